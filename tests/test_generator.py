"""The generator backend: limb plans verified by the PTX interpreter, emitted headers reproducible."""
import os
import random

import pytest

from modarith_b200.primes import ALL_PRIMES as PRIMES, Prime
from modarith_b200.gen.plan import (make_plan, PseudoMersenne, PseudoMersenne33, PseudoMersenneBits, GenMersenne, Montgomery,
                                    MontgomeryFriendly, MontgomeryFull, words, value)
from modarith_b200.gen.ptx import Asm, LostCarry
from modarith_b200.gen import satmul
from modarith_b200.gen.emit import emit_field_header

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_plan_selection():
    assert isinstance(make_plan(PRIMES["X25519"]), PseudoMersenne)
    assert isinstance(make_plan(PRIMES["X448"]), GenMersenne)
    assert isinstance(make_plan(PRIMES["NIST256"]), Montgomery)
    assert make_plan(PRIMES["X25519"]).L == 8 and make_plan(PRIMES["X448"]).L == 14


@pytest.mark.parametrize("name", list(PRIMES))
def test_plan_self_check(name):
    """mul/sqr/mli/add/sub/neg/canon blocks against bignum arithmetic, lost carries trapped."""
    plan = make_plan(PRIMES[name])
    plan.build()
    assert plan.self_check(trials=250, seed=2024)


@pytest.mark.parametrize("name", list(PRIMES))
def test_product_counts(name):
    """Wide multiplies per call = algorithmic L^2 / L(L+1)/2 (+ fold multiplies for 2^n-c)."""
    plan = make_plan(PRIMES[name])
    b = plan.build()
    L = plan.L
    fold = L if name == "X25519" else (L + 1 if isinstance(plan, PseudoMersenne33) else 0)
    if isinstance(plan, MontgomeryFull):      # separated-operand REDC: + L(L+1)/2 - L (low half) + L^2 (Q*p)
        assert b["mul"].stats()[0] <= L * L + L * (L + 1) // 2 + L * L
        return
    assert b["mul"].stats()[0] == L * L + fold
    assert b["sqr"].stats()[0] == L * (L + 1) // 2 + fold
    assert b["add"].stats()[0] == 0 and b["sub"].stats()[0] == 0


def test_interpreter_traps_lost_carry():
    a = Asm("t")
    a.inp("x", "y")
    d = a.tmp()
    a.add(d, "x", "y")               # no carry-out declared
    a.out("r", d)
    assert a.run({"x": 1, "y": 2})["r"] == 3
    with pytest.raises(LostCarry):
        a.run({"x": 0xFFFFFFFF, "y": 1})


@pytest.mark.parametrize("L", [2, 4, 8, 14])
def test_wide_products(L):
    rng = random.Random(L)
    A = Asm("mul")
    a = ["a[%d]" % i for i in range(L)]
    b = ["b[%d]" % i for i in range(L)]
    A.inp(*a)
    A.inp(*b)
    for k, t in enumerate(satmul.product(A, a, b)):
        A.out("r[%d]" % k, t)
    S = Asm("sqr")
    S.inp(*a)
    for k, t in enumerate(satmul.square(S, a)):
        S.out("r[%d]" % k, t)
    top = (1 << (32 * L)) - 1
    cases = [(top, top), (0, top), (1, 1)] + [(rng.getrandbits(32 * L), rng.getrandbits(32 * L)) for _ in range(60)]
    for x, y in cases:
        env = {a[i]: w for i, w in enumerate(words(x, L))}
        env.update({b[i]: w for i, w in enumerate(words(y, L))})
        o = A.run(env)
        assert value([o["r[%d]" % k] for k in range(2 * L)]) == x * y
        o = S.run(env)
        assert value([o["r[%d]" % k] for k in range(2 * L)]) == x * x
    assert A.stats()[0] == L * L and S.stats()[0] == L * (L + 1) // 2


@pytest.mark.parametrize("name", list(PRIMES))
def test_emitted_header_is_reproducible(name):
    """The committed csrc/gen/field_<P>.cuh is exactly what the generator prints today."""
    text = emit_field_header(make_plan(PRIMES[name]))
    path = os.path.join(ROOT, "modarith_b200", "csrc", "gen", "field_%s.cuh" % name)
    assert open(path).read() == text
    assert "asm(" in text and "MAB_HOSTSIM" in text and "madc.hi.cc.u32" in text


def test_fallback_plan_accepts_any_odd_modulus():
    """Every kind of prime the reference's named tables hold gets a plan that passes the bignum
    self-check: odd limb counts, unshaped Montgomery moduli, group orders (monty.py:1961-2127)."""
    assert isinstance(make_plan(PRIMES["SECP256K1"]), PseudoMersenne33)       # 2^256 - 2^32 - 977: shaped after all
    assert MontgomeryFull(PRIMES["SECP256K1"]).build() and MontgomeryFull(PRIMES["SECP256K1"]).L == 8
    assert isinstance(make_plan(PRIMES["NIST256ORDER"]), MontgomeryFull)
    for nm, p in {"NIST384": 2**384 - 2**128 - 2**96 + 2**32 - 1, "NIST521": 2**521 - 1, "PM266": 2**266 - 3,
                  "NIST224": 2**224 - 2**96 + 1, "GM480": 2**480 - 2**240 - 1,
                  "ED25519ORDER": 2**252 + 27742317777372353535851937790883648493}.items():
        plan = make_plan(Prime(nm, p, "monty"))
        assert plan.self_check(trials=40, seed=7)
    for nm, p in {"PM383": 2**383 - 187, "PM512": 2**512 - 569}.items():
        plan = make_plan(Prime(nm, p, "pseudo"))
        assert isinstance(plan, PseudoMersenne) and plan.self_check(trials=40, seed=7)


@pytest.mark.parametrize("nm,expr", [("C41417", "2**414-17"), ("NIST521", "2**521-1"), ("PM266", "2**266-3"),
                                     ("C2065", "2**206-5"), ("PM336", "2**336-3"), ("M221", "2**221-3")])
def test_bit_level_pseudo_mersenne_plan(nm, expr):
    """2^n - c with n not a multiple of 32 and any limb count (most of pseudo.py's named table, pseudo.py:1487-1550):
    folded at bit n, L^2 + L wide multiplies per modmul where the fall-back plan needs 2 L^2 + L.  Stored values stay
    below 2^n + 2^32; canon covers raw imports up to 2^(32L) (checked by self_check over that whole range)."""
    p = eval(expr)
    plan = make_plan(Prime(nm, p, "pseudo"))
    assert isinstance(plan, PseudoMersenneBits)
    L = plan.L
    assert plan.bound == (1 << p.bit_length()) + (1 << 32) and plan.R == 1
    b = plan.blocks
    assert b["mul"].stats()[0] == L * L + L and b["sqr"].stats()[0] == L * (L + 1) // 2 + L
    assert b["add"].stats()[0] == 0 and b["sub"].stats()[0] == 0
    for seed in (3, 4):
        assert plan.self_check(trials=300, seed=seed)
    # the fall-back plan still takes the same modulus (comparison builds: MAB_PMBITS=0)
    assert MontgomeryFull(Prime(nm, p, "monty")).build()


@pytest.mark.parametrize("nm,expr,z", [("ED248", "5*2**248-1", 7), ("MFP4", "3*67*(2**246)-1", 7), ("SIDH434", "2**216*3**137-1", 6),
                                       ("GM270", "2**270-2**162-1", 5), ("GM512", "2**512-2**127-1", 3),
                                       ("NIST384", "2**384-2**128-2**96+2**32-1", 1)])
def test_montgomery_friendly_plan(nm, expr, z):
    """p = -1 (mod 2^(32z)): the quotient digits are the low words themselves and only the L - z high words of
    (p + 1) / 2^(32z) are multiplied -- L^2 + L(L - z) wide multiplies, no plain ones (monty.py:740-751 is the
    one-word case; isogeny, MFP and 2^n - 2^m - 1 primes of monty.py:1961-2108 have z up to L - 1)."""
    p = eval(expr)
    plan = make_plan(Prime(nm, p, "monty"))
    assert isinstance(plan, MontgomeryFriendly) and plan.z == z
    L = plan.L
    assert plan.blocks["mul"].stats()[:2] == (L * L + L * (L - z), 0)
    assert plan.self_check(trials=120, seed=5)
    full = MontgomeryFull(Prime(nm, p, "monty"))
    full.build()
    assert full.blocks["mul"].stats()[0] == 2 * L * L


def test_carry_capture_pipe_table(monkeypatch):
    """Which pipe a function's carry captures run on is a per-plan table (Plan.capture_ops): the ALU-bound P-256 plan
    writes them as madc.lo d, 0, 0, 0 (IMAD.X) everywhere but in mul_w, the multiplier-bound plans leave `addc d, 0, 0`
    to ptxas; MAB_CAPOP_<PRIME>_<FN> overrides one function (profiles/r2_p256_capop.txt, r2_ecn_capop.txt)."""
    def captures(asm):
        lines, _, _ = asm.emit_ptx()
        return (sum(1 for l in lines if l.startswith("madc.lo.u32") and l.endswith(", 0, 0, 0;")),
                sum(1 for l in lines if l.startswith("addc.u32") and l.endswith(", 0x0, 0x0;")))

    p256 = make_plan(PRIMES["NIST256"])
    b = p256.build()
    for fn in ("mul", "sqr", "add", "sub", "sqr_w"):
        assert b[fn].capture_op == "madc"
    assert b["mul_w"].capture_op == "addc"
    m, a = captures(b["sqr_w"])
    assert m >= 5 and a == 0
    m, a = captures(b["mul_w"])
    assert m == 0 and a >= 5
    for name in ("X25519", "X448", "SECP256K1", "NIST256ORDER"):
        for fn, asm in make_plan(PRIMES[name]).build().items():
            assert asm.capture_op == "addc", (name, fn)
            assert captures(asm)[0] == 0
    monkeypatch.setenv("MAB_CAPOP_NIST256_SQR_W", "addc")
    monkeypatch.setenv("MAB_CAPOP_X25519_SQR", "madc")
    assert make_plan(PRIMES["NIST256"]).build()["sqr_w"].capture_op == "addc"
    x = make_plan(PRIMES["X25519"]).build()
    assert x["sqr"].capture_op == "madc" and x["mul"].capture_op == "addc"
    # the interpreter sees the same instruction either way: results do not depend on the table
    assert make_plan(PRIMES["NIST256"]).build() and p256.self_check(trials=50, seed=7)


def test_ladder_launch_bounds():
    """8-limb ladders are compiled for three resident CTAs per SM (ptxas then takes the registers it wants:
    profiles/r2_x25519_occupancy.txt), longer ones for two; MAB_MINBLOCKS_<PRIME> overrides."""
    assert make_plan(PRIMES["X25519"]).ladder_minblocks == 3
    assert make_plan(PRIMES["X448"]).ladder_minblocks == 2
    hdr = open(os.path.join(ROOT, "modarith_b200", "csrc", "gen", "field_X25519.cuh")).read()
    assert "LADDER_MINBLOCKS = 3;" in hdr
