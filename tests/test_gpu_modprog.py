"""mab_<P>_modprog: a sequence of generated-API calls executed in one launch with the variables on chip.  Every
program is checked three ways: against the same sequence issued as separate API calls (bit-identical limb planes),
against the value-level oracle, and -- for the reference's own complete point addition -- against the group law."""
import random

import numpy as np
import pytest
import torch

from field_oracle import FieldOracle
from modarith_b200.primes import ALL_PRIMES as PRIMES

pytestmark = pytest.mark.gpu
NAMES = list(PRIMES)


def _field(name):
    from modarith_b200 import Field
    return Field(name)


def run_separately(F, code, inputs, nreg=16):
    """The same program as one API call per instruction."""
    n = inputs[0].shape[1]
    R = [F.alloc(n) for _ in range(nreg)]
    for r in R:
        F.modzer(r)
    for k, t in enumerate(inputs):
        R[k].copy_(t)
    for ins in code:
        op, d, a, b = ins[:4]
        imm = ins[4] if len(ins) > 4 else 0
        if op == "add": F.modadd(R[a], R[b], R[d])
        elif op == "sub": F.modsub(R[a], R[b], R[d])
        elif op == "neg": F.modneg(R[a], R[d])
        elif op == "mul": F.modmul(R[a], R[b], R[d])
        elif op == "sqr": F.modsqr(R[a], R[d])
        elif op == "mli": F.modmli(R[a], imm, R[d])
        elif op == "cpy": F.modcpy(R[a], R[d])
        elif op == "nsqr": F.modcpy(R[a], R[d]); F.modnsqr(R[d], imm)
        elif op == "pro": F.modpro(R[a], R[d])
        elif op == "inv": F.modinv_perelement(R[a], R[d])
        elif op == "sqrt": F.modsqrt(R[a], None, R[d])
        elif op == "zer": F.modzer(R[d])
        elif op == "one": F.modone(R[d])
        elif op == "int": F.modint(imm, R[d])
        elif op == "haf": F.modcpy(R[a], R[d]); F.modhaf(R[d])
    return R


@pytest.mark.parametrize("name", NAMES)
def test_random_programs_against_separate_calls_and_oracle(name):
    F = _field(name)
    O = FieldOracle(name)
    p = O.p
    rng = random.Random(77)
    n = 300
    vals = [[rng.randrange(p) for _ in range(n)] for _ in range(3)]
    for v in vals:
        v[0], v[1], v[2] = 0, 1, p - 1
    inputs = [F.from_ints(v) for v in vals]
    ops2 = ["add", "sub", "mul"]
    ops1 = ["neg", "sqr", "cpy", "haf"]
    for trial in range(6):
        code, live = [], 3
        for _ in range(40):
            r = rng.random()
            d = rng.randrange(16)
            if r < 0.6:
                code.append((rng.choice(ops2), d, rng.randrange(live), rng.randrange(live)))
            elif r < 0.8:
                code.append((rng.choice(ops1), d, rng.randrange(live), 0))
            elif r < 0.87:
                code.append(("mli", d, rng.randrange(live), 0, rng.choice([0, 1, 2, 121665, 39081, (1 << 31) - 1])))
            elif r < 0.92:
                code.append(("nsqr", d, rng.randrange(live), 0, rng.randrange(0, 5)))
            elif r < 0.96:
                code.append((rng.choice(["zer", "one"]), d, 0, 0))
            else:
                code.append(("int", d, 0, 0, rng.randrange(1 << 20)))
            live = max(live, min(16, d + 1)) if d <= live else live
        if trial == 0:
            code += [("inv", 5, 0, 0), ("sqrt", 6, 1, 0), ("pro", 7, 2, 0)]
        outs = list(range(16))
        got = F.modprog(code, inputs, outs)
        ref = run_separately(F, code, inputs)
        for r in range(16):
            assert torch.equal(F.modexp(got[r]), F.modexp(ref[r])), (name, trial, r)
        # value-level oracle on a few elements
        for i in (0, 1, 2, 17, n - 1):
            R = [0] * 16
            for k in range(3):
                R[k] = vals[k][i]
            for ins in code:
                op, d, a, b = ins[:4]
                imm = ins[4] if len(ins) > 4 else 0
                R[d] = {"add": lambda: O.modadd(R[a], R[b]), "sub": lambda: O.modsub(R[a], R[b]), "mul": lambda: O.modmul(R[a], R[b]),
                        "neg": lambda: O.modneg(R[a]), "sqr": lambda: O.modsqr(R[a]), "cpy": lambda: R[a], "haf": lambda: O.modhaf(R[a]),
                        "mli": lambda: O.modmli(R[a], imm), "nsqr": lambda: O.modnsqr(R[a], imm), "zer": lambda: 0, "one": lambda: 1,
                        "int": lambda: imm % p, "inv": lambda: O.modinv(R[a]), "sqrt": lambda: O.modsqrt(R[a]),
                        "pro": lambda: O.modpro(R[a])}[op]() % p
            ints = [F.to_ints(got[r][:, i:i + 1].contiguous())[0] for r in range(16)]
            assert ints == R, (name, trial, i)


@pytest.mark.parametrize("jit", [False, True], ids=["interpreted", "compiled"])
def test_weierstrass_addition_as_one_program(jit):
    """The complete projective addition for a = -3 (eprint 2015/1060 Algorithm 4, which weierstrass.c:69-160 transcribes)
    written as a modprog program over 12 registers: 12 multiplications, 2 by the curve constant b, 29 additions /
    subtractions -- one launch instead of 43.  Checked against affine arithmetic on the curve."""
    F = _field("NIST256")
    P = PRIMES["NIST256"]
    p, b = P.p, P.wb
    O = FieldOracle("NIST256")

    def affine_add(P1, P2):
        if P1 is None: return P2
        if P2 is None: return P1
        (x1, y1), (x2, y2) = P1, P2
        if x1 == x2 and (y1 + y2) % p == 0: return None
        lam = (3 * x1 * x1 - 3) * pow(2 * y1, -1, p) % p if P1 == P2 else (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        return x3, (lam * (x1 - x3) - y1) % p

    def mul(k, Pt):
        R = None
        while k:
            if k & 1: R = affine_add(R, Pt)
            Pt = affine_add(Pt, Pt)
            k >>= 1
        return R

    G = (P.wgx, P.wgy)
    rng = random.Random(5)
    n = 64
    pts1 = [mul(rng.randrange(1, P.worder), G) for _ in range(n)]
    pts2 = [mul(rng.randrange(1, P.worder), G) for _ in range(n)]
    pts2[0] = pts1[0]                                          # doubling through the addition law (complete formulas)
    pts2[1] = (pts1[1][0], (-pts1[1][1]) % p)                  # P + (-P) = O
    X1, Y1 = F.from_ints([q[0] for q in pts1]), F.from_ints([q[1] for q in pts1])
    X2, Y2 = F.from_ints([q[0] for q in pts2]), F.from_ints([q[1] for q in pts2])
    Bc = F.from_ints([b] * n)
    one = F.alloc(n); F.modone(one)
    # registers: 0 X1, 1 Y1, 2 Z1, 3 X2, 4 Y2, 5 Z2, 6 b, 7..11 t0..t4, 12 X3, 13 Y3, 14 Z3
    X1_, Y1_, Z1_, X2_, Y2_, Z2_, B_, t0, t1, t2, t3, t4, X3, Y3, Z3 = range(15)
    code = [("mul", t0, X1_, X2_), ("mul", t1, Y1_, Y2_), ("mul", t2, Z1_, Z2_), ("add", t3, X1_, Y1_), ("add", t4, X2_, Y2_),
            ("mul", t3, t3, t4), ("add", t4, t0, t1), ("sub", t3, t3, t4), ("add", t4, Y1_, Z1_), ("add", X3, Y2_, Z2_),
            ("mul", t4, t4, X3), ("add", X3, t1, t2), ("sub", t4, t4, X3), ("add", X3, X1_, Z1_), ("add", Y3, X2_, Z2_),
            ("mul", X3, X3, Y3), ("add", Y3, t0, t2), ("sub", Y3, X3, Y3), ("mul", Z3, B_, t2), ("sub", X3, Y3, Z3),
            ("add", Z3, X3, X3), ("add", X3, X3, Z3), ("sub", Z3, t1, X3), ("add", X3, t1, X3), ("mul", Y3, B_, Y3),
            ("add", t1, t2, t2), ("add", t2, t1, t2), ("sub", Y3, Y3, t2), ("sub", Y3, Y3, t0), ("add", t1, Y3, Y3),
            ("add", Y3, t1, Y3), ("add", t1, t0, t0), ("add", t0, t1, t0), ("sub", t0, t0, t2), ("mul", t1, t4, Y3),
            ("mul", t2, t0, Y3), ("mul", Y3, X3, Z3), ("add", Y3, Y3, t2), ("mul", X3, t3, X3), ("sub", X3, X3, t1),
            ("mul", Z3, t4, Z3), ("mul", t1, t3, t0), ("add", Z3, Z3, t1),
            # affine: x = X3/Z3, y = Y3/Z3 (0 -> 0 gives (0, 0) for the point at infinity)
            ("inv", t0, Z3, 0), ("mul", X3, X3, t0), ("mul", Y3, Y3, t0)]
    outs = F.modprog(code, [X1, Y1, one, X2, Y2, one, Bc], [X3, Y3, Z3], jit=jit)
    xs, ys, zs = F.to_ints(outs[0]), F.to_ints(outs[1]), F.to_ints(outs[2])
    for i in range(n):
        want = affine_add(pts1[i], pts2[i])
        if want is None:
            assert zs[i] == 0
        else:
            assert (xs[i], ys[i]) == want, i
    assert zs[1] == 0 and (xs[0], ys[0]) == affine_add(pts1[0], pts1[0])


@pytest.mark.parametrize("jit", [False, True], ids=["interpreted", "compiled"])
def test_modprog_arguments(jit):
    F = _field("X25519")
    x = F.from_ints([1, 2, 3])
    with pytest.raises(ValueError):
        F.modprog([], [x], [0], jit=jit)
    with pytest.raises(ValueError):
        F.modprog([("mul", 16, 0, 0)], [x], [0], jit=jit)
    with pytest.raises(ValueError):
        F.modprog([("frobnicate", 1, 0, 0)], [x], [0], jit=jit)
    with pytest.raises(ValueError):
        F.modprog([("mli", 1, 0, 0, -1)], [x], [0], jit=jit)
    with pytest.raises(ValueError):
        F.modprog([("mul", 1, 0, 0)] * 400, [x], [1], jit=jit)
    out = F.modprog([("mul", 1, 0, 0), ("add", 1, 1, 0)], [x], [1], jit=jit)
    assert F.to_ints(out[0]) == [2, 6, 12]
    # in place: the output tensor is an input
    F.modprog([("sqr", 0, 0, 0)], [x], [0], outputs=[x], jit=jit)
    assert F.to_ints(x) == [1, 4, 9]
    # ragged size across several blocks, pitch wider than n
    n = 1000
    wide = torch.zeros((F.Nlimbs, 3 * n), dtype=torch.int32, device="cuda")
    a, c = wide[:, :n], wide[:, 2 * n:]
    a.copy_(F.from_ints(list(range(n))))
    F.modprog([("mli", 1, 0, 0, 3), ("add", 1, 1, 0)], [a], [1], outputs=[c], jit=jit)
    assert F.to_ints(c.contiguous()) == [4 * v for v in range(n)]
    assert int(wide[:, n:2 * n].abs().sum()) == 0


def _random_program(rng, length, heavy):
    ops2, ops1 = ["add", "sub", "mul"], ["neg", "sqr", "cpy", "haf"]
    code, live = [], 3
    for _ in range(length):
        r = rng.random()
        d = rng.randrange(16)
        if r < 0.6:
            code.append((rng.choice(ops2), d, rng.randrange(live), rng.randrange(live)))
        elif r < 0.8:
            code.append((rng.choice(ops1), d, rng.randrange(live), 0))
        elif r < 0.87:
            code.append(("mli", d, rng.randrange(live), 0, rng.choice([0, 1, 2, 121665, 39081, (1 << 31) - 1])))
        elif r < 0.92:
            code.append(("nsqr", d, rng.randrange(live), 0, rng.randrange(0, 5)))
        elif r < 0.96:
            code.append((rng.choice(["zer", "one"]), d, 0, 0))
        else:
            code.append(("int", d, 0, 0, rng.randrange(1 << 20)))
        live = max(live, min(16, d + 1)) if d <= live else live
    if heavy:
        code += [("inv", 5, 0, 0), ("sqrt", 6, 1, 0), ("pro", 7, 2, 0)]
    return code


@pytest.mark.parametrize("name", NAMES)
def test_compiled_programs_equal_interpreted_ones(name):
    """mab_<P>_modprog_jit against mab_<P>_modprog on random programs: the same generated functions called in the same
    order, so the limb planes -- not just the canonical values -- are bit-identical.  Registers that are never written
    read as zero in both; a program is compiled once and then served from the cache (second call, other batch size)."""
    F = _field(name)
    p = FieldOracle(name).p
    rng = random.Random(2024)
    n = 1000
    vals = [[rng.randrange(p) for _ in range(n)] for _ in range(3)]
    for v in vals:
        v[0], v[1], v[2] = 0, 1, p - 1
    inputs = [F.from_ints(v) for v in vals]
    for trial in range(2):
        # the long fixed exponentiations inline ~270 products each; on the fall-back plan (500 instructions per
        # product) that is a minute of compilation, so the group order gets the light programs only
        code = _random_program(rng, 30, heavy=(trial == 0 and name != "NIST256ORDER"))
        outs = list(range(16))
        want = F.modprog(code, inputs, outs)
        got = F.modprog(code, inputs, outs, jit=True)
        for r in range(16):
            assert torch.equal(got[r], want[r]), (name, trial, r)
        small = [t[:, :77].contiguous() for t in inputs]
        got2 = F.modprog(code, small, outs, jit=True)
        for r in range(16):
            assert torch.equal(got2[r], want[r][:, :77]), (name, trial, r, "cached kernel, other n")
