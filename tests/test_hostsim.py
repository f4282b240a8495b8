"""Device-side logic on the CPU: the generated headers + hand-written templates compiled with the
PTX blocks transcribed to C (tests/hostsim), checked against the oracle and the golden vectors.
(The real PTX build is checked by the -m gpu tests through the C ABI.)"""
import ctypes
import random

import numpy as np
import pytest

from field_oracle import FieldOracle, rfc7748
from modarith_b200.primes import ALL_PRIMES as PRIMES
import util

CURVES = ("X25519", "X448")


@pytest.mark.parametrize("curve", CURVES)
def test_ladder_golden(hostsim, golden_rfc, curve):
    nb = golden_rfc[curve]["nbytes"]
    fn = getattr(hostsim, "sim_%s_rfc7748" % curve)
    rows = golden_rfc[curve]["edge"] + golden_rfc[curve]["random"]
    v = golden_rfc[curve]["rfc"]
    g = PRIMES[curve].generator.to_bytes(nb, "little")
    rows = rows + [{"k": v["sk1"], "u": g.hex(), "out": v["pk1"]}, {"k": v["sk2"], "u": g.hex(), "out": v["pk2"]},
                   {"k": v["sk1"], "u": v["pk2"], "out": v["shared"]}]
    for r in rows:
        out = ctypes.create_string_buffer(nb)
        fn(bytes.fromhex(r["k"]), bytes.fromhex(r["u"]), out)
        assert out.raw[:nb].hex() == r["out"], r


@pytest.mark.parametrize("curve", CURVES)
def test_ladder_validation_tail(hostsim, golden_rfc, curve):
    """The ladder variant with the point-validation tail (rfc7748.c:228-251)."""
    nb = golden_rfc[curve]["nbytes"]
    fn = getattr(hostsim, "sim_%s_rfc7748_validate" % curve)
    for r in golden_rfc[curve]["validate"]:
        out = ctypes.create_string_buffer(nb)
        fn(bytes.fromhex(r["k"]), bytes.fromhex(r["u"]), out)
        assert out.raw[:nb].hex() == r["out"], r


@pytest.mark.parametrize("curve", CURVES)
def test_shared_inversion(hostsim, golden_rfc, curve):
    """K ladders finished with ONE inversion (Rfc7748<F>::finish_batch): same bytes as one inversion per
    key, including keys whose z2 is zero (low-order inputs: u = 0, 1, p-1 ...) mixed into the batch."""
    nb = golden_rfc[curve]["nbytes"]
    fn = getattr(hostsim, "sim_%s_rfc7748_shared" % curve)
    rows = golden_rfc[curve]["edge"] + golden_rfc[curve]["random"][:13]
    zero_rows = [r for r in rows if int(r["out"], 16) == 0]
    assert zero_rows, "the edge set must contain low-order inputs"
    rng = random.Random(4)
    for K in (1, 2, 3, 4):
        for _ in range(6):
            pick = rng.sample(rows, K)
            if K > 1:
                pick[rng.randrange(K)] = rng.choice(zero_rows)
            k = b"".join(bytes.fromhex(r["k"]) for r in pick)
            u = b"".join(bytes.fromhex(r["u"]) for r in pick)
            out = ctypes.create_string_buffer(nb * K)
            fn(k, u, out, K)
            assert [out.raw[nb * j:nb * (j + 1)].hex() for j in range(K)] == [r["out"] for r in pick], (curve, K)
    # all keys zero
    pick = [zero_rows[0]] * 4
    out = ctypes.create_string_buffer(nb * 4)
    fn(b"".join(bytes.fromhex(r["k"]) for r in pick), b"".join(bytes.fromhex(r["u"]) for r in pick), out, 4)
    assert out.raw[:nb * 4] == bytes(nb * 4)


def test_ladder_demo_loop(hostsim, golden_rfc):
    """rfc7748.c:main's 5000x2 chained loop (X25519): every output feeds the next call."""
    d = golden_rfc["X25519"]["demo"]
    bk = bytes.fromhex(d["key"])
    bu = (9).to_bytes(32, "little")
    out = ctypes.create_string_buffer(32)
    for _ in range(5000):
        hostsim.sim_X25519_rfc7748(bk, bu, out)
        bv = out.raw[:32]
        hostsim.sim_X25519_rfc7748(bk, bv, out)
        bu = out.raw[:32]
    assert bu.hex() == d["loop5000"]


@pytest.mark.parametrize("curve", CURVES)
def test_ladder_random_vs_oracle(hostsim, curve):
    nb = PRIMES[curve].nbytes
    n = 24
    k, u = util.random_bytes(31, n, nb), util.random_bytes(32, n, nb)
    out = np.zeros_like(k)
    getattr(hostsim, "sim_%s_rfc7748_batch" % curve)(k.ctypes.data_as(ctypes.c_void_p), u.ctypes.data_as(ctypes.c_void_p),
                                                      out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n))
    for i in range(n):
        assert out[i].tobytes() == rfc7748(curve, k[i].tobytes(), u[i].tobytes())


@pytest.mark.parametrize("curve", CURVES)
def test_ladder_saturated_word_patterns(hostsim, curve):
    """Scalars and u-coordinates made of all-ones / all-zero / single-bit words: the operands most likely to
    break a bound the ladder step relies on (the 2^255-19 step adds and subtracts products with add_tt /
    sub_tt, whose preconditions the host simulation checks on every call and aborts on)."""
    nb = PRIMES[curve].nbytes
    rng = np.random.default_rng(255)
    pats = np.array([0x00, 0xff, 0x80, 0x01, 0x7f, 0xfe], dtype=np.uint8)
    n = 48
    k = np.repeat(pats[rng.integers(0, len(pats), (n, nb // 4))], 4, axis=1).astype(np.uint8)
    u = np.repeat(pats[rng.integers(0, len(pats), (n, nb // 4))], 4, axis=1).astype(np.uint8)
    k[:4] = 0xff
    u[:2] = 0xff
    u[2:4] = 0
    p = PRIMES[curve].p
    for i, v in enumerate((p - 1, p + 1, 2 * p - 1 if curve == "X25519" else p - 2, (1 << (8 * nb - 1)) - 1)):
        u[4 + i] = np.frombuffer((v % (1 << (8 * nb))).to_bytes(nb, "little"), dtype=np.uint8)
    out = np.zeros_like(k)
    getattr(hostsim, "sim_%s_rfc7748_batch" % curve)(k.ctypes.data_as(ctypes.c_void_p), u.ctypes.data_as(ctypes.c_void_p),
                                                      out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n))
    for i in range(n):
        assert out[i].tobytes() == rfc7748(curve, k[i].tobytes(), u[i].tobytes()), i


@pytest.mark.parametrize("name", list(PRIMES))
def test_field_golden(hostsim, golden_field, name):
    g = golden_field[name]
    S = util.Sim(hostsim, name)
    for i, (ah, bh) in enumerate(zip(g["a"], g["b"])):
        a, b = int(ah, 16), int(bh, 16)
        sa, lt = S.imp(a)
        sb, _ = S.imp(b)
        assert lt == g["ops"]["id"]["status"][i]
        res = {
            "mul": S.raw("MUL", sa, sb)[0], "sqr": S.raw("SQR", sa)[0], "inv": S.raw("INV", sa)[0],
            "sqrt": S.raw("SQRT", sa)[0], "add": S.raw("ADD", sa, sb)[0], "sub": S.raw("SUB", sa, sb)[0],
            "neg": S.raw("NEG", sa)[0], "pro": S.raw("PRO", sa)[0], "id": sa,
            "mli": S.raw("MLI", sa, scalar=g["mli_int"])[0], "haf": S.raw("HAF", sa)[0],
        }
        for op, s in res.items():
            assert "%0*x" % (2 * g["nbytes"], S.exp(s)) == g["ops"][op]["out"][i], (name, op, i)
        assert S.raw("QR", sa)[1] == g["ops"]["qr"]["status"][i]


@pytest.mark.parametrize("name", list(PRIMES))
def test_every_api_function_vs_oracle(hostsim, name):
    S = util.Sim(hostsim, name)
    F = FieldOracle(name)
    p = F.p
    rng = random.Random(9)
    vals = [0, 1, 2, p - 1, p - 2, (p + 1) // 2, 4, 9] + [rng.randrange(p) for _ in range(12)]
    for x in vals:
        y = rng.randrange(p)
        sx, sy = S.imp(x)[0], S.imp(y)[0]
        assert S.exp(sx) == x
        h = S.raw("PRO", sx)[0]
        assert S.exp(h) == F.modpro(x)
        assert S.exp(S.raw("INVH", sx, h)[0]) == F.modinv(x) == (pow(x, -1, p) if x else 0)
        assert S.exp(S.raw("SQRTH", sx, h)[0]) == F.modsqrt(x)
        assert S.raw("QRH", h, sx)[1] == F.modqr(None, x) == int(x == 0 or pow(x, (p - 1) // 2, p) == 1)
        sq = F.modsqrt(F.modsqr(x))
        assert sq in (x, p - x)
        assert S.raw("IS1", sx)[1] == int(x == 1) and S.raw("IS0", sx)[1] == int(x == 0)
        assert S.exp(S.raw("NSQR", sx, scalar=5)[0]) == pow(x, 32, p)
        assert S.exp(S.raw("ONE")[0]) == 1
        assert S.exp(S.raw("INT", scalar=x & 0x7FFFFFFF)[0]) == (x & 0x7FFFFFFF) % p
        assert S.exp(S.raw("NRES", x)[0]) == x          # nres takes the plain canonical words
        assert S.raw("REDC", sx)[0] == x                # redc returns them
        r, _, r2 = S.raw("CSW", sx, sy, scalar=1)
        assert (S.exp(r), S.exp(r2)) == (y, x)
        r, _, r2 = S.raw("CSW", sx, sy, scalar=0)
        assert (S.exp(r), S.exp(r2)) == (x, y)
        assert S.exp(S.raw("CMV", sx, sy, scalar=1)[0]) == x and S.exp(S.raw("CMV", sx, sy, scalar=0)[0]) == y
        assert S.exp(S.raw("SHL", sx, scalar=3)[0]) == F.modshl(3, x)
        assert S.exp(S.raw("HAF", sx)[0]) == F.modhaf(x)
        assert S.raw("SIGN", sx)[1] == F.modsign(x)
        assert S.raw("CMP", sx, sy)[1] == F.modcmp(x, y) and S.raw("CMP", sx, sx)[1] == 1
        st, lt, _ = S.raw("FSB", sx)
        assert lt == 1
    for r in (0, 1, 31, 32, 100, F.P.nbits - 1, 8 * F.nbytes - 1, 8 * F.nbytes, 8 * F.nbytes + 5):
        assert S.exp(S.raw("2R", scalar=r)[0]) == F.mod2r(r)
    # modshr is defined on the canonical STORED value (plain value for the non-Montgomery moduli)
    if name in ("X25519", "X448"):
        for x in vals:
            r, out, _ = S.raw("SHR", S.imp(x)[0], scalar=8)
            assert (S.exp(r), out) == F.modshr(8, x)


@pytest.mark.parametrize("name", list(PRIMES))
def test_reference_selftest_sequence(hostsim, name):
    """The generator self-test call sequence (pseudo.py:1783-1796) on the simulated device code;
    modshl/modshr excepted -- see DESIGN.md on saturated limbs."""
    S = util.Sim(hostsim, name)
    p = PRIMES[name].p
    rng = random.Random(77)
    for _ in range(10):
        x, y = rng.randrange(p), rng.randrange(p)
        ax, ay = S.raw("NRES", x)[0], S.raw("NRES", y)[0]
        t = S.raw("ADD", ax, ay)[0]
        z = S.raw("SUB", ax, ay)[0]
        ax = S.raw("MUL", t, z)[0]
        z = S.raw("SQR", ax)[0]
        z = S.raw("INV", z)[0]
        z = S.raw("SQRT", z)[0]
        z = S.raw("SQR", z)[0]
        z = S.raw("HAF", z)[0]
        z = S.raw("ADD", z, z)[0]
        z = S.raw("REDC", z)[0]
        assert z == pow(((x - y) * (x + y)) ** 2 % p, -1, p)
