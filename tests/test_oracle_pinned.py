"""Pin the oracle (oracle/field_oracle.py) before trusting it: every golden vector, known-answer
test and fixture the reference holds for this path."""
import os
import random

import numpy as np
import pytest

from field_oracle import FieldOracle, rfc7748
from modarith_b200.primes import PRIMES, ALL_PRIMES
from modarith_b200 import addchain
import util

CURVES = ("X25519", "X448")


def test_rfc7748_section6_vectors(golden_rfc):
    """rfc7748.c:271,274 + simd/rfc7748_simt.cu:245,249; expected values are RFC 7748 6.1/6.2."""
    expect = {
        "X25519": ("8520f0098930a754748b7ddcb43ef75a0dbf3a0d26381af4eba4a98eaa9b4e6a",
                   "de9edb7d7b7dc1b4d35b61c2ece435373f8343c85b78674dadfc7e146f882b4f",
                   "4a5d9d5ba4ce2de1728e3bf480350f25e07e21c947d19e3376f09b3c1e161742"),
        "X448": ("9b08f7cc31b7e3e67d22d5aea121074a273bd2b83de09c63faa73d2c22c5d9bbc836647241d953d40c5b12da88120d53177f80e532c41fa0",
                 "3eb7a829b0cd20f5bcfc0b599b6feccf6da4627107bdb0d4f345b43027d8b972fc3e34fb4232a13ca706dcb57aec3dae07bdc1c67bf33609",
                 "07fff4181ac6cc95ec1c16a94a0f74d12da232ce40a77552281d282bb60c0b56fd2464c335543936521c24403085d59a449a5037514a879d"),
    }
    for c in CURVES:
        v = golden_rfc[c]["rfc"]
        nb = golden_rfc[c]["nbytes"]
        g = PRIMES[c].generator.to_bytes(nb, "little")
        k1, k2 = bytes.fromhex(v["sk1"]), bytes.fromhex(v["sk2"])
        pk1, pk2 = rfc7748(c, k1, g), rfc7748(c, k2, g)
        assert (pk1.hex(), pk2.hex()) == expect[c][:2] == (v["pk1"], v["pk2"])
        assert rfc7748(c, k1, pk2).hex() == rfc7748(c, k2, pk1).hex() == expect[c][2] == v["shared"]


def test_rfc7748_iterated_vector():
    """RFC 7748 5.2: one iteration of k <- X(k, u), u <- old k starting from the base point."""
    k = u = (9).to_bytes(32, "little")
    k, u = rfc7748("X25519", k, u), k
    assert k.hex() == "422c8e7a6227d7bca1350b3e2bb7279f7897b87bb6854b783c60e80311ae3079"
    k = u = (5).to_bytes(56, "little")
    k, u = rfc7748("X448", k, u), k
    assert k.hex() == ("3f482c8a9f19b01e6c46ee9711d9dc14fd4bf67af30765c2ae2b846a4d23a8cd0db897086239492caf350b51f833868b9bc2b3bca9cf4113")


def test_demo_main_outputs(golden_rfc):
    """Deterministic outputs of rfc7748.c:main (LCG keys): the DH secrets (the 5000x2 loop value is
    checked against the GPU and the host simulation, which are fast enough to run it)."""
    for c in CURVES:
        d = golden_rfc[c]["demo"]
        nb = golden_rfc[c]["nbytes"]
        g = PRIMES[c].generator.to_bytes(nb, "little")
        a, b = bytes.fromhex(d["alice"]), bytes.fromhex(d["bob"])
        apk, bpk = rfc7748(c, a, g), rfc7748(c, b, g)
        assert rfc7748(c, a, bpk).hex() == d["ssa"] == d["ssb"] == rfc7748(c, b, apk).hex()
    assert golden_rfc["X25519"]["demo"]["loop5000"] == "2ac5ee2022d2eed9b890983a064f3e0521f81dfec78b0c2b933f620b43b1d41c"
    assert golden_rfc["X25519"]["demo"]["ssa"] == "24ff9d34a8dfa8013dba79c34b64a4fd51f56f888b6490a9fb97f0b4500d9713"


@pytest.mark.parametrize("curve", CURVES)
def test_rfc7748_edge_and_random_rows(golden_rfc, curve):
    rows = golden_rfc[curve]["edge"] + golden_rfc[curve]["random"][:24]
    for r in rows:
        assert rfc7748(curve, bytes.fromhex(r["k"]), bytes.fromhex(r["u"])).hex() == r["out"], r


@pytest.mark.parametrize("curve", CURVES)
def test_rfc7748_validation_tail(golden_rfc, curve):
    """rfc7748.c:228-251 (driver built without TWIST_SECURE): zero for points off the curve."""
    rows = golden_rfc[curve]["validate"]
    assert 10 < sum(1 for r in rows if int(r["out"], 16) == 0) < len(rows) - 10
    for r in rows[:30] + rows[40:60]:
        assert rfc7748(curve, bytes.fromhex(r["k"]), bytes.fromhex(r["u"]), twist_secure=False).hex() == r["out"], r


@pytest.mark.parametrize("name", list(ALL_PRIMES))
def test_field_golden(golden_field, name):
    """Field-level vectors produced by the reference's generated 64-bit C (tests/golden/make_golden.py)."""
    g = golden_field[name]
    F = FieldOracle(name)
    nb = g["nbytes"]
    for op, res in g["ops"].items():
        for i, (ah, bh) in enumerate(zip(g["a"], g["b"])):
            a, b = int(ah, 16), int(bh, 16)
            v, st = util.oracle_field_op(F, op, a, b, g["mli_int"])
            assert v.to_bytes(nb, "big").hex() == res["out"][i], (name, op, i)
            assert st == res["status"][i], (name, op, i, "status")


def _matpow(M, e, mod):
    n = len(M)
    R = [[int(i == j) for j in range(n)] for i in range(n)]
    while e:
        if e & 1:
            R = [[sum(R[i][k] * M[k][j] for k in range(n)) % mod for j in range(n)] for i in range(n)]
        M = [[sum(M[i][k] * M[k][j] for k in range(n)) % mod for j in range(n)] for i in range(n)]
        e >>= 1
    return R


# low 24 bits printed by the reference's ./time (pseudo.py:1250,1318,1382), random.seed(42) operands
# (pseudo.py:1862-1866); regenerated in this container: oracle/_ref/build_*.log
TIME_CHECKSUMS = {
    "X25519": (0x116640, 0x675A88, 0xE70A06),
    "X448": (0xBCDDE4, 0xA8450D, 0x189F52),
    "NIST256": (0xA47501, 0x717A99, 0xE1E067),
}


@pytest.mark.parametrize("name", list(PRIMES))
def test_time_checksums(name):
    """The 10^8-multiply loops of time.c (pseudo.py:1235-1242,1306-1310,1371-1374) collapse to
    exponent arithmetic; the oracle's modmul/modsqr/modinv must land on the printed checksums."""
    F = FieldOracle(name)
    p = F.p
    random.seed(42)
    ra, rb, rs, ri = (random.randint(0, p - 1) for _ in range(4))

    def step(x, y):
        z = [a + b for a, b in zip(x, y)]
        y = [a + b for a, b in zip(z, x)]
        x = [a + b for a, b in zip(y, z)]
        z = [a + b for a, b in zip(x, y)]
        y = [a + b for a, b in zip(z, x)]
        return x, y, z

    # sanity of the exponent model against the oracle's own modmul on three iterations
    x, y = ra, rb
    for _ in range(3):
        z = F.modmul(x, y); y = F.modmul(z, x); x = F.modmul(y, z); z = F.modmul(x, y); y = F.modmul(z, x)
    ex, ey, ez = [1, 0], [0, 1], None
    for _ in range(3):
        ex, ey, ez = step(ex, ey)
    assert z == pow(ra, ez[0], p) * pow(rb, ez[1], p) % p

    ex, ey, _ = step([1, 0], [0, 1])
    Mk = _matpow([ex, ey], 100000 * 200 - 1, p - 1)
    _, _, z = step(Mk[0], Mk[1])
    zm = pow(ra, z[0] % (p - 1), p) * pow(rb, z[1] % (p - 1), p) % p
    zs = pow(rs, pow(2, 2 * 100000 * 500 - 1, p - 1), p)
    zi = F.modinv(ri)
    assert (F.redc(zm) & 0xFFFFFF, F.redc(zs) & 0xFFFFFF, F.redc(zi) & 0xFFFFFF) == TIME_CHECKSUMS[name]


@pytest.mark.parametrize("name", list(PRIMES))
def test_generator_selftest_identity(name):
    """pseudo.py:1762-1767,1783-1796: the 14-call sequence ends on ((x-y)(x+y))^-2."""
    F = FieldOracle(name)
    p = F.p
    rng = random.Random(5)
    for _ in range(20):
        x, y = rng.randrange(2 * p), rng.randrange(2 * p)
        ax, ay = F.nres(x), F.nres(y)
        t = F.modadd(ax, ay)
        z = F.modsub(ax, ay)
        ax = F.modmul(t, z)
        z = F.modsqr(ax)
        z = F.modinv(z)
        z = F.modsqrt(z)
        z = F.modsqr(z)
        z = F.modhaf(z)
        z = F.modadd(z, z)
        want = pow(((x - y) * (x + y)) ** 2 % p, -1, p) if (x - y) * (x + y) % p else 0
        assert F.redc(z) == want


def test_oracle_against_reference_build(ref_libs):
    """oracle == the reference's own C (oracle/_ref) on random raw inputs, ladder and field ops."""
    if not ref_libs:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py)")
    for c in CURVES:
        if c not in ref_libs:
            continue
        nb = PRIMES[c].nbytes
        k, u = util.random_bytes(11, 48, nb), util.random_bytes(12, 48, nb)
        out = util.ref_rfc7748_batch(ref_libs[c], k, u)
        for i in range(48):
            assert rfc7748(c, k[i].tobytes(), u[i].tobytes()) == out[i].tobytes()
    for name, key in (("X25519", "X25519_generic"), ("X448", "X448_generic"), ("NIST256", "NIST256"),
                      ("SECP256K1", "SECP256K1"), ("NIST256ORDER", "NIST256ORDER")):
        if key not in ref_libs:
            continue
        F = FieldOracle(name)
        nb = F.nbytes
        a, b = util.random_bytes(21, 64, nb), util.random_bytes(22, 64, nb)
        if name == "X448":
            a[:, 0] &= 0x7F      # keep the integers < 2p as modimp requires (pseudo.py:1142)
        for op in ("mul", "sqr", "inv", "sqrt", "add", "sub", "neg", "haf", "mli", "qr"):
            out, st = util.ref_field_batch(ref_libs[key], op, a, b if op in ("mul", "add", "sub") else None, 39081)
            for i in range(64):
                av, bv = int.from_bytes(a[i].tobytes(), "big"), int.from_bytes(b[i].tobytes(), "big")
                v, s = util.oracle_field_op(F, op, av, bv, 39081)
                assert v.to_bytes(nb, "big") == out[i].tobytes(), (name, op, i)
                assert s == st[i]


def test_addchain():
    for P in PRIMES.values():
        prog = addchain.find_chain(P.pe)
        s, m = addchain.cost(prog)
        assert s == P.pe.bit_length() - 1 and m <= 16
        for x in (2, 3, P.p - 2):
            assert addchain.evaluate(prog, x, P.p) == pow(x, P.pe, P.p)
    rng = random.Random(1)
    for _ in range(100):
        e = rng.getrandbits(rng.randrange(1, 200)) | 1
        assert addchain.evaluate(addchain.find_chain(e), 5, 2**127 - 1) == pow(5, e, 2**127 - 1)
    txt = addchain.to_reference_text(addchain.find_chain(PRIMES["X25519"].pe))
    assert txt.startswith("tmp t0") and "shift" in txt and "add z z x" in txt


# ---- the plain-C restatement (oracle/oracle.c) -----------------------------------------------------
def test_c_oracle_golden_and_python_oracle(golden_rfc, golden_field):
    """oracle.c against every golden vector and against the Python restatement on random inputs."""
    import c_oracle
    for c in CURVES:
        g = golden_rfc[c]
        nb = g["nbytes"]
        v = g["rfc"]
        gen = PRIMES[c].generator.to_bytes(nb, "little").hex()
        rows = g["edge"] + g["random"] + [{"k": v["sk1"], "u": gen, "out": v["pk1"]}, {"k": v["sk2"], "u": gen, "out": v["pk2"]},
                                          {"k": v["sk1"], "u": v["pk2"], "out": v["shared"]},
                                          {"k": g["demo"]["alice"], "u": gen, "out": None}]
        k = np.frombuffer(b"".join(bytes.fromhex(r["k"]) for r in rows), dtype=np.uint8).reshape(-1, nb)
        u = np.frombuffer(b"".join(bytes.fromhex(r["u"]) for r in rows), dtype=np.uint8).reshape(-1, nb)
        out = c_oracle.rfc7748_batch(c, k, u)
        for i, r in enumerate(rows):
            if r["out"] is not None:
                assert out[i].tobytes().hex() == r["out"], (c, i)
        k, u = util.random_bytes(51, 12, nb), util.random_bytes(52, 12, nb)
        out = c_oracle.rfc7748_batch(c, k, u)
        for i in range(12):
            assert out[i].tobytes() == rfc7748(c, k[i].tobytes(), u[i].tobytes())
    for name, g in golden_field.items():
        nb = g["nbytes"]
        a = np.frombuffer(bytes.fromhex("".join(g["a"])), dtype=np.uint8).reshape(-1, nb)
        b = np.frombuffer(bytes.fromhex("".join(g["b"])), dtype=np.uint8).reshape(-1, nb)
        for op, res in g["ops"].items():
            out, st = c_oracle.field_batch(name, op, a, b if op in ("mul", "add", "sub") else None, g["mli_int"])
            assert [out[i].tobytes().hex() for i in range(a.shape[0])] == res["out"], (name, op)
            assert list(st) == res["status"], (name, op)
        F = FieldOracle(name)
        a, b = util.random_bytes(61, 40, nb), util.random_bytes(62, 40, nb)
        for op in ("mul", "inv", "sqrt", "sub", "haf", "qr", "pro"):
            out, st = c_oracle.field_batch(name, op, a, b if op in ("mul", "sub") else None, 0)
            for i in range(40):
                av, bv = int.from_bytes(a[i].tobytes(), "big"), int.from_bytes(b[i].tobytes(), "big")
                v, s = util.oracle_field_op(F, op, av, bv, 0)
                assert v.to_bytes(nb, "big") == out[i].tobytes() and s == st[i], (name, op, i)


def test_c_oracle_against_reference_build(ref_libs):
    if "X25519" not in ref_libs:
        pytest.skip("oracle/_ref not built")
    import c_oracle
    for c in CURVES:
        nb = PRIMES[c].nbytes
        n = 512 if c == "X25519" else 128
        k, u = util.random_bytes(71, n, nb), util.random_bytes(72, n, nb)
        assert np.array_equal(c_oracle.rfc7748_batch(c, k, u), util.ref_rfc7748_batch(ref_libs[c], k, u))


def test_oracle_tables_are_its_own_and_agree_with_the_product():
    """The oracle carries its own moduli and curve constants (oracle/oracle_primes.py, restated from the reference's
    tables) and its own modpro; nothing under oracle/ imports the product package.  The two tables must agree."""
    import re
    import oracle_primes
    from modarith_b200.primes import ALL_PRIMES
    assert set(oracle_primes.TABLE) == set(ALL_PRIMES)
    for name, P in ALL_PRIMES.items():
        Q = oracle_primes.TABLE[name]
        for key in ("p", "nbits", "nbytes", "pm1d2", "pe", "roi", "a24", "cof", "generator", "ed_d", "ed_gx", "ed_gy",
                    "ed_order", "wb", "wgx", "wgy", "worder"):
            assert getattr(Q, key) == getattr(P, key), (name, key)
    root = os.path.join(os.path.dirname(__file__), "..", "oracle")
    for fn in os.listdir(root):
        # addchain_standin.py is not part of the checker: it feeds the REFERENCE generators an addition chain when
        # oracle/_ref is built (the order of squarings / multiplications inside the reference's modpro, never a value)
        # and deliberately hands them the chain the product uses, so that the CPU baseline is not slowed by a worse one
        if fn.endswith(".py") and fn != "addchain_standin.py":
            src = open(os.path.join(root, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+modarith_b200", src, re.M), fn
    # modpro by the oracle's own square-and-multiply equals the plain power for every modulus
    for name in ALL_PRIMES:
        O = FieldOracle(name)
        for w in (0, 1, 2, 3, O.p - 1, 0x123456789ABCDEF % O.p):
            assert O.modpro(w) == pow(w, O.pe, O.p)
