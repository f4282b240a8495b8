"""Double scalar multiplication e*P + f*Q (ecnXXXmul2, weierstrass.c:545-572 / edwards.c:486-513; the
verification building block of SURVEY.md 8f row 1): oracle pinned against the reference's own routine,
device logic on the host simulation, CUDA build through the C ABI."""
import ctypes

import numpy as np
import pytest

from field_oracle import ecnmul as oracle_ecnmul, ecnmul_edwards as oracle_ecnmul_edwards, ecnmul2 as oracle_ecnmul2
from modarith_b200.primes import ALL_PRIMES
import util

CURVES = [("NIST256", "NIST256"), ("ED25519", "X25519")]
KEYS = ("e", "x1", "y1", "f", "x2", "y2")
ONE = (1).to_bytes(32, "big")


def _gen(curve):
    Q = ALL_PRIMES["NIST256" if curve == "NIST256" else "X25519"]
    return (Q.wgx, Q.wgy) if curve == "NIST256" else (Q.ed_gx, Q.ed_gy)


@pytest.mark.parametrize("curve,prime", CURVES)
def test_oracle_against_reference_vectors(golden_ecn2, curve, prime):
    for r in golden_ecn2[curve]:
        xo, yo = oracle_ecnmul2(prime, *(bytes.fromhex(r[k]) for k in KEYS))
        assert (xo.hex(), yo.hex()) == (r["xo"], r["yo"]), r
    # the case the reference cannot run (it reads before its digit array): both scalars zero -> identity
    r = golden_ecn2[curve][0]
    z = bytes(32)
    assert oracle_ecnmul2(prime, z, bytes.fromhex(r["x1"]), bytes.fromhex(r["y1"]), z, bytes.fromhex(r["x2"]),
                          bytes.fromhex(r["y2"])) == (z, ONE)


@pytest.mark.parametrize("curve,prime", CURVES)
def test_hostsim_against_reference_vectors(hostsim, golden_ecn2, curve, prime):
    fn = getattr(hostsim, "sim_%s_ecnmul2" % curve)
    rows = list(golden_ecn2[curve])
    rows.append(dict(rows[0], e="00" * 32, f="00" * 32, xo="00" * 32, yo=ONE.hex()))
    for r in rows:
        xo, yo = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        fn(*(bytes.fromhex(r[k]) for k in KEYS), xo, yo)
        assert (xo.raw[:32].hex(), yo.raw[:32].hex()) == (r["xo"], r["yo"]), r


def _gpu(curve, *arrays):
    import torch
    from modarith_b200.ecn import ecnmul2
    xo, yo = ecnmul2(curve, *(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays))
    torch.cuda.synchronize()
    return xo.cpu().numpy(), yo.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("curve,prime", CURVES)
def test_gpu_against_reference_vectors(golden_ecn2, curve, prime):
    rows = list(golden_ecn2[curve])
    rows.append(dict(rows[0], e="00" * 32, f="00" * 32, xo="00" * 32, yo=ONE.hex()))
    cols = [np.frombuffer(b"".join(bytes.fromhex(r[k]) for r in rows), dtype=np.uint8).reshape(-1, 32).copy() for k in KEYS]
    xo, yo = _gpu(curve, *cols)
    for i, r in enumerate(rows):
        assert (xo[i].tobytes().hex(), yo[i].tobytes().hex()) == (r["xo"], r["yo"]), (i, r)


@pytest.mark.gpu
@pytest.mark.parametrize("curve,prime", CURVES)
def test_gpu_random_pairs_against_single_multiplications_and_reference_build(ref_libs, curve, prime):
    """3000 random pairs (ragged): e*P + f*Q with P = a*G, Q = b*G must equal (e*a + f*b)*G computed by the
    single multiplication (whose parity is pinned elsewhere); every row also against the reference's mul2."""
    import torch
    from modarith_b200.ecn import ecnmul
    n = 3000 + 37
    gxv, gyv = _gen(curve)
    order = ALL_PRIMES["NIST256"].worder if curve == "NIST256" else ALL_PRIMES["X25519"].ed_order
    gx = np.tile(np.frombuffer(gxv.to_bytes(32, "big"), dtype=np.uint8), (n, 1))
    gy = np.tile(np.frombuffer(gyv.to_bytes(32, "big"), dtype=np.uint8), (n, 1))
    a, b, e, f = (util.random_bytes(900 + i, n, 32) for i in range(4))
    dev = lambda v: torch.from_numpy(v).cuda()
    px, py = (t.cpu().numpy() for t in ecnmul(curve, dev(a), dev(gx), dev(gy)))
    qx, qy = (t.cpu().numpy() for t in ecnmul(curve, dev(b), dev(gx), dev(gy)))
    xo, yo = _gpu(curve, e, px, py, f, qx, qy)
    iv = lambda row: int.from_bytes(row.tobytes(), "big")
    k = np.frombuffer(b"".join(((iv(e[i]) * iv(a[i]) + iv(f[i]) * iv(b[i])) % order).to_bytes(32, "big") for i in range(n)),
                      dtype=np.uint8).reshape(n, 32).copy()
    sx, sy = (t.cpu().numpy() for t in ecnmul(curve, dev(k), dev(gx), dev(gy)))
    assert np.array_equal(xo, sx) and np.array_equal(yo, sy)
    for i in range(0, n, 1000):
        assert (xo[i].tobytes(), yo[i].tobytes()) == oracle_ecnmul2(prime, e[i].tobytes(), px[i].tobytes(), py[i].tobytes(),
                                                                    f[i].tobytes(), qx[i].tobytes(), qy[i].tobytes())
    key = curve + "_curve"
    if key in ref_libs and hasattr(ref_libs[key], "ref_ecnmul2_batch"):
        rx, ry = np.zeros_like(xo), np.zeros_like(yo)
        c = lambda v: v.ctypes.data_as(ctypes.c_char_p)
        ref_libs[key].ref_ecnmul2_batch(c(e), c(px), c(py), c(f), c(qx), c(qy), c(rx), c(ry), ctypes.c_size_t(n), ctypes.c_int(0))
        assert np.array_equal(rx, xo) and np.array_equal(ry, yo)


@pytest.mark.gpu
@pytest.mark.parametrize("curve,prime", CURVES)
def test_gpu_large_batch_is_position_independent(curve, prime):
    gxv, gyv = _gen(curve)
    base = 500
    gx = np.tile(np.frombuffer(gxv.to_bytes(32, "big"), dtype=np.uint8), (base, 1))
    gy = np.tile(np.frombuffer(gyv.to_bytes(32, "big"), dtype=np.uint8), (base, 1))
    e, f = util.random_bytes(931, base, 32), util.random_bytes(932, base, 32)
    e[0] = 0
    f[0] = 0
    x0, y0 = _gpu(curve, e, gx, gy, f, gx, gy)
    assert x0[0].tobytes() == bytes(32) and y0[0].tobytes() == ONE
    n = 150 * base + 19
    idx = (np.arange(n) * 7919) % base
    xo, yo = _gpu(curve, e[idx], gx[idx], gy[idx], f[idx], gx[idx], gy[idx])
    assert np.array_equal(xo, x0[idx]) and np.array_equal(yo, y0[idx])
    xe, ye = _gpu(curve, e[:0], gx[:0], gy[:0], f[:0], gx[:0], gy[:0])
    assert xe.shape == (0, 32)


# ---- the reference's own group test, testcurve.c main (generator, order*G = O, r1*G + r2*G = O, the two
# ---- iterated loops P = n1*P and P = n1*P + n2*G run 150 times each by the reference's own functions) ----
def _chain(mul, mul2, t):
    be = lambda h: int(h, 16).to_bytes(32, "big")
    g = (bytes.fromhex(t["gx"]), bytes.fromhex(t["gy"]))
    assert mul(be(t["order"]), *g) == (bytes(32), ONE)                                   # "MUL test"
    assert mul2(be(t["r1"]), *g, be(t["r2"]), *g) == (bytes(32), ONE)                    # "MUL2 test"
    a, b = be(t["n1"]), be(t["n2"])
    P = g
    for _ in range(t["iters"]):
        P = mul(a, *P)
    assert (P[0].hex(), P[1].hex()) == (t["x1"], t["y1"])
    for _ in range(t["iters"]):
        P = mul2(a, *P, b, *g)
    assert (P[0].hex(), P[1].hex()) == (t["x2"], t["y2"])


@pytest.mark.parametrize("curve,prime", CURVES)
def test_testcurve_flow_oracle(golden_testcurve, curve, prime):
    one = oracle_ecnmul if curve == "NIST256" else oracle_ecnmul_edwards
    _chain(lambda e, x, y: one(prime, e, x, y), lambda e, x1, y1, f, x2, y2: oracle_ecnmul2(prime, e, x1, y1, f, x2, y2),
           golden_testcurve[curve])


@pytest.mark.parametrize("curve,prime", CURVES)
def test_testcurve_flow_hostsim(hostsim, golden_testcurve, curve, prime):
    f1, f2 = getattr(hostsim, "sim_%s_ecnmul" % curve), getattr(hostsim, "sim_%s_ecnmul2" % curve)

    def mul(e, x, y):
        xo, yo = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        f1(e, x, y, xo, yo)
        return xo.raw[:32], yo.raw[:32]

    def mul2(e, x1, y1, f, x2, y2):
        xo, yo = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        f2(e, x1, y1, f, x2, y2, xo, yo)
        return xo.raw[:32], yo.raw[:32]

    _chain(mul, mul2, golden_testcurve[curve])


@pytest.mark.gpu
@pytest.mark.parametrize("curve,prime", CURVES)
def test_testcurve_flow_gpu(golden_testcurve, curve, prime):
    import torch
    from modarith_b200.ecn import ecnmul, ecnmul2
    row = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).reshape(1, 32).cuda()
    back = lambda t: t.cpu().numpy().tobytes()

    def mul(e, x, y):
        xo, yo = ecnmul(curve, row(e), row(x), row(y))
        return back(xo), back(yo)

    def mul2(e, x1, y1, f, x2, y2):
        xo, yo = ecnmul2(curve, row(e), row(x1), row(y1), row(f), row(x2), row(y2))
        return back(xo), back(yo)

    _chain(mul, mul2, golden_testcurve[curve])


def test_argument_checks_need_no_gpu():
    from modarith_b200 import ecn
    a = np.zeros((3, 32), dtype=np.uint8)
    with pytest.raises(ValueError):
        ecn.ecnmul("SECP256K1", a, a, a)
    import torch
    if not torch.cuda.is_available():
        from modarith_b200.lib import MabError
        with pytest.raises(MabError):
            ecn.ecnmul("NIST256", a, a, a)                 # no CPU fallback


@pytest.mark.gpu
def test_gpu_host_arrays_round_trip(golden_ecn2):
    from modarith_b200 import ecn
    rows = golden_ecn2["NIST256"][:8]
    cols = [np.frombuffer(b"".join(bytes.fromhex(r[k]) for r in rows), dtype=np.uint8).reshape(-1, 32).copy() for k in KEYS]
    xo, yo = ecn.ecnmul2("NIST256", *cols)
    assert isinstance(xo, np.ndarray)
    assert [xo[i].tobytes().hex() for i in range(8)] == [r["xo"] for r in rows]
    with pytest.raises(ValueError):
        ecn.ecnmul2("NIST256", cols[0][:, :31], *cols[1:])
    with pytest.raises(ValueError):
        ecn.ecnmul("NIST256", cols[0], cols[1][:4], cols[2])
