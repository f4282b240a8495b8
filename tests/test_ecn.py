"""Short-Weierstrass scalar multiplication: oracle pinned against vectors produced
by the reference's own weierstrass.c (via its curve.py), device logic on the host simulation, CUDA build
through the C ABI."""
import ctypes

import numpy as np
import pytest

from field_oracle import ecnmul as oracle_ecnmul, ecnmul_edwards as oracle_ecnmul_edwards
from modarith_b200.primes import ALL_PRIMES
import util

P = ALL_PRIMES["NIST256"]


def test_oracle_against_reference_vectors(golden_ecn):
    rows = golden_ecn["NIST256"]
    assert sum(1 for r in rows if int(r["xo"], 16) == 0 and int(r["yo"], 16) == 1) >= 5      # infinities present
    for r in rows:
        xo, yo = oracle_ecnmul("NIST256", bytes.fromhex(r["e"]), bytes.fromhex(r["x"]), bytes.fromhex(r["y"]))
        assert (xo.hex(), yo.hex()) == (r["xo"], r["yo"]), r
    # known answer: 2G (SEC 2 / NIST test vectors)
    g = (P.wgx.to_bytes(32, "big"), P.wgy.to_bytes(32, "big"))
    assert oracle_ecnmul("NIST256", (2).to_bytes(32, "big"), *g)[0].hex() == \
        "7cf27b188d034f7e8a52380304b51ac3c08969e277f21b35a60b48fc47669978"


def test_edwards_oracle_against_reference_vectors(golden_ecn):
    rows = golden_ecn["ED25519"]
    assert sum(1 for r in rows if int(r["xo"], 16) == 0 and int(r["yo"], 16) == 1) >= 5
    for r in rows:
        xo, yo = oracle_ecnmul_edwards("X25519", bytes.fromhex(r["e"]), bytes.fromhex(r["x"]), bytes.fromhex(r["y"]))
        assert (xo.hex(), yo.hex()) == (r["xo"], r["yo"]), r


@pytest.mark.parametrize("curve", ["NIST256", "ED25519"])
def test_hostsim_against_reference_vectors(hostsim, golden_ecn, curve):
    fn = getattr(hostsim, "sim_%s_ecnmul" % curve)
    for r in golden_ecn[curve]:
        xo, yo = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        fn(bytes.fromhex(r["e"]), bytes.fromhex(r["x"]), bytes.fromhex(r["y"]), xo, yo)
        assert (xo.raw[:32].hex(), yo.raw[:32].hex()) == (r["xo"], r["yo"]), r


def _gpu(e, x, y, curve="NIST256"):
    import torch
    from modarith_b200.ecn import ecnmul
    xo, yo = ecnmul(curve, torch.from_numpy(e).cuda(), torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())
    torch.cuda.synchronize()
    return xo.cpu().numpy(), yo.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("curve", ["NIST256", "ED25519"])
def test_gpu_against_reference_vectors(golden_ecn, curve):
    rows = golden_ecn[curve]
    f = lambda k: np.frombuffer(b"".join(bytes.fromhex(r[k]) for r in rows), dtype=np.uint8).reshape(-1, 32).copy()
    xo, yo = _gpu(f("e"), f("x"), f("y"), curve)
    for i, r in enumerate(rows):
        assert (xo[i].tobytes().hex(), yo[i].tobytes().hex()) == (r["xo"], r["yo"]), (i, r)


@pytest.mark.gpu
def test_gpu_ecdh_property_and_reference_build(ref_libs):
    """Diffie-Hellman on 4096 random pairs: a*(b*G) == b*(a*G); every element against the reference build."""
    n = 4096
    a, b = util.random_bytes(601, n, 32), util.random_bytes(602, n, 32)
    gx = np.tile(np.frombuffer(P.wgx.to_bytes(32, "big"), dtype=np.uint8), (n, 1))
    gy = np.tile(np.frombuffer(P.wgy.to_bytes(32, "big"), dtype=np.uint8), (n, 1))
    ax, ay = _gpu(a, gx, gy)
    bx, by = _gpu(b, gx, gy)
    s1 = _gpu(a, bx, by)
    s2 = _gpu(b, ax, ay)
    assert np.array_equal(s1[0], s2[0]) and np.array_equal(s1[1], s2[1])
    for i in range(0, n, 512):
        assert (ax[i].tobytes(), ay[i].tobytes()) == oracle_ecnmul("NIST256", a[i].tobytes(), gx[i].tobytes(), gy[i].tobytes())
    if "NIST256_curve" in ref_libs:
        lib = ref_libs["NIST256_curve"]
        xo, yo = np.zeros_like(ax), np.zeros_like(ay)
        lib.ref_ecnmul_batch(a.ctypes.data_as(ctypes.c_char_p), bx.ctypes.data_as(ctypes.c_char_p),
                             by.ctypes.data_as(ctypes.c_char_p), xo.ctypes.data_as(ctypes.c_char_p),
                             yo.ctypes.data_as(ctypes.c_char_p), ctypes.c_size_t(n), ctypes.c_int(0))
        assert np.array_equal(xo, s1[0]) and np.array_equal(yo, s1[1])


@pytest.mark.gpu
def test_gpu_ed25519_property_and_reference_build(ref_libs):
    """Ed25519: a*(b*G) == b*(a*G) on 4096 random pairs; every element against the reference's edwards.c build."""
    Q = ALL_PRIMES["X25519"]
    n = 4096
    a, b = util.random_bytes(701, n, 32), util.random_bytes(702, n, 32)
    gx = np.tile(np.frombuffer(Q.ed_gx.to_bytes(32, "big"), dtype=np.uint8), (n, 1))
    gy = np.tile(np.frombuffer(Q.ed_gy.to_bytes(32, "big"), dtype=np.uint8), (n, 1))
    ax, ay = _gpu(a, gx, gy, "ED25519")
    bx, by = _gpu(b, gx, gy, "ED25519")
    s1 = _gpu(a, bx, by, "ED25519")
    s2 = _gpu(b, ax, ay, "ED25519")
    assert np.array_equal(s1[0], s2[0]) and np.array_equal(s1[1], s2[1])
    for i in range(0, n, 512):
        assert (ax[i].tobytes(), ay[i].tobytes()) == oracle_ecnmul_edwards("X25519", a[i].tobytes(), gx[i].tobytes(), gy[i].tobytes())
    if "ED25519_curve" in ref_libs:
        lib = ref_libs["ED25519_curve"]
        xo, yo = np.zeros_like(ax), np.zeros_like(ay)
        lib.ref_ecnmul_batch(a.ctypes.data_as(ctypes.c_char_p), bx.ctypes.data_as(ctypes.c_char_p),
                             by.ctypes.data_as(ctypes.c_char_p), xo.ctypes.data_as(ctypes.c_char_p),
                             yo.ctypes.data_as(ctypes.c_char_p), ctypes.c_size_t(n), ctypes.c_int(0))
        assert np.array_equal(xo, s1[0]) and np.array_equal(yo, s1[1])


@pytest.mark.gpu
@pytest.mark.parametrize("curve", ["NIST256", "ED25519"])
def test_gpu_large_ragged_batch_is_position_independent(curve):
    """More 128-point blocks than the persistent grid has CTAs (each CTA walks several blocks and re-uses
    its table slice), a ragged last block, and an empty batch: every row must equal the row of the small
    batch it was copied from, which the other tests pin to the reference."""
    Q = ALL_PRIMES["X25519"]
    gxv, gyv = (P.wgx, P.wgy) if curve == "NIST256" else (Q.ed_gx, Q.ed_gy)
    base = 1000
    e0 = util.random_bytes(811, base, 32)
    e0[0] = 0                                              # zero scalar -> (0, 1)
    gx = np.tile(np.frombuffer(gxv.to_bytes(32, "big"), dtype=np.uint8), (base, 1))
    gy = np.tile(np.frombuffer(gyv.to_bytes(32, "big"), dtype=np.uint8), (base, 1))
    gy[1, 31] ^= 1                                         # a point that is not on the curve -> (0, 1)
    x0, y0 = _gpu(e0, gx, gy, curve)
    assert x0[0].tobytes() == bytes(32) and y0[0].tobytes() == (1).to_bytes(32, "big")
    assert x0[1].tobytes() == bytes(32) and y0[1].tobytes() == (1).to_bytes(32, "big")
    ora = oracle_ecnmul if curve == "NIST256" else oracle_ecnmul_edwards
    name = "NIST256" if curve == "NIST256" else "X25519"
    for i in (2, 499, 999):
        assert (x0[i].tobytes(), y0[i].tobytes()) == ora(name, e0[i].tobytes(), gx[i].tobytes(), gy[i].tobytes())
    n = 230 * base + 77                                    # 1798 blocks of 128 for a grid of a few hundred CTAs
    idx = (np.arange(n) * 7919) % base
    xo, yo = _gpu(e0[idx], gx[idx], gy[idx], curve)
    assert np.array_equal(xo, x0[idx]) and np.array_equal(yo, y0[idx])
    xe, ye = _gpu(e0[:0], gx[:0], gy[:0], curve)
    assert xe.shape == (0, 32) and ye.shape == (0, 32)


def test_hostsim_ed25519_random_points_check_operand_bounds(hostsim):
    """Random multiples of the generator (so that the table and the accumulator see arbitrary coordinates) with
    random and saturated scalars, single and double multiplication: the host simulation aborts if an add_tt /
    sub_tt of the Edwards law ever receives an operand above the product bound; results against the oracle."""
    from field_oracle import ecnmul2 as oracle_ecnmul2
    Q = ALL_PRIMES["X25519"]
    g = (Q.ed_gx.to_bytes(32, "big"), Q.ed_gy.to_bytes(32, "big"))
    rng = np.random.default_rng(25519)
    sc = [rng.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(10)] + [b"\xff" * 32, b"\x88" * 32, b"\x77" * 32,
                                                                                   b"\x80" + bytes(31), bytes(31) + b"\x08"]
    f1, f2 = hostsim.sim_ED25519_ecnmul, hostsim.sim_ED25519_ecnmul2
    pts = [g]
    for e in sc:
        x, y = pts[-1]
        xo, yo = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        f1(e, x, y, xo, yo)
        assert (xo.raw[:32], yo.raw[:32]) == oracle_ecnmul_edwards("X25519", e, x, y)
        if xo.raw[:32] != bytes(32):
            pts.append((xo.raw[:32], yo.raw[:32]))
    for i in range(len(sc) - 1):
        (x1, y1), (x2, y2) = pts[i % len(pts)], pts[(i + 3) % len(pts)]
        xo, yo = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        f2(sc[i], x1, y1, sc[i + 1], x2, y2, xo, yo)
        assert (xo.raw[:32], yo.raw[:32]) == oracle_ecnmul2("X25519", sc[i], x1, y1, sc[i + 1], x2, y2)
