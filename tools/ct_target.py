#!/usr/bin/env python3
"""One launch of each secret-handling kernel on inputs of a chosen pattern, for the dynamic half of the
constant-time check (tests/test_gpu_ct.py runs this under `ncu --metrics smsp__inst_executed.sum`):

    python tools/ct_target.py zero|ones|random|lowbits

zero / ones: every key and every point byte 0x00 / 0xff; random: PCG64(1); lowbits: only the lowest bit of every
32-bit word set.  The batch shapes are the same for every pattern, so any difference in executed instructions is a
data-dependent path."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from modarith_b200 import Field, lib as mlib  # noqa: E402
from modarith_b200.ecn import ecnmul  # noqa: E402
from modarith_b200.primes import NIST256, X25519  # noqa: E402


def pattern(kind, n, nb, seed):
    if kind == "zero":
        return np.zeros((n, nb), dtype=np.uint8)
    if kind == "ones":
        return np.full((n, nb), 0xFF, dtype=np.uint8)
    if kind == "lowbits":
        a = np.zeros((n, nb), dtype=np.uint8)
        a[:, ::4] = 1
        return a
    return np.random.Generator(np.random.PCG64(seed)).integers(0, 256, (n, nb), dtype=np.uint8)


def main(kind):
    l = mlib.load()
    st = torch.cuda.current_stream().cuda_stream
    n = 4096
    for curve, nb in (("X25519", 32), ("X448", 56)):
        k = torch.from_numpy(pattern(kind, n, nb, 1)).cuda()
        u = torch.from_numpy(pattern(kind, n, nb, 2)).cuda()
        v = torch.empty_like(k)
        mlib.check(getattr(l, "mab_%s_rfc7748_perkey" % curve)(k.data_ptr(), u.data_ptr(), v.data_ptr(), n, st))
        mlib.check(getattr(l, "mab_%s_rfc7748" % curve)(k.data_ptr(), u.data_ptr(), v.data_ptr(), n, st))
    # scalar multiplications: the scalar is the secret; the point stays the generator so that it is on the curve
    for curve, gx, gy in (("NIST256", NIST256.wgx, NIST256.wgy), ("ED25519", X25519.ed_gx, X25519.ed_gy)):
        e = torch.from_numpy(pattern(kind, 2048, 32, 3)).cuda()
        x = torch.from_numpy(np.tile(np.frombuffer(gx.to_bytes(32, "big"), dtype=np.uint8), (2048, 1))).cuda()
        y = torch.from_numpy(np.tile(np.frombuffer(gy.to_bytes(32, "big"), dtype=np.uint8), (2048, 1))).cuda()
        ecnmul(curve, e, x, y)
    # field: inversion and square root of secret elements (shared-chain and per-element kernels)
    F = Field("NIST256")
    a, _ = F.modimp(torch.from_numpy(pattern(kind, 4096, 32, 4)).cuda())
    r = F.alloc(4096)
    F.modinv(a, None, r)
    F.modinv_perelement(a, r)
    F.modsqrt(a, None, r)
    torch.cuda.synchronize()
    print("done", kind)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "random")
