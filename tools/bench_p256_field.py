#!/usr/bin/env python3
"""P-256 field chains and scalar multiplications of ONE library build (MODARITH_B200_LIB selects a variant built by
tools/variants.py): register-resident modmul / modnsqr chains, modinv per element, modsqrt, ecnmul, ecnmul2, each
with a hash of its output so that variants can be shown to agree.

    python tools/bench_p256_field.py [tag]
"""
import hashlib
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from modarith_b200 import Field  # noqa: E402
from modarith_b200.ecn import ecnmul, ecnmul2  # noqa: E402
from modarith_b200.primes import NIST256 as P256  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "shipped"
dev = torch.device("cuda", 0)
PEAK = 148 * 4 * 8 * 1.965e9          # IMAD.WIDE lanes per clock x SM clock: 9.31 Tprod/s nominal


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def h(t):
    return hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:12]


gen = torch.Generator(device=dev).manual_seed(256)
F = Field("NIST256", dev)
m, iters = 1 << 21, 512
x, _ = F.modimp(torch.randint(0, 256, (m, 32), dtype=torch.uint8, device=dev, generator=gen))
y, _ = F.modimp(torch.randint(0, 256, (m, 32), dtype=torch.uint8, device=dev, generator=gen))
r = F.alloc(m)
t = timed(lambda: F.bench_modmul(x, y, r, iters))
print("%-10s modmul chain   %8.2f Gop/s  %.3f of nominal IMAD peak  %s" % (tag, m * iters / t / 1e9, m * iters / t * 64 / PEAK, h(r)), flush=True)


def nsq():
    r.copy_(x)
    F.modnsqr(r, iters)


t0 = timed(lambda: r.copy_(x))
t = timed(nsq) - t0
print("%-10s modnsqr chain  %8.2f Gop/s  %.3f  %s" % (tag, m * iters / t / 1e9, m * iters / t * 36 / PEAK, h(r)), flush=True)
t = timed(lambda: F.modinv_perelement(x, r))
print("%-10s modinv/element %8.2f Mop/s  %s" % (tag, m / t / 1e6, h(r)), flush=True)
t = timed(lambda: F.modsqrt(x, None, r))
print("%-10s modsqrt        %8.2f Mop/s  %s" % (tag, m / t / 1e6, h(r)), flush=True)
ne = 1 << 18
e = torch.randint(0, 256, (ne, 32), dtype=torch.uint8, device=dev, generator=gen)
f2 = torch.randint(0, 256, (ne, 32), dtype=torch.uint8, device=dev, generator=gen)
gx = torch.from_numpy(np.tile(np.frombuffer(P256.wgx.to_bytes(32, "big"), dtype=np.uint8), (ne, 1))).to(dev)
gy = torch.from_numpy(np.tile(np.frombuffer(P256.wgy.to_bytes(32, "big"), dtype=np.uint8), (ne, 1))).to(dev)
out = [None]


def em():
    out[0] = ecnmul("NIST256", e, gx, gy)


t = timed(em, 2)
print("%-10s ecnmul         %8.2f M/s  %s" % (tag, ne / t / 1e6, h(out[0][0])), flush=True)
n2 = ne // 2


def em2():
    out[0] = ecnmul2("NIST256", e[:n2], gx[:n2], gy[:n2], f2[:n2], gx[:n2], gy[:n2])


t = timed(em2, 2)
print("%-10s ecnmul2        %8.2f M/s  %s" % (tag, n2 / t / 1e6, h(out[0][0])), flush=True)
