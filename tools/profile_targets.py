#!/usr/bin/env python3
"""Launch each hot kernel a few times so `ncu -k regex:... ` can capture it (used under gpurun).

    ncu --set full --clock-control none --import-source on -k regex:'k_rfc7748|k_field' -c 12 \
        -o gpurun_out/prof python tools/profile_targets.py
"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from modarith_b200 import Field  # noqa: E402
from modarith_b200.rfc7748 import rfc7748  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
for curve, nb, n in (("X25519", 32, 1 << 20), ("X448", 56, 1 << 19)):
    k = torch.randint(0, 256, (n, nb), dtype=torch.uint8, device=dev, generator=g)
    u = torch.randint(0, 256, (n, nb), dtype=torch.uint8, device=dev, generator=g)
    for _ in range(2):
        rfc7748(curve, k, u)
torch.cuda.synchronize()
if "--no-ecn" not in sys.argv:
    import numpy as np
    from modarith_b200.ecn import ecnmul
    from modarith_b200.primes import PRIMES, X25519
    n = 1 << 18
    e = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    for curve, gx, gy in (("NIST256", PRIMES["NIST256"].wgx, PRIMES["NIST256"].wgy), ("ED25519", X25519.ed_gx, X25519.ed_gy)):
        x = torch.from_numpy(np.tile(np.frombuffer(gx.to_bytes(32, "big"), dtype=np.uint8), (n, 1))).to(dev)
        y = torch.from_numpy(np.tile(np.frombuffer(gy.to_bytes(32, "big"), dtype=np.uint8), (n, 1))).to(dev)
        for _ in range(2):
            ecnmul(curve, e, x, y)
    torch.cuda.synchronize()
if "--ladders-only" in sys.argv:
    print("done")
    sys.exit(0)
for name, n in (("NIST256", 1 << 22), ("X25519", 1 << 22)):
    F = Field(name)
    a = torch.randint(0, 256, (n, F.Nbytes), dtype=torch.uint8, device=dev, generator=g)
    b = torch.randint(0, 256, (n, F.Nbytes), dtype=torch.uint8, device=dev, generator=g)
    x, _ = F.modimp(a)
    y, _ = F.modimp(b)
    r = F.alloc(n)
    F.modmul(x, y, r)
    F.modadd(x, y, r)
    m = 1 << 20
    xs, ys, rs = x[:, :m].contiguous(), y[:, :m].contiguous(), r[:, :m].contiguous()
    F.bench_modmul(xs, ys, rs, 256)
    F.modnsqr(rs, 256)
    F.modinv(xs, None, rs)
    F.modsqrt(xs, None, rs)
torch.cuda.synchronize()
print("done")
