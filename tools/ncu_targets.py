#!/usr/bin/env python3
"""Launch the kernels the round-2 work is about, twice each, so that ncu can capture the second launch:

    ncu --set full --clock-control none --import-source on -k regex:'k_rfc7748_rounds|k_field' ... \
        python tools/ncu_targets.py [x25519] [x448] [p256] [k1] [order] [ecn]
"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from modarith_b200 import Field  # noqa: E402
from modarith_b200.rfc7748 import rfc7748  # noqa: E402

what = set(sys.argv[1:]) or {"x25519", "p256"}
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
for curve, nb, n in (("X25519", 32, 1 << 20), ("X448", 56, 1 << 19)):
    if curve.lower() not in what:
        continue
    k = torch.randint(0, 256, (n, nb), dtype=torch.uint8, device=dev, generator=g)
    u = torch.randint(0, 256, (n, nb), dtype=torch.uint8, device=dev, generator=g)
    for _ in range(2):
        rfc7748(curve, k, u)
    torch.cuda.synchronize()
for key, name in (("p256", "NIST256"), ("k1", "SECP256K1"), ("order", "NIST256ORDER"), ("f25519", "X25519")):
    if key not in what:
        continue
    F = Field(name)
    n = 1 << 20
    a = torch.randint(0, 256, (n, F.Nbytes), dtype=torch.uint8, device=dev, generator=g)
    b = torch.randint(0, 256, (n, F.Nbytes), dtype=torch.uint8, device=dev, generator=g)
    x, _ = F.modimp(a)
    y, _ = F.modimp(b)
    r = F.alloc(n)
    for _ in range(2):
        F.bench_modmul(x, y, r, 256)        # k_field<F, 31>
    for _ in range(2):
        F.modnsqr(r, 256)                   # k_field<F, 7>
    for _ in range(2):
        F.modinv_perelement(x, r)           # k_field<F, 9>
    for _ in range(2):
        F.modsqrt(x, None, r)               # k_field<F, 13>
    torch.cuda.synchronize()
if "ecn" in what:
    from modarith_b200.ecn import ecnmul, ecnmul2
    from modarith_b200.primes import PRIMES, X25519
    n = 1 << 18
    e = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    f = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    for curve, gx, gy in (("NIST256", PRIMES["NIST256"].wgx, PRIMES["NIST256"].wgy), ("ED25519", X25519.ed_gx, X25519.ed_gy)):
        x = torch.from_numpy(np.tile(np.frombuffer(gx.to_bytes(32, "big"), dtype=np.uint8), (n, 1))).to(dev)
        y = torch.from_numpy(np.tile(np.frombuffer(gy.to_bytes(32, "big"), dtype=np.uint8), (n, 1))).to(dev)
        for _ in range(2):
            ecnmul(curve, e, x, y)
        h = n // 2
        for _ in range(2):
            ecnmul2(curve, e[:h], x[:h], y[:h], f[:h], x[:h], y[:h])
    torch.cuda.synchronize()
if "jit" in what:
    # compiled field programs (mab_<P>_modprog_jit): the P-256 point addition, the interpreter's kernel beside it
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from bench_jit import POINT_ADD
    F = Field("NIST256")
    n = 1 << 21
    ops = [F.modimp(torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g))[0] for _ in range(7)]
    outs = [F.alloc(n) for _ in range(3)]
    for jit in (True, False):
        for _ in range(2):
            F.modprog(POINT_ADD, ops, [12, 13, 14], outputs=outs, jit=jit)     # k_prog_jit / k_prog<F_NIST256>
    torch.cuda.synchronize()
if "addon" in what:
    # add-on moduli on the bit-level pseudo-Mersenne plan and the user-curve ladder
    for name in ("C41417", "NIST521"):
        F = Field(name)
        n = 1 << 20
        x, _ = F.modimp(torch.randint(0, 128, (n, F.Nbytes), dtype=torch.uint8, device=dev, generator=g))
        y, _ = F.modimp(torch.randint(0, 128, (n, F.Nbytes), dtype=torch.uint8, device=dev, generator=g))
        r = F.alloc(n)
        for _ in range(2):
            F.bench_modmul(x, y, r, 128)
    k = torch.randint(0, 256, (1 << 19, 48), dtype=torch.uint8, device=dev, generator=g)
    u = torch.randint(0, 256, (1 << 19, 48), dtype=torch.uint8, device=dev, generator=g)
    for _ in range(2):
        rfc7748("M383", k, u)
    torch.cuda.synchronize()
print("done")
