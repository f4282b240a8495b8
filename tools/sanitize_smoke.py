#!/usr/bin/env python3
"""Small batches through every kernel family, for `compute-sanitizer --tool memcheck|racecheck|initcheck`.

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
(The reference has no sanitizer story -- SURVEY.md section 5; its CUDA demo has a cross-thread race
in modcsw/modcmv's `static spint R`, section 2a.  Ragged sizes are used on purpose.)
"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from modarith_b200 import Field  # noqa: E402
from modarith_b200.rfc7748 import rfc7748  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(3)
for curve, nb in (("X25519", 32), ("X448", 56)):
    n = 257
    k = torch.randint(0, 256, (n, nb), dtype=torch.uint8, device="cuda", generator=g)
    u = torch.randint(0, 256, (n, nb), dtype=torch.uint8, device="cuda", generator=g)
    a = rfc7748(curve, k, u)
    b = rfc7748(curve, k, u, validate=True)
    h = rfc7748(curve, k.cpu().numpy(), u.cpu().numpy())                 # pageable host arrays: staged pipeline
    assert (a.cpu().numpy() == h).all()
    z = rfc7748(curve, k.cpu().pin_memory(), u.cpu().pin_memory())        # pinned: the kernel reads / writes host memory
    assert (a.cpu() == z).all()
for name in ("X25519", "X448", "NIST256"):
    F = Field(name)
    n = 131
    x, st = F.modimp(torch.randint(0, 256, (n, F.Nbytes), dtype=torch.uint8, device="cuda", generator=g))
    y, _ = F.modimp(torch.randint(0, 256, (n, F.Nbytes), dtype=torch.uint8, device="cuda", generator=g))
    r = F.alloc(n)
    F.modmul(x, y, r); F.modsqr(x, r); F.modadd(x, y, r); F.modsub(x, y, r); F.modneg(x, r)
    F.modmli(x, 121665, r); F.modinv(x, None, r); F.modsqrt(x, None, r); F.modpro(x, r)
    F.modqr(None, x); F.modis0(x); F.modis1(x); F.modsign(x); F.modcmp(x, y); F.modfsb(r)
    F.modhaf(r); F.modshl(3, r); F.modshr(3, r); F.mod2r(77, r); F.modint(5, r); F.modone(r); F.modzer(r)
    bits = torch.randint(0, 2, (n,), dtype=torch.int32, device="cuda", generator=g)
    F.modcsw(bits, x, y); F.modcmv(bits, x, y); F.nres(x, r); F.redc(x, r); F.modnsqr(r, 3); F.modcpy(x, r)
    F.modexp(r)
# scalar multiplication: ragged batch, table slices in the pooled global workspace (P-256) / shared memory (Ed25519)
import numpy as np  # noqa: E402
from modarith_b200.ecn import ecnmul  # noqa: E402
from modarith_b200.primes import PRIMES, X25519  # noqa: E402
n = 300
e = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
for curve, gx, gy in (("NIST256", PRIMES["NIST256"].wgx, PRIMES["NIST256"].wgy), ("ED25519", X25519.ed_gx, X25519.ed_gy)):
    x = torch.from_numpy(np.tile(np.frombuffer(gx.to_bytes(32, "big"), dtype=np.uint8), (n, 1))).cuda()
    y = torch.from_numpy(np.tile(np.frombuffer(gy.to_bytes(32, "big"), dtype=np.uint8), (n, 1))).cuda()
    xo, yo = ecnmul(curve, e, x, y)
    xo2, yo2 = ecnmul(curve, e, x, y)
    assert torch.equal(xo, xo2) and torch.equal(yo, yo2)
    from modarith_b200.ecn import ecnmul2  # noqa: E402
    f = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    ecnmul2(curve, e, x, y, f, xo, yo)
# a batch large enough for the per-sub-partition work queues with stealing (>= 8 groups per SM)
n = 148 * 8 * 32 + 77
k = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
u = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
a = rfc7748("X25519", k, u)
b = rfc7748("X25519", k[:1000].contiguous(), u[:1000].contiguous())
assert torch.equal(a[:1000], b)
# a straight-line program with its registers in shared memory
F = Field("NIST256")
x, _ = F.modimp(torch.randint(0, 256, (333, 32), dtype=torch.uint8, device="cuda", generator=g))
y, _ = F.modimp(torch.randint(0, 256, (333, 32), dtype=torch.uint8, device="cuda", generator=g))
F.modprog([("mul", 2, 0, 1), ("add", 3, 2, 0), ("sqr", 3, 3, 0), ("sub", 4, 3, 1), ("inv", 5, 4, 0), ("mli", 6, 5, 0, 7)], [x, y], [3, 5, 6])
# the same program compiled at run time (NVRTC) into a kernel of its own, and on an add-on modulus
prog = [("mul", 2, 0, 1), ("add", 3, 2, 0), ("sqr", 3, 3, 0), ("sub", 4, 3, 1), ("inv", 5, 4, 0), ("mli", 6, 5, 0, 7)]
a = F.modprog(prog, [x, y], [3, 5, 6])
b = F.modprog(prog, [x, y], [3, 5, 6], jit=True)
assert all(torch.equal(p, q) for p, q in zip(a, b))
import os
from modarith_b200 import lib as mlib
if os.path.exists(mlib.extra_lib_path("NIST384")):
    F3 = Field("NIST384")
    x3, _ = F3.modimp(torch.randint(0, 256, (517, 48), dtype=torch.uint8, device="cuda", generator=g))
    r3 = F3.alloc(517)
    F3.modmul(x3, x3, r3)
    F3.modinv(r3, None, r3)
    F3.modprog([("sqr", 1, 0, 0), ("sub", 1, 1, 0)], [x3], [1], jit=True)
torch.cuda.synchronize()
print("sanitize smoke done")
