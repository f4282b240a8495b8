#!/usr/bin/env python3
"""Time the round-structured ladder (shared inversions) against the one-key-per-thread kernel."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from modarith_b200 import lib as mlib  # noqa: E402

l = mlib.load()
st = torch.cuda.current_stream().cuda_stream
for curve, nb in (("X25519", 32), ("X448", 56)):
    for lg in [int(v) for v in os.environ.get("LGS", "20 22").split()]:
        n = 1 << lg
        g = torch.Generator(device="cuda").manual_seed(1)
        k = torch.randint(0, 256, (n, nb), dtype=torch.uint8, device="cuda", generator=g)
        u = torch.randint(0, 256, (n, nb), dtype=torch.uint8, device="cuda", generator=g)
        o = torch.empty_like(k)
        for name in ("rfc7748", "rfc7748_perkey"):
            fn = getattr(l, "mab_%s_%s" % (curve, name))
            for _ in range(2):
                fn(k.data_ptr(), u.data_ptr(), o.data_ptr(), n, st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                fn(k.data_ptr(), u.data_ptr(), o.data_ptr(), n, st)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print("%-7s 2^%d %-16s %8.3f ms  %7.2f M/s" % (curve, lg, name, ms, n / ms / 1e3), flush=True)
