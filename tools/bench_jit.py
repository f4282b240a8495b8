#!/usr/bin/env python3
"""Compiled field programs (mab_<P>_modprog_jit) against the interpreter (mab_<P>_modprog) and one launch per call.

    python tools/bench_jit.py [lg_n]          # prints one line per program; MAB_JIT_MINBLOCKS=k is honoured

Programs: the complete P-256 point addition (weierstrass.c:69-160) and one Montgomery ladder step
(rfc7748.c:186-221 without the swaps) on the 2^255-19 and 2^448-2^224-1 fields."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modarith_b200 import Field            # noqa: E402
from modarith_b200 import lib as mlib      # noqa: E402

X1_, Y1_, Z1_, X2_, Y2_, Z2_, B_, t0, t1, t2, t3, t4, X3, Y3, Z3 = range(15)
POINT_ADD = [("mul", t0, X1_, X2_), ("mul", t1, Y1_, Y2_), ("mul", t2, Z1_, Z2_), ("add", t3, X1_, Y1_), ("add", t4, X2_, Y2_),
             ("mul", t3, t3, t4), ("add", t4, t0, t1), ("sub", t3, t3, t4), ("add", t4, Y1_, Z1_), ("add", X3, Y2_, Z2_),
             ("mul", t4, t4, X3), ("add", X3, t1, t2), ("sub", t4, t4, X3), ("add", X3, X1_, Z1_), ("add", Y3, X2_, Z2_),
             ("mul", X3, X3, Y3), ("add", Y3, t0, t2), ("sub", Y3, X3, Y3), ("mul", Z3, B_, t2), ("sub", X3, Y3, Z3),
             ("add", Z3, X3, X3), ("add", X3, X3, Z3), ("sub", Z3, t1, X3), ("add", X3, t1, X3), ("mul", Y3, B_, Y3),
             ("add", t1, t2, t2), ("add", t2, t1, t2), ("sub", Y3, Y3, t2), ("sub", Y3, Y3, t0), ("add", t1, Y3, Y3),
             ("add", Y3, t1, Y3), ("add", t1, t0, t0), ("add", t0, t1, t0), ("sub", t0, t0, t2), ("mul", t1, t4, Y3),
             ("mul", t2, t0, Y3), ("mul", Y3, X3, Z3), ("add", Y3, Y3, t2), ("mul", X3, t3, X3), ("sub", X3, X3, t1),
             ("mul", Z3, t4, Z3), ("mul", t1, t3, t0), ("add", Z3, Z3, t1)]


def ladder_step(a24):
    # registers: 0 x1, 1 x2, 2 z2, 3 x3, 4 z3; 5 A, 6 B, 7 C, 8 D, 9 t
    x1, x2, z2, x3, z3, A, B, C, D, t = range(10)
    return [("add", A, x2, z2), ("sub", B, x2, z2), ("add", C, x3, z3), ("sub", D, x3, z3), ("mul", D, D, A), ("mul", C, C, B),
            ("sqr", A, A, 0), ("sqr", B, B, 0), ("add", x3, D, C), ("sub", z3, D, C), ("sqr", x3, x3, 0), ("sqr", z3, z3, 0),
            ("mul", z3, z3, x1), ("mul", x2, A, B), ("sub", B, A, B), ("mli", t, B, 0, a24), ("add", t, t, A), ("mul", z2, t, B)]


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


def peak(dev):
    import ctypes
    lib = mlib.load()
    ms, ins = ctypes.c_float(), ctypes.c_double()
    best = 0.0
    for _ in range(3):
        mlib.check(lib.mab_imad_peak(0, 4000, 148 * 8, 256, ctypes.byref(ms), ctypes.byref(ins), None))
        best = max(best, ins.value / (ms.value * 1e-3))
    return best


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 21
    dev = torch.device("cuda:0")
    m = 1 << lg
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    pk = peak(dev)
    print("# n = 2^%d elements, IMAD.WIDE peak %.2f T/s, MAB_JIT_MINBLOCKS=%s" % (lg, pk / 1e12, os.environ.get("MAB_JIT_MINBLOCKS", "-")))
    jobs = [("NIST256", "complete point addition (14 mul, 29 add/sub)", POINT_ADD, 7, [X3, Y3, Z3]),
            ("X25519", "ladder step (5 mul, 4 sqr, mli, 8 add/sub)", ladder_step(121665), 5, [1, 2, 3, 4]),
            ("X448", "ladder step (5 mul, 4 sqr, mli, 8 add/sub)", ladder_step(39081), 5, [1, 2, 3, 4])]
    for prime, what, code, nin, outs in jobs:
        F = Field(prime, dev)
        L = F.Nlimbs
        ops = [F.modimp(torch.randint(0, 256, (m, F.Nbytes), dtype=torch.uint8, device=dev, generator=gen))[0] for _ in range(nin)]
        o1 = [F.alloc(m) for _ in outs]
        o2 = [F.alloc(m) for _ in outs]
        ti = timeit(lambda: F.modprog(code, ops, outs, outputs=o1), 5)
        t0 = time.time()
        F.modprog(code, ops, outs, outputs=o2, jit=True)
        torch.cuda.synchronize()
        tc = time.time() - t0
        tj = timeit(lambda: F.modprog(code, ops, outs, outputs=o2, jit=True), 5)
        same = all(bool(torch.equal(a, b)) for a, b in zip(o1, o2))
        prods = sum({"mul": L * L, "sqr": L * (L + 1) // 2, "mli": L}.get(c[0], 0) for c in code)
        print("%-8s %-46s interpreted %7.1f M/s (%.3f)   compiled %7.1f M/s (%.3f of the IMAD roofline)   x%.2f   first call %.2f s   identical %s"
              % (prime, what, m / ti / 1e6, m / ti * prods / pk, m / tj / 1e6, m / tj * prods / pk, ti / tj, tc, same))
        del ops, o1, o2


if __name__ == "__main__":
    main()
