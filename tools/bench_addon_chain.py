#!/usr/bin/env python3
"""Register-resident modmul chain (mab_<P>_bench_modmul) of add-on moduli, as a fraction of the IMAD roofline.

    python tools/bench_addon_chain.py NAME [NAME ...]     # libraries built with python -m modarith_b200.build --prime"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modarith_b200 import Field            # noqa: E402
from modarith_b200 import lib as mlib      # noqa: E402


def main():
    dev = torch.device("cuda:0")
    lib = mlib.load()
    ms, ins = ctypes.c_float(), ctypes.c_double()
    pk = 0.0
    for _ in range(3):
        mlib.check(lib.mab_imad_peak(0, 4000, 148 * 8, 256, ctypes.byref(ms), ctypes.byref(ins), None))
        pk = max(pk, ins.value / (ms.value * 1e-3))
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    for name in sys.argv[1:]:
        F = Field(name, dev)
        m, iters, L = 1 << 20, 200, F.Nlimbs
        x = F.modimp(torch.randint(0, 128, (m, F.Nbytes), dtype=torch.uint8, device=dev, generator=gen))[0]
        y = F.modimp(torch.randint(0, 128, (m, F.Nbytes), dtype=torch.uint8, device=dev, generator=gen))[0]
        r = F.alloc(m)
        fn = getattr(F.lib, "mab_%s_bench_modmul" % name)
        st = torch.cuda.current_stream(dev).cuda_stream

        def run():
            mlib.check(fn(x.data_ptr(), y.data_ptr(), r.data_ptr(), iters, m, m, st), "bench_modmul", F.lib)
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        run()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 2e3
        print("%-10s %4d bits %2d limbs   modmul chain %7.1f Gop/s   %.3f of the IMAD roofline (L^2 = %d products)"
              % (name, F.Nbits, L, m * iters / t / 1e9, m * iters / t * L * L / pk, L * L))


if __name__ == "__main__":
    main()
