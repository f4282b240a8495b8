#!/usr/bin/env python3
"""Run the instruction-mix probes on the GPU and print cycles per PTX block per warp.

    python tools/run_probe.py [warps_per_smsp ...]     (default 1 2 4 8)
"""
import ctypes
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from modarith_b200 import lib as mlib  # noqa: E402


def main():
    import torch
    lib = mlib.load()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    clk = 1.965e9
    iters = 20000
    res = []
    wlist = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
    for v in range(64):
        row = None
        for w in wlist:                       # warps per SMSP: block of 128 threads = 1 warp per SMSP
            ms, name, nw, na = ctypes.c_float(), ctypes.c_char_p(), ctypes.c_int(), ctypes.c_int()
            rc = lib.mab_pipe_probe(v, iters, sms * w, 128, ctypes.byref(ms), ctypes.byref(name), ctypes.byref(nw),
                                    ctypes.byref(na), None)
            if rc != 0:
                break
            cyc_per_block_per_smsp = ms.value * 1e-3 * clk / iters / w     # SMSP cycles per warp-block
            if row is None:
                row = {"variant": v, "name": name.value.decode(), "wide": nw.value, "alu": na.value, "cycles": {}}
            row["cycles"][w] = cyc_per_block_per_smsp
        if row is None:
            break
        res.append(row)
        print("%2d %-42s " % (row["variant"], row["name"]) +
              "  ".join("w%d: %6.1f" % (w, c) for w, c in row["cycles"].items()), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
