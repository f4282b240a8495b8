#!/usr/bin/env python3
"""Compare gpurun_out/issue_probe.txt (tools/probe/run.sh) with the issue model used by tools/sass_loopmix.py:

    T = 4*W + 2*I + 0.7*min(A, W+I) + 2*max(0, A - (W+I))      cycles per warp on one sub-partition

W = IMAD.WIDE[.X] (multiplier pipe, 4 cycles), I = other multiplier-pipe instructions (IMAD, IMAD.X, IMAD.MOV ...,
2 cycles; IMAD.HI 4), A = ALU-pipe instructions (IADD3[.X], LOP3, SHF, SEL, MOV ..., 2 cycles on their own pipe).
Reading: an ALU instruction hides behind a multiplier-pipe instruction for ~0.7 cycles, one per multiplier-pipe
instruction; ALU instructions beyond that cost their full 2 cycles.  usage: fit_model.py [issue_probe.txt]"""
import re
import sys


def main(path):
    cyc, mix, names = {}, {}, {}
    for line in open(path):
        m = re.match(r"\s+(\d+)\s+(.*?)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s*$", line)
        if m:
            cyc[int(m.group(1))] = [float(m.group(k)) for k in (3, 4, 5, 6)]
            names[int(m.group(1))] = m.group(2)
        m = re.match(r"(\d+)\s+n=(\d+)\s+(.*)$", line)
        if m:
            mix[int(m.group(1))] = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in m.group(3).split(", "))
    print("# id  W   I   A   measured W=4 / W=8   model   ratio(W=4)  ratio(W=8)   block")
    for v in sorted(cyc):
        if v not in mix or names[v].startswith("lat:"):
            continue
        W = sum(n for k, n in mix[v].items() if k.startswith("IMAD.WIDE"))
        H = sum(n for k, n in mix[v].items() if k.startswith("IMAD.HI"))
        I = sum(n for k, n in mix[v].items() if (k.startswith("IMAD") or k.startswith("HFMA2")) and not k.startswith("IMAD.WIDE") and not k.startswith("IMAD.HI"))
        A = sum(n for k, n in mix[v].items() if not k.startswith(("IMAD", "HFMA2", "BRA", "UIADD3", "UISETP")))
        nf = W + I + H
        t = 4 * (W + H) + 2 * I + 0.7 * min(A, nf) + 2.0 * max(0, A - nf)
        print("  %-3d %3d %3d %3d   %8.1f / %6.1f   %6.1f     %5.2f       %5.2f      %s" % (
            v, W + H, I, A, cyc[v][2], cyc[v][3], t, cyc[v][2] / t, cyc[v][3] / t, names[v]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/issue_probe.txt")
