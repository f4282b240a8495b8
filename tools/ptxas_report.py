#!/usr/bin/env python3
"""Registers / spills per kernel from the build logs (`-Xptxas -v`), optionally filtered by substring."""
import glob, os, re, sys
pat = sys.argv[1] if len(sys.argv) > 1 else ""
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "modarith_b200", "build")
for log in sorted(glob.glob(os.path.join(root, "*.o.log"))):
    t = open(log).read()
    for b in re.split(r"ptxas info\s+: Compiling entry function ", t)[1:]:
        name = b.split("'")[1]
        if pat not in name:
            continue
        m = re.search(r"Used (\d+) registers", b)
        sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
        print("%-22s %-70s regs=%s spill=%s/%s" % (os.path.basename(log)[:-6], name[:70], m.group(1), sp.group(1), sp.group(2)))
