#!/usr/bin/env python3
"""Opcode mix of every loop of one kernel in an object file, classified by issue pipe (no GPU needed).

    python tools/sass_loopmix.py modarith_b200/build/mab_capi_X25519.o k_rfc7748_rounds

For each backward branch the instructions between its target and the branch are counted:
  W     IMAD.WIDE[.X]           (multiplier pipe, 4 cycles per warp instruction)
  I     other IMAD.* / IMAD.HI  (multiplier pipe, 2 cycles; .HI 4)
  A     everything else that executes on the ALU pipe (IADD3, LOP3, SHF, SEL, MOV, ...)
and the issue model fitted in profiles/r2_issue_probe.txt is evaluated on them.

The count is STATIC: a loop whose body holds both arms of a run-time choice shows both (the X448 ladder step carries
the z3 = x1 * (DA-CB)^2 product twice, once per arm of `if (stash)`: 1600 wide multiplies in the listing, 1404 executed
per step -- DESIGN.md section 9, item 4).
"""
import collections
import re
import subprocess
import sys


def loops(obj, pattern):
    txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    for f in re.split(r'\n\s*Function : ', txt)[1:]:
        name = f.split('\n')[0]
        if pattern not in name:
            continue
        ins = []
        for l in f.split('\n'):
            mm = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
            if mm:
                ins.append((int(mm.group(1), 16), mm.group(3), l))
        found = []
        for a, op, l in ins:
            if op.startswith('BRA'):
                t = re.search(r'0x([0-9a-f]+)', l.split('BRA')[1])
                if t and int(t.group(1), 16) < a:
                    found.append((int(t.group(1), 16), a))
        yield name, ins, found


def classify(op):
    if op.startswith("IMAD.WIDE"):
        return "W"
    if op.startswith("IMAD") or op.startswith("HFMA2") or op.startswith("FFMA"):
        return "I"
    if op.startswith(("BRA", "NOP", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "LD", "ST", "ATOM", "RED", "S2R", "SHFL", "BAR", "CS2R", "R2UR", "CALL", "RET")):
        return "X"
    return "A"


def model(W, I, A):
    """cycles per warp on one sub-partition (fit of profiles/r2_issue_probe.txt)"""
    nf = W + I
    return 4 * W + 2 * I + 0.7 * min(A, nf) + 2.0 * max(0, A - nf)


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    for name, ins, found in loops(obj, pat):
        print("==", name)
        for lo, hi in sorted(found, key=lambda r: r[1] - r[0]):
            body = [op for a, op, l in ins if lo <= a <= hi]
            if len(body) < 20:
                continue
            c = collections.Counter(re.sub(r'\.U32|\.reuse', '', o) for o in body)
            k = collections.Counter(classify(o) for o in body)
            print("loop 0x%x..0x%x: n=%d  W=%d I=%d A=%d other=%d  model=%.0f cycles (4W=%d)" % (
                lo, hi, len(body), k["W"], k["I"], k["A"], k["X"], model(k["W"], k["I"], k["A"]), 4 * k["W"]))
            print("   " + ", ".join("%s=%d" % kv for kv in sorted(c.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
