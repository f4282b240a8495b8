#!/usr/bin/env python3
"""Instruction mix of the loops in a `cuobjdump -sass` listing (static analysis, no GPU).

usage: cuobjdump -sass -fun <kernel> file.o | python tools/sass_loops.py
For each backward branch prints the span, instruction count and the count per opcode
family, with IMAD.WIDE separated out (it occupies the multiplier pipe twice as long).
"""
import re
import sys
from collections import Counter

ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
print("total instructions:", len(ins))
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt < a and tgt in addr_index:
        body = ins[addr_index[tgt]:i + 1]
        c = Counter()
        for _, s in body:
            s = re.sub(r"^@!?U?P\w+\s+", "", s)
            op = s.split()[0]
            if op.startswith("IMAD.WIDE"):
                key = "IMAD.WIDE"
            elif op.startswith("IMAD.MOV") or op.startswith("IMAD.SHL") or op.startswith("IMAD.IADD"):
                key = "IMAD(mov/shl/iadd)"
            else:
                key = op.split(".")[0]
            c[key] += 1
        print("loop 0x%x..0x%x: %d instructions" % (tgt, a, len(body)))
        print("   ", ", ".join("%s=%d" % kv for kv in c.most_common()))
