#!/usr/bin/env python3
"""SASS opcode mix of the loop body of every k_probe<V> in a cuobjdump -sass listing (stdin or file)."""
import collections
import re
import sys


def main(path):
    txt = open(path).read() if path != "-" else sys.stdin.read()
    funcs = re.split(r'\n\s*Function : ', txt)[1:]
    res = {}
    for f in funcs:
        name = f.split('\n')[0]
        m = re.search(r'ILi(\d+)E', name)
        if not m:
            continue
        v = int(m.group(1))
        ins = []
        for l in f.split('\n'):
            mm = re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
            if mm:
                ins.append((int(mm.group(1), 16), mm.group(3), l))
        loop = None
        for a, op, l in ins:
            if op.startswith('BRA'):
                t = re.search(r'0x([0-9a-f]+)', l.split('BRA')[1])
                if t and int(t.group(1), 16) < a:
                    loop = (int(t.group(1), 16), a)
        if loop:
            body = [op for a, op, l in ins if loop[0] <= a <= loop[1]]
            c = collections.Counter(re.sub(r'\.U32|\.reuse|\.LUT|\.L\.W\.HI|\.NE\.AND|\.U$', '', o) for o in body)
            res[v] = (len(body), c)
    for v in sorted(res):
        n, c = res[v]
        print("%-3d n=%-3d %s" % (v, n, ", ".join("%s=%d" % kv for kv in sorted(c.items(), key=lambda kv: -kv[1]))))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "-")
