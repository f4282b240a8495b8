#!/usr/bin/env python3
"""SASS instruction mix of each generated field function in isolation (no GPU needed).

usage: python tools/sass_count.py NIST256 [X25519 ...]
Compiles one tiny kernel per function (operands loaded from / stored to global memory) and
prints the instruction mix minus the fixed load/store/address overhead of an empty kernel."""
import os, re, subprocess, sys, tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "modarith_b200", "csrc")
FUNCS = {
    "nop": "for (int i = 0; i < L; i++) r[i] = a[i] ^ b[i];",
    "mul": "F::mul(r, a, b);", "sqr": "F::sqr(r, a); (void)b;", "add": "F::add(r, a, b);",
    "sub": "F::sub(r, a, b);", "mli": "F::mli(r, a, b[0]);", "mla": "F::mla(r, a, 121665u, b);",
    "neg": "F::neg(r, a); (void)b;", "add_tt": "F::add_tt(r, a, b);", "sub_tt": "F::sub_tt(r, a, b);",
}

def mix(name):
    src = ['#include "gen/field_%s.cuh"' % name, "typedef F_%s F; constexpr int L = F::L;" % name]
    for fn, body in FUNCS.items():
        src.append("""extern "C" __global__ void k_%s(const uint32_t* pa, const uint32_t* pb, uint32_t* pr) {
  uint32_t a[L], b[L], r[L]; int t = threadIdx.x;
  for (int i = 0; i < L; i++) { a[i] = pa[i * 32 + t]; b[i] = pb[i * 32 + t]; }
  %s
  for (int i = 0; i < L; i++) pr[i * 32 + t] = r[i];
}""" % (fn, body))
    with tempfile.TemporaryDirectory() as d:
        cu, cubin = os.path.join(d, "t.cu"), os.path.join(d, "t.cubin")
        open(cu, "w").write("\n".join(src))
        subprocess.check_call(["nvcc", "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-gencode",
                               "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-cubin", cu, "-o", cubin])
        out = {}
        for fn in FUNCS:
            t = subprocess.run(["cuobjdump", "-sass", "-fun", "k_" + fn, cubin], stdout=subprocess.PIPE, text=True).stdout
            c = Counter()
            for line in t.splitlines():
                m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
                if not m:
                    continue
                s = re.sub(r"^@!?U?P\w+\s+", "", m.group(1).strip())
                op = s.split()[0]
                if op.startswith("IMAD.WIDE"): key = "WIDE"
                elif op.startswith(("IMAD.MOV", "IMAD.SHL", "IMAD.IADD")): key = "IMAD.mov"
                else: key = op.split(".")[0]
                c[key] += 1
            out[fn] = c
    return out

for name in sys.argv[1:]:
    m = mix(name)
    base = m.pop("nop")
    base["LOP3"] -= 0
    print("== %s (minus the load/store skeleton: %d instructions)" % (name, sum(base.values()) - base.get("LOP3", 0)))
    for fn, c in m.items():
        d = Counter(c); d.subtract(base); d["LOP3"] += base.get("LOP3", 0)
        d = {k: v for k, v in d.items() if v}
        print("  %-4s total %4d : %s" % (fn, sum(d.values()), ", ".join("%s=%d" % kv for kv in sorted(d.items(), key=lambda kv: -kv[1]))))
