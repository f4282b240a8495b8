#!/usr/bin/env python3
"""Generate tools/probe/issue_probe.cu: a stand-alone micro-benchmark that settles how an sm_100a
sub-partition issues the integer instructions the field code is made of.

Every variant is ONE inline-PTX block executed in a counted loop by W warps per sub-partition on every SM;
the program prints cycles per block per warp-slot, i.e. the time one sub-partition needs to issue the block
once for one warp when W warps keep it busy.  tools/probe/run.sh prints the SASS opcode mix of every loop
next to it, because ptxas re-pipes integer adds at will (IADD3 <-> IMAD.IADD / IMAD.X): the SASS, not the
PTX, is what the numbers belong to.

Register file of a block: %0..%15 = sixteen loop-carried 32-bit registers, %16 = y, %17 = z (never written).
"""
import os

NA, NM, NR = 8, 8, 16


def ar(i): return "%%%d" % i
Y, Z = "%16", "%17"
NR = 16


# Multiplier work is written the way the field code writes it: the 64-bit accumulators are block-local
# (lo, hi) register pairs born from a product (mul.lo + mul.hi = IMAD.WIDE with a zero addend), a second
# product is accumulated on each (mad.lo.cc + madc.hi = IMAD.WIDE with the pair as addend), and the pairs are
# folded into the loop-carried ALU registers with one 3-input LOP3 each, so that nothing is loop-invariant
# and ptxas has no reason to move registers around.  One unit = 16 IMAD.WIDE + 8 LOP3.
def wide_unit(chains=False):
    first = [["mul.lo.u32 l%d, %s, %s;" % (w, ar(w), Y), "mul.hi.u32 h%d, %s, %s;" % (w, ar(w), Y)] for w in range(8)]
    if not chains:
        second = [["mad.lo.cc.u32 l%d, %s, %s, l%d;" % (w, ar(8 + w), Z, w), "madc.hi.u32 h%d, %s, %s, h%d;" % (w, ar(8 + w), Z, w)]
                  for w in range(8)]
    else:
        second = []
        for c in range(2):
            g = []
            for n in range(4):
                w = 4 * c + n
                g.append("%s l%d, %s, %s, l%d;" % ("mad.lo.cc.u32" if n == 0 else "madc.lo.cc.u32", w, ar(8 + w), Z, w))
                g.append("%s h%d, %s, %s, h%d;" % ("madc.hi.cc.u32" if n < 3 else "madc.hi.u32", w, ar(8 + w), Z, w))
            second.append(g)
    fold = [["lop3.b32 %s, %s, l%d, h%d, 0x96;" % (ar(w), ar(w), w, w)] for w in range(8)]
    return first, second, fold


def addc_chain(idx):
    out = []
    for n, j in enumerate(idx):
        op = "add.cc.u32" if n == 0 else ("addc.cc.u32" if n < len(idx) - 1 else "addc.u32")
        out.append("%s %s, %s, %s;" % (op, ar(j), ar(j), ar((j + 3) % NR)))
    return out


def subc_chain(idx):
    out = []
    for n, j in enumerate(idx):
        op = "sub.cc.u32" if n == 0 else ("subc.cc.u32" if n < len(idx) - 1 else "subc.u32")
        out.append("%s %s, %s, %s;" % (op, ar(j), ar(j), ar((j + 3) % NR)))
    return out


def lop3(j): return ["lop3.b32 %s, %s, %s, %s, 0x96;" % (ar(j), ar(j), ar((j + 1) % NR), ar((j + 2) % NR))]
def shf(j): return ["shf.l.wrap.b32 %s, %s, %s, 7;" % (ar(j), ar(j), ar((j + 1) % NR))]
def sel(j): return ["selp.u32 %s, %s, %s, p;" % (ar(j), ar(j), ar((j + 5) % NR))]
def imad32(j): return ["mad.lo.u32 %s, %s, %s, %s;" % (ar(j), ar(j), Y, ar((j + 1) % NR))]
def imadhi(j): return ["mad.hi.u32 %s, %s, %s, %s;" % (ar(j), ar(j), Y, ar((j + 1) % NR))]


def interleave(a_groups, b_groups):
    """spread the b groups evenly between the a groups"""
    out = []
    nb = len(b_groups)
    na = max(1, len(a_groups))
    k = 0
    for i, g in enumerate(a_groups):
        out += g
        want = (i + 1) * nb // na
        while k < want:
            out += b_groups[k]
            k += 1
    while k < nb:
        out += b_groups[k]
        k += 1
    return out


def with_wides(alu_groups, chains=False):
    f, s_, fold = wide_unit(chains)
    na = len(alu_groups)
    h = na // 2
    return interleave(f, alu_groups[:h]) + interleave(s_, alu_groups[h:]) + sum(fold, [])


V = []
V.append(("16 IMAD.WIDE + 8 LOP3", with_wides([])))
V.append(("8 WIDE + 2x4 WIDE.X chains + 8 LOP3", with_wides([], chains=True)))
V.append(("16 add-with-carry (2 chains x 8)", addc_chain(range(0, 8)) + addc_chain(range(8, 16))))
V.append(("16 add-with-carry (4 chains x 4)", sum([addc_chain(range(4 * c, 4 * c + 4)) for c in range(4)], [])))
V.append(("16 sub-with-borrow (2 chains x 8)", subc_chain(range(0, 8)) + subc_chain(range(8, 16))))
V.append(("16 LOP3", sum([lop3(j) for j in range(16)], [])))
V.append(("16 SHF", sum([shf(j) for j in range(16)], [])))
V.append(("16 SEL", sum([sel(j) for j in range(16)], [])))
V.append(("16 IMAD.LO (32-bit)", sum([imad32(j) for j in range(16)], [])))
V.append(("16 IMAD.HI", sum([imadhi(j) for j in range(16)], [])))
V.append(("8 LOP3 + 8 SHF", sum([lop3(j) + shf(j + 8) for j in range(8)], [])))
for n_alu in (8, 16, 24, 32, 48):
    V.append(("16 WIDE + 8 LOP3 + %d add-with-carry (chains of 4)" % n_alu,
              with_wides([addc_chain([(4 * c + t) % 16 for t in range(4)]) for c in range(n_alu // 4)])))
for n_alu in (8, 24, 40):
    V.append(("16 WIDE + %d LOP3" % (8 + n_alu), with_wides([lop3(j % 16) for j in range(n_alu)])))
V.append(("16 WIDE + 8 LOP3 + 16 SHF", with_wides([shf(j) for j in range(16)])))
V.append(("16 WIDE + 8 LOP3 + 16 SEL", with_wides([sel(j) for j in range(16)])))
V.append(("16 WIDE + 8 LOP3 + 16 IMAD.LO", with_wides([imad32(j) for j in range(16)])))
V.append(("16 WIDE + 8 LOP3 + 32 IMAD.LO", with_wides([imad32(j % 16) for j in range(32)])))
V.append(("8 WIDE + 2x4 WIDE.X + 8 LOP3 + 32 add-with-carry (4x8)",
          with_wides([addc_chain(range(0, 8)), addc_chain(range(8, 16)), addc_chain(range(0, 8)), addc_chain(range(8, 16))], chains=True)))
V.append(("16 IMAD.LO + 16 LOP3", interleave([imad32(j) for j in range(16)], [lop3(j) for j in range(16)])))
V.append(("16 IMAD.LO + 16 add-with-carry (4x4)", interleave([imad32(j) for j in range(16)],
                                                             [addc_chain(range(4 * c, 4 * c + 4)) for c in range(4)])))
# latency probes: ONE dependent chain (read the W=1 column)
V.append(("lat: 8 dependent IMAD.WIDE (low word feeds the next multiplicand)",
          ["mul.lo.u32 l0, %0, %16;", "mul.hi.u32 h0, %0, %16;"] +
          sum([["mad.lo.cc.u32 l%d, l%d, %%16, l%d;" % (k + 1, k, k), "madc.hi.u32 h%d, l%d, %%16, h%d;" % (k + 1, k, k)] for k in range(7)], []) +
          ["lop3.b32 %0, %0, l7, h7, 0x96;"]))
V.append(("lat: 16-long add-with-carry chain", addc_chain(range(16))))
V.append(("lat: 16 dependent LOP3", sum([["lop3.b32 %0, %0, %1, %2, 0x96;"] for _ in range(16)], [])))
V.append(("lat: 16 dependent IMAD.LO", sum([["mad.lo.u32 %0, %0, %16, %1;"] for _ in range(16)], [])))


def main():
    out = []
    out.append("// generated by tools/probe/gen_issue_probe.py -- do not edit\n")
    out.append("#include <cuda_runtime.h>\n#include <stdint.h>\n#include <stdio.h>\n#include <stdlib.h>\n#include <string.h>\n\n")
    out.append("#define NV %d\n" % len(V))
    out.append("static const char* kNames[NV] = {\n")
    for name, _ in V:
        out.append('  "%s",\n' % name)
    out.append("};\n\n")
    out.append("template <int VAR> __global__ void __launch_bounds__(1024, 1) k_probe(const uint32_t* seed, uint32_t* sink, long long* cyc, int iters) {\n")
    out.append("  uint32_t r[18];\n#pragma unroll\n  for (int i = 0; i < 18; i++) r[i] = seed[(threadIdx.x + i) & 63] + i;\n")
    out.append("  long long t0 = clock64();\n#pragma unroll 1\n  for (int it = 0; it < iters; it++) {\n")
    ops = ", ".join('"+r"(r[%d])' % i for i in range(16))
    for v, (name, lines) in enumerate(V):
        out.append("    if constexpr (VAR == %d) {\n      asm volatile(\"{\\n\\t.reg .pred p;\\n\\t.reg .u32 l<8>, h<8>;\\n\\tsetp.ne.u32 p, %%17, 0;\\n\\t\"\n" % v)
        for ln in lines:
            out.append('        "%s\\n\\t"\n' % ln)
        out.append('        "}"\n        : %s\n        : "r"(r[16]), "r"(r[17]));\n    }\n' % ops)
    out.append("  }\n  long long t1 = clock64();\n  uint32_t acc = 0;\n#pragma unroll\n  for (int i = 0; i < 18; i++) acc ^= r[i];\n")
    out.append("  if (acc == 0x12345678u) sink[0] = acc;\n")
    out.append("  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;\n}\n\n")
    out.append("template <int VAR> static void run(int v, int warps, int iters, int sms, const uint32_t* d, uint32_t* sink, long long* cyc, double* out) {\n")
    out.append("  if (v == VAR) {\n    for (int rep = 0; rep < 2; rep++) k_probe<VAR><<<sms, warps * 128>>>(d, sink, cyc, iters);\n")
    out.append("    cudaDeviceSynchronize();\n    int nw = sms * warps * 4;\n    long long* h = (long long*)malloc(nw * sizeof(long long));\n")
    out.append("    cudaMemcpy(h, cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost);\n    double s = 0;\n    for (int i = 0; i < nw; i++) s += (double)h[i];\n")
    out.append("    free(h);\n    *out = s / nw / iters / warps;\n  }\n  if constexpr (VAR + 1 < NV) run<VAR + 1>(v, warps, iters, sms, d, sink, cyc, out);\n}\n\n")
    out.append("""int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 20000;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t host[64];
  for (int i = 0; i < 64; i++) host[i] = 0x9e3779b9u * (i + 1) | 1u;
  uint32_t* d; long long* cyc;
  cudaMalloc(&d, 80 * sizeof(uint32_t));
  cudaMalloc(&cyc, sms * 32 * sizeof(long long));
  cudaMemcpy(d, host, sizeof(host), cudaMemcpyHostToDevice);
  printf("# cycles per block per warp slot: one sub-partition issuing the block once, W warps resident on it (%d SMs, %d iterations)\\n", sms, iters);
  printf("# %-3s %-52s %8s %8s %8s %8s\\n", "id", "block", "W=1", "W=2", "W=4", "W=8");
  for (int v = 0; v < NV; v++) {
    printf("  %-3d %-52s", v, kNames[v]);
    for (int w = 1; w <= 8; w *= 2) {
      double c = 0;
      run<0>(v, w, iters, sms, d, d + 64, cyc, &c);
      printf(" %8.1f", c);
    }
    printf("\\n");
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
""")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "issue_probe.cu")
    with open(path, "w") as f:
        f.write("".join(out))
    print("wrote", path, len(V), "variants")


if __name__ == "__main__":
    main()
