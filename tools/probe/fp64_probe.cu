// fp64_probe.cu -- does the FP64 pipe of an sm_100a sub-partition run beside the integer multiplier pipe?
//
// The field code keeps the multiplier (FMA-heavy) pipe 85 % busy and leaves the FP64 pipe idle.  This probe
// measures what a sub-partition does with warps that issue DFMA / DADD streams next to warps that issue the
// ladder's IMAD.WIDE + LOP3 mix: cycles per loop iteration for each kind of warp, for several residencies.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu && ./fp64_probe
//
// One block per SM; warp w of a block runs on sub-partition w % 4, so "row" r = w / 4 holds one warp per
// sub-partition.  Rows [0, ni) run the integer block, rows [ni, ni + nf) the floating-point block.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

enum { B_INT = 0, B_DFMA = 1, B_DADD = 2, B_MIXED = 3, B_DFMA_LAT = 4, B_DFMA8 = 5, B_INT_ADC = 6 };

template <int KIND> __device__ __forceinline__ void block(uint32_t* a, uint64_t* acc, uint32_t& lg, double* x, double* d, uint32_t b, double y) {
  if constexpr (KIND == B_INT || KIND == B_MIXED || KIND == B_INT_ADC) {
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[(i + r) & 7]), "r"(a[(i + 3 + 2 * r) & 7]));   // 16 IMAD.WIDE.U32
  }
  if constexpr (KIND == B_INT || KIND == B_MIXED) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(a[i]) : "r"(lg), "r"(b));      // 8 LOP3
  }
  if constexpr (KIND == B_INT_ADC) {
    // 16 add-with-carry on the ALU pipe (4 chains of 4 words)
#pragma unroll
    for (int c = 0; c < 4; c++)
      asm volatile("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %6;\n\taddc.u32 %3, %3, %7;"
                   : "+r"(a[(4 * c) & 7]), "+r"(a[(4 * c + 1) & 7]), "+r"(a[(4 * c + 2) & 7]), "+r"(a[(4 * c + 3) & 7])
                   : "r"(b), "r"(lg), "r"(b), "r"(lg));
  }
  if constexpr (KIND == B_DFMA || KIND == B_MIXED) {
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(x[i & 1]), "d"(y));  // 16 DFMA, 8 accumulators
  }
  if constexpr (KIND == B_DFMA8) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int i = 0; i < 4; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(x[i & 1]), "d"(y));  // 16 DFMA, 4 accumulators
  }
  if constexpr (KIND == B_DADD) {
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(x[i & 1]));             // 16 DADD
  }
  if constexpr (KIND == B_DFMA_LAT) {
#pragma unroll
    for (int i = 0; i < 16; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[0]) : "d"(y), "d"(x[i & 1]));  // 16 dependent DFMA
  }
}

template <int KI, int KF> __global__ void __launch_bounds__(1024, 1)
k_probe(const uint32_t* seed, uint32_t* sink, long long* cyc, int iters, int ni) {
  uint32_t a[8];
  uint64_t acc[8];
  double x[2], d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = seed[(threadIdx.x + i) & 63] + i;
    acc[i] = a[i];
    d[i] = (double)i;
  }
  x[0] = 1.0 + (double)(a[0] & 1023) * (1.0 / 1048576.0);
  x[1] = 1.0 + (double)(a[1] & 1023) * (1.0 / 1048576.0);
  uint32_t lg = a[3], b = seed[5] | 1u;
  const double y = 1.0 + 1.0 / 1073741824.0;
  const int row = threadIdx.x >> 7;
  __syncthreads();
  long long t0 = clock64();
  if (row < ni) {
#pragma unroll 1
    for (int it = 0; it < iters; it++) block<KI>(a, acc, lg, x, d, b, y);
  } else {
#pragma unroll 1
    for (int it = 0; it < iters; it++) block<KF>(a, acc, lg, x, d, b, y);
  }
  long long t1 = clock64();
  uint32_t s = lg;
  double ds = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { s ^= (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32) ^ a[i]; ds += d[i]; }
  if (s == seed[9] && ds == (double)seed[7]) sink[0] = s;                                  // keep everything alive
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int KI, int KF> static void run(const char* name, int ni, int nf, int iters, int sms, uint32_t* d, long long* cyc) {
  const int rows = ni + nf;
  k_probe<KI, KF><<<sms, rows * 128>>>(d, d + 64, cyc, iters, ni);
  cudaDeviceSynchronize();
  k_probe<KI, KF><<<sms, rows * 128>>>(d, d + 64, cyc, iters, ni);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long* h = (long long*)malloc(sizeof(long long) * sms * 32);
  cudaMemcpy(h, cyc, sizeof(long long) * sms * 32, cudaMemcpyDeviceToHost);
  double si = 0, sf = 0;
  for (int s = 0; s < sms; s++)
    for (int w = 0; w < rows * 4; w++) ((w / 4) < ni ? si : sf) += (double)h[s * 32 + w];
  free(h);
  const double ci = ni ? si / (sms * ni * 4) / iters : 0, cf = nf ? sf / (sms * nf * 4) / iters : 0;
  // per sub-partition: blocks per 1000 cycles of each kind
  printf("  %-46s int warps %d  fp warps %d | cycles/iteration int %7.1f  fp %7.1f | blocks per 1000 cycles per sub-partition: int %6.2f  fp %6.2f\n",
         name, ni, nf, ci, cf, ni ? 1000.0 * ni / ci : 0.0, nf ? 1000.0 * nf / cf : 0.0);
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 20000;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t host[64];
  for (int i = 0; i < 64; i++) host[i] = 0x9e3779b9u * (i + 1) | 1u;
  uint32_t* d; long long* cyc;
  cudaMalloc(&d, 80 * sizeof(uint32_t));
  cudaMalloc(&cyc, sms * 32 * sizeof(long long));
  cudaMemcpy(d, host, sizeof(host), cudaMemcpyHostToDevice);
  printf("# fp64_probe: %d SMs, %d iterations; int block = 16 IMAD.WIDE + 16 IADD3[.X] + 8 LOP3 (ptxas adds pairs of products with three-input adds), fp block = 16 DFMA (or as named)\n", sms, iters);
  for (int w = 1; w <= 8; w *= 2) run<B_INT, B_INT>("int only", w, 0, iters, sms, d, cyc);
  for (int w = 1; w <= 8; w *= 2) run<B_DFMA, B_DFMA>("DFMA 16, 8 accumulators", 0, w, iters, sms, d, cyc);
  for (int w = 1; w <= 4; w *= 2) run<B_DFMA8, B_DFMA8>("DFMA 16, 4 accumulators", 0, w, iters, sms, d, cyc);
  for (int w = 1; w <= 4; w *= 2) run<B_DADD, B_DADD>("DADD 16, 8 accumulators", 0, w, iters, sms, d, cyc);
  for (int w = 1; w <= 2; w *= 2) run<B_DFMA_LAT, B_DFMA_LAT>("16 dependent DFMA (latency)", 0, w, iters, sms, d, cyc);
  for (int w = 1; w <= 4; w *= 2) run<B_MIXED, B_MIXED>("same warp: 16 IMAD.WIDE + 8 LOP3 + 16 DFMA", w, 0, iters, sms, d, cyc);
  const int mix[][2] = {{1, 1}, {2, 1}, {3, 1}, {4, 1}, {2, 2}, {3, 2}, {4, 2}, {4, 4}, {6, 2}};
  for (auto& m : mix) run<B_INT, B_DFMA>("int warps beside DFMA warps", m[0], m[1], iters, sms, d, cyc);
  for (auto& m : mix) run<B_INT_ADC, B_DFMA>("int(16 WIDE + 16 add-with-carry) beside DFMA", m[0], m[1], iters, sms, d, cyc);
  const int mix2[][2] = {{2, 1}, {4, 1}, {4, 2}};
  for (auto& m : mix2) run<B_INT, B_DADD>("int warps beside DADD warps", m[0], m[1], iters, sms, d, cyc);
  return 0;
}
