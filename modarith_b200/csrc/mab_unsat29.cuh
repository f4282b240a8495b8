// mab_unsat29.cuh -- COMPARISON KERNEL, not on the product path.
//
// 2^255-19 modmul on the reference's own WL=32 limb plan: unsaturated radix 2^29, 9 limbs
// (getbase, pseudo.py:124-140; 81 limb products + 9 fold multiplies per modmul with the high half
// carried before it is folded because (b-1)^2*mm*N >= 2^64, pseudo.py:1640-1648).  Columns are
// accumulated in 64-bit registers with IMAD.WIDE (no carry flag needed), carries are propagated with
// shifts and masks.  It exists so that the choice of SATURATED limbs (gen/plan.py) is backed by a
// measurement on the same GPU instead of an argument: bench.py times a register-resident chain of
// these next to the saturated one, tests/test_gpu_field.py checks it against the oracle.
#pragma once
#include <stdint.h>

struct Unsat29 {
  static constexpr int N = 9;
  static constexpr uint32_t MASK = (1u << 29) - 1u;
  static constexpr uint32_t FOLD = 19u << 6;          // 2^261 = 2^(9*29) == 19 * 2^6  (mod 2^255-19)

  // c = a*b, limbs of the result < 2^29 (+2^15 for limb 1); inputs may be that loose
  static __device__ __forceinline__ void mul(uint32_t (&c)[N], const uint32_t (&a)[N], const uint32_t (&b)[N]) {
    uint64_t t[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) t[k] = 0;
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j < N; j++) t[i + j] += (uint64_t)a[i] * b[j];
    // carry the high columns so that each fits 29 bits before it is multiplied by FOLD
#pragma unroll
    for (int k = N; k < 2 * N - 1; k++) {
      t[k + 1] += t[k] >> 29;
      t[k] &= MASK;
    }
    // fold: column k+9 counts FOLD times at column k (t[9..16] < 2^29, t[17] < 2^30 after the carries)
#pragma unroll
    for (int k = 0; k < N; k++) t[k] += (uint64_t)FOLD * (uint32_t)t[k + N];
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < N; k++) {
      t[k] += carry;
      carry = t[k] >> 29;
      c[k] = (uint32_t)t[k] & MASK;
    }
    // carry out of limb 8 wraps with FOLD
    uint64_t w = carry * FOLD + c[0];
    c[0] = (uint32_t)w & MASK;
    w = (w >> 29) + c[1];
    c[1] = (uint32_t)w;                               // < 2^29 + 2^15: left loose
  }
};

// register-resident chain c = a * b^iters on 9-limb planes (limb j of element i at p[j*stride+i])
__global__ void __launch_bounds__(128) k_unsat29_mulchain(const uint32_t* a, const uint32_t* b, uint32_t* c,
                                                          unsigned iters, size_t n, size_t stride) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x[Unsat29::N], y[Unsat29::N];
#pragma unroll
  for (int j = 0; j < Unsat29::N; j++) { x[j] = a[j * stride + i]; y[j] = b[j * stride + i]; }
#pragma unroll 1
  for (unsigned it = 0; it < iters; it++) Unsat29::mul(x, x, y);
#pragma unroll
  for (int j = 0; j < Unsat29::N; j++) c[j * stride + i] = x[j];
}
