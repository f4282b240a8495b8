// mab_common.cuh -- shared definitions for the generated field headers.
//
// Two build modes:
//   default        device code for sm_100a (nvcc); the arithmetic blocks are inline PTX.
//   MAB_HOSTSIM    TEST SCAFFOLDING ONLY: the same headers compiled by g++ with each PTX
//                  block replaced by its plain-C transcription (gen/ptx.py emit_sim), so
//                  the CPU test-suite can run the device-side logic without a GPU.  The
//                  shipped library (csrc/mab_capi.cu) never defines it and has no CPU path.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef MAB_HOSTSIM
#include <stdio.h>
#include <stdlib.h>
#define MAB_DEV inline
#define MAB_NOUNROLL
// call-site invariants of the generated code are CHECKED in the host simulation (the device build trusts them)
#define MAB_SIM_REQUIRE(cond, what) do { if (!(cond)) { fprintf(stderr, "hostsim: violated precondition: %s\n", what); abort(); } } while (0)
struct uint4 { uint32_t x, y, z, w; };           // the 16-byte vector type of the device build
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v = {x, y, z, w}; return v; }
static inline uint32_t mab_sim_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint8_t by[8];
  for (int i = 0; i < 4; i++) { by[i] = (uint8_t)(a >> (8 * i)); by[4 + i] = (uint8_t)(b >> (8 * i)); }
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)by[(sel >> (4 * i)) & 7] << (8 * i);
  return r;
}
static inline uint32_t mab_bswap(uint32_t x) { return mab_sim_prmt(x, 0, 0x0123); }
static inline uint32_t mab_shf_l(uint32_t lo, uint32_t hi, unsigned n) {   // high word of (hi:lo) << n, n in [0,31]
  return (uint32_t)(((((uint64_t)hi << 32) | lo) << n) >> 32);
}
static inline uint32_t mab_shf_r(uint32_t lo, uint32_t hi, unsigned n) {   // low word of (hi:lo) >> n, n in [0,31]
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> n);
}
#else
#define MAB_DEV __device__ __forceinline__
#define MAB_NOUNROLL _Pragma("unroll 1")
static __device__ __forceinline__ uint32_t mab_bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
static __device__ __forceinline__ uint32_t mab_shf_l(uint32_t lo, uint32_t hi, unsigned n) {
  return __funnelshift_l(lo, hi, n);
}
static __device__ __forceinline__ uint32_t mab_shf_r(uint32_t lo, uint32_t hi, unsigned n) {
  return __funnelshift_r(lo, hi, n);
}
#endif
