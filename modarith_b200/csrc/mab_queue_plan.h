// Chunk plan of one work queue of the round-structured ladder (mab_kernels.cuh: k_rfc7748_rounds): plain C++ shared by
// the host (mab_capi.inc cuts the plan), the device (the kernel walks it) and the CPU test suite
// (tests/test_queue_plan.py compiles this header with g++ and checks that every plan covers every group exactly once).
#pragma once
#if defined(__CUDACC__)
#define MAB_HD __host__ __device__ __forceinline__
#else
#define MAB_HD inline
#endif

struct MabQueuePlan {
  unsigned c4, c2;                 // K = 4 chunks first, then K = 2 chunks; single groups follow
};

// A queue of g groups served by w warps: whole rounds of w chunks at K = 4 (every warp gets the same number of them),
// then whole rounds at K = 2, and at least one round's worth of single groups last, which the warps that ran ahead
// take more of.  Partial rounds of big chunks are what to avoid: a warp that draws one more K = 4 chunk than its
// neighbours finishes four key-times late.  (Measured at 2^20 keys in round 1: handing the remainder out two groups at
// a time instead of one costs 2.9 %.)
// A batch too short for K = 4 chunks (c4 == 0: at most a handful of groups per warp) has nothing for the reserved
// round of single groups to balance, and every single group pays a whole inversion: there every whole round of K = 2
// chunks the remainder holds is cut (2^17 keys: 48.7 -> 50.4 M/s X25519; 2^16 keys X448: 8.55 -> 8.95 M/s; longer
// batches unchanged -- profiles/r2_tail_rule.txt).  tail = 0 keeps the reserve always, 1 is the rule just described
// (default), 2 drops the reserve always (loses 3 % on X448 at 2^19 keys).  kmax < 4 cuts no K = 4 chunks.
MAB_HD MabQueuePlan mab_queue_plan(unsigned g, unsigned w, int kmax, int tail) {
  MabQueuePlan p;
  p.c4 = (kmax >= 4 && g > w) ? (g - w) / (4 * w) * w : 0;
  const unsigned left = g - 4 * p.c4;
  p.c2 = (left > w) ? (left - w) / (2 * w) * w : 0;
  if (tail == 2 || (tail == 1 && p.c4 == 0)) p.c2 = left / (2 * w) * w;
  return p;
}

// number of chunks of a queue of g groups under the plan (c4, c2)
MAB_HD unsigned mab_queue_nchunks(unsigned g, unsigned c4, unsigned c2) { return c4 + c2 + (g - 4 * c4 - 2 * c2); }

// chunk ci of a queue: returns K (0 = past the end) and the first group of the chunk inside the queue
MAB_HD int mab_queue_chunk(unsigned g, unsigned c4, unsigned c2, unsigned long long ci, unsigned& first) {
  if (ci < c4) { first = (unsigned)ci * 4; return 4; }
  if (ci < (unsigned long long)c4 + c2) { first = c4 * 4 + (unsigned)(ci - c4) * 2; return 2; }
  const unsigned long long s = (unsigned long long)c4 * 4 + (unsigned long long)c2 * 2 + (ci - c4 - c2);
  if (s >= g) return 0;
  first = (unsigned)s;
  return 1;
}
