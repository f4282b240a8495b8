// mab_kernels.cuh -- batched kernels: one field element / one key per thread.
//
// HBM layout.  Field elements are structure-of-arrays LIMB PLANES: limb j of element i is
// plane[j*stride + i] (uint32), so a warp's load of limb j is one coalesced 128-byte
// line; every kernel issues its F::L (x operands) independent plane loads up front.
// Byte strings keep the reference's array-of-structures layout (element i at
// bytes[i*Nbytes ..], modimp/modexp big-endian, rfc7748 little-endian) and are moved with
// the widest vector access the pointer alignment allows.
#pragma once
#include "mab_queue_plan.h"
#include <cuda_runtime.h>
#include "modarith_b200.h"
#include "rfc7748_sm100.cuh"
#include "weierstrass_sm100.cuh"
#include "edwards_sm100.cuh"

enum MabOp {
  OP_ADD, OP_SUB, OP_NEG, OP_MUL, OP_SQR, OP_MLI, OP_CPY, OP_NSQR, OP_PRO, OP_INV, OP_INVH,
  OP_QR, OP_QRH, OP_SQRT, OP_SQRTH, OP_IS1, OP_IS0, OP_ZER, OP_ONE, OP_INT, OP_NRES, OP_REDC,
  OP_CSW, OP_CMV, OP_SHL, OP_SHR, OP_HAF, OP_2R, OP_SIGN, OP_CMP, OP_FSB, OP_MULCHAIN
};

struct MabArgs {
  const uint32_t* a;     // first input planes
  const uint32_t* b;     // second input planes (or h)
  uint32_t* r;           // output planes (may alias a or b)
  uint32_t* r2;          // second in/out planes (modcsw)
  int* iout;             // per-element int results
  const int* iin;        // per-element int inputs (swap / move bits)
  uint32_t scalar;       // small-integer argument
  size_t n, stride;
};

template <int L> static __device__ __forceinline__ void plane_ld(uint32_t (&x)[L], const uint32_t* p, size_t stride, size_t i) {
#pragma unroll
  for (int j = 0; j < L; j++) x[j] = p[j * stride + i];
}
template <int L> static __device__ __forceinline__ void plane_st(uint32_t* p, size_t stride, size_t i, const uint32_t (&x)[L]) {
#pragma unroll
  for (int j = 0; j < L; j++) p[j * stride + i] = x[j];
}

template <class F, int OP> __global__ void __launch_bounds__(128) k_field(MabArgs p) {
  constexpr int L = F::L;
  typedef Field<F> Fd;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  uint32_t a[L], b[L], r[L];
  if constexpr (OP == OP_ADD || OP == OP_SUB || OP == OP_MUL || OP == OP_CMP || OP == OP_INVH || OP == OP_MULCHAIN ||
                OP == OP_SQRTH || OP == OP_QRH) {
    plane_ld<L>(a, p.a, p.stride, i);
    plane_ld<L>(b, p.b, p.stride, i);
  } else if constexpr (OP == OP_ZER || OP == OP_ONE || OP == OP_INT || OP == OP_2R) {
  } else if constexpr (OP == OP_CSW || OP == OP_CMV) {
    plane_ld<L>(a, p.a, p.stride, i);
    plane_ld<L>(b, p.r2, p.stride, i);
  } else {
    plane_ld<L>(a, p.a, p.stride, i);
  }
  if constexpr (OP == OP_ADD) F::add(r, a, b);
  if constexpr (OP == OP_SUB) F::sub(r, a, b);
  if constexpr (OP == OP_NEG) F::neg(r, a);
  if constexpr (OP == OP_MUL) F::mul(r, a, b);
  if constexpr (OP == OP_SQR) F::sqr(r, a);
  if constexpr (OP == OP_MULCHAIN) {        // measurement helper: r = a * b^scalar, register resident
    Fd::cpy(r, a);
    MAB_NOUNROLL
    for (uint32_t it = 0; it < p.scalar; it++) F::mul_w(r, r, b);
    if (F::WEAK) (void)F::canon(r, r);
  }
  if constexpr (OP == OP_MLI) F::mli(r, a, p.scalar);
  if constexpr (OP == OP_CPY) Fd::cpy(r, a);
  if constexpr (OP == OP_NSQR) { Fd::cpy(r, a); Fd::nsqr(r, (int)p.scalar); }
  if constexpr (OP == OP_PRO) F::pro(r, a);
  if constexpr (OP == OP_INV) Fd::template inv<false>(r, a, a);
  if constexpr (OP == OP_INVH) Fd::template inv<true>(r, a, b);
  if constexpr (OP == OP_SQRT) Fd::template sqrt<false>(r, a, a);
  if constexpr (OP == OP_SQRTH) Fd::template sqrt<true>(r, a, b);
  if constexpr (OP == OP_ZER) Fd::zer(r);
  if constexpr (OP == OP_ONE) Fd::one(r);
  if constexpr (OP == OP_INT) Fd::from_int(r, p.scalar);
  if constexpr (OP == OP_2R) Fd::pow2(r, p.scalar);
  if constexpr (OP == OP_NRES) F::nres(r, a);
  if constexpr (OP == OP_REDC) Fd::to_words(r, a);
  if constexpr (OP == OP_HAF) { Fd::cpy(r, a); Fd::haf(r); }
  if constexpr (OP == OP_SHL) { Fd::cpy(r, a); Fd::shl(r, p.scalar); }
  if constexpr (OP == OP_SHR) { Fd::cpy(r, a); uint32_t o = Fd::shr(r, p.scalar); if (p.iout) p.iout[i] = (int)o; }
  if constexpr (OP == OP_FSB) { Fd::cpy(r, a); uint32_t o = Fd::fsb(r); if (p.iout) p.iout[i] = (int)o; }
  if constexpr (OP == OP_QR) { p.iout[i] = (int)Fd::template qr<false>(a, a); return; }
  if constexpr (OP == OP_QRH) { p.iout[i] = (int)Fd::template qr<true>(a, b); return; }   // a = h, b = x
  if constexpr (OP == OP_IS1) { p.iout[i] = (int)Fd::is1(a); return; }
  if constexpr (OP == OP_IS0) { p.iout[i] = (int)Fd::is0(a); return; }
  if constexpr (OP == OP_SIGN) { p.iout[i] = (int)Fd::sign(a); return; }
  if constexpr (OP == OP_CMP) { p.iout[i] = (int)Fd::cmp(a, b); return; }
  if constexpr (OP == OP_CSW) {
    Fd::csw((uint32_t)p.iin[i] & 1u, a, b);
    plane_st<L>(p.r, p.stride, i, a);
    plane_st<L>(p.r2, p.stride, i, b);
    return;
  }
  if constexpr (OP == OP_CMV) {          // f(r2) <- g(a) iff bit
    Fd::cmv((uint32_t)p.iin[i] & 1u, a, b);
    plane_st<L>(p.r2, p.stride, i, b);
    return;
  }
  plane_st<L>(p.r, p.stride, i, r);
}

// modinv (pseudo.py:788-812) over a batch with ONE progenitor chain per K elements: Montgomery's
// simultaneous inversion inside each thread (K prefix products in registers, inputs re-read from HBM
// on the way back -- 4L more bytes per element against ~(Nbits/K) fewer squarings).  Zero inputs are
// replaced by one inside the product and give zero, as modinv(0) = 0 does; outputs may alias inputs.
template <class F, int K> __global__ void __launch_bounds__(128) k_inv_shared(MabArgs p) {
  constexpr int L = F::L;
  typedef Field<F> Fd;
  const size_t first = (size_t)blockIdx.x * K * blockDim.x + threadIdx.x;
  uint32_t P[K][L], a[L], acc[L], one[L];
  Fd::one(one);
  uint32_t flags = 0;
#pragma unroll
  for (int j = 0; j < K; j++) {
    const size_t i = first + (size_t)j * blockDim.x;
    if (i < p.n) plane_ld<L>(a, p.a, p.stride, i); else Fd::cpy(a, one);
    const uint32_t f = Fd::is0_stored(a);
    flags |= f << j;
    Fd::cmv(f, one, a);
    if (j == 0) Fd::cpy(P[0], a); else F::mul(P[j], P[j - 1], a);
  }
  {
    uint32_t h[L];
    F::pro(h, P[K - 1]);
    Fd::template inv<true>(acc, P[K - 1], h);
  }
  // The bound is re-read behind a compiler barrier: ptxas otherwise keeps the K comparisons of the first loop
  // alive across the whole inversion by packing them into the register that also holds the (secret) zero flags,
  // and branches on bits of that register -- correct, but impossible to tell from a secret-dependent branch in
  // the SASS (tools/ct_audit.py).
  size_t nn = p.n;
  asm volatile("" : "+l"(nn));
#pragma unroll
  for (int j = K - 1; j >= 0; j--) {
    const size_t i = first + (size_t)j * blockDim.x;
    uint32_t r[L];
    if (j > 0) {
      if (i < nn) plane_ld<L>(a, p.a, p.stride, i); else Fd::cpy(a, one);
      Fd::cmv((flags >> j) & 1u, one, a);
      F::mul(r, acc, P[j - 1]);
      F::mul(acc, acc, a);
    } else {
      Fd::cpy(r, acc);
    }
    uint32_t zero[L];
    Fd::zer(zero);
    Fd::cmv((flags >> j) & 1u, zero, r);
    if (i < nn) plane_st<L>(p.r, p.stride, i, r);
  }
}

// ---- modprog: a straight-line program of field operations in ONE launch ----------------------------
// The reference's consumers are sequences of field calls on a handful of elements (the complete addition of
// weierstrass.c:69-160 is 14 multiplications and 23 additions on 10 variables); one launch per call would move
// 3 x 4L bytes per element through HBM for every one of them.  Here the variables of a program live on chip:
// MAB_PROG_NREG field registers per thread in shared memory (column layout: word w of register r of thread t at
// [(r*L + w)*128 + t], conflict-free), operands are read into machine registers, the generated arithmetic runs,
// and the result goes back to shared memory.  The program is uniform across the batch (it arrives in kernel
// parameter space, i.e. the constant bank), so there is no divergence and nothing data-dependent in control
// flow or addressing.  Inputs are loaded from limb planes into registers 0 .. nin-1 before the first
// instruction; any registers can be stored to limb planes after the last.
#define MAB_PROG_NREG 16
#define MAB_PROG_MAX 320
#define MAB_PROG_THREADS 128
struct MabProg {
  const uint32_t* in[MAB_PROG_NREG];
  uint32_t* out[MAB_PROG_NREG];
  unsigned char out_reg[MAB_PROG_NREG];
  int nin, nout, ncode, nreg;      // nreg: registers the program touches (sizes the shared-memory register file)
  size_t n, stride;
  mab_insn code[MAB_PROG_MAX];
};
// vector type of one shared-memory access of the register file: 16 bytes when L is a multiple of 4, else 8 or 4
template <int L> struct MabProgVec { typedef uint32_t type; static constexpr int W = 1; };
template <> struct MabProgVec<8> { typedef uint4 type; static constexpr int W = 4; };
template <> struct MabProgVec<12> { typedef uint4 type; static constexpr int W = 4; };
template <> struct MabProgVec<16> { typedef uint4 type; static constexpr int W = 4; };
template <> struct MabProgVec<14> { typedef uint2 type; static constexpr int W = 2; };
template <> struct MabProgVec<10> { typedef uint2 type; static constexpr int W = 2; };
template <> struct MabProgVec<6> { typedef uint2 type; static constexpr int W = 2; };

template <class F> __global__ void __launch_bounds__(MAB_PROG_THREADS) k_prog(const __grid_constant__ MabProg P) {
  constexpr int L = F::L;
  constexpr int T = MAB_PROG_THREADS;
  typedef Field<F> Fd;
  typedef typename MabProgVec<L>::type V;
  constexpr int VW = MabProgVec<L>::W, NV = L / VW;       // a register is NV vectors of VW words
  static_assert(NV * VW == L, "vector width must divide the limb count");
  // chunk c of register r of thread t at rf[(r*NV + c)*T + t]: consecutive threads, consecutive vectors -- every
  // access of a warp is one conflict-free wavefront per 128 bytes
  extern __shared__ uint4 mab_smem4[];
  V* rf = reinterpret_cast<V*>(mab_smem4) + threadIdx.x;
  const size_t i = (size_t)blockIdx.x * T + threadIdx.x;
  const bool live = i < P.n;
  auto ld = [&](uint32_t (&x)[L], int r) {
#pragma unroll
    for (int c = 0; c < NV; c++) {
      const V v = rf[(r * NV + c) * T];
      const uint32_t* pv = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
      for (int w = 0; w < VW; w++) x[c * VW + w] = pv[w];
    }
  };
  auto st = [&](int r, const uint32_t (&x)[L]) {
#pragma unroll
    for (int c = 0; c < NV; c++) {
      V v;
      uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
      for (int w = 0; w < VW; w++) pv[w] = x[c * VW + w];
      rf[(r * NV + c) * T] = v;
    }
  };
  // registers 0 .. nin-1 from the input planes, the others zero (a program that reads a register it never wrote
  // sees 0, not what an earlier CTA left in shared memory)
  for (int k = 0; k < P.nreg; k++) {
    uint32_t x[L];
#pragma unroll
    for (int w = 0; w < L; w++) x[w] = (live && k < P.nin) ? P.in[k][w * P.stride + i] : 0u;
    st(k, x);
  }
  // the result of the previous instruction stays in machine registers and is forwarded to the next instruction's
  // operands (chains like t = t*u; t = t - v read one operand less from shared memory); the choice is uniform
  // across the batch, so it is a branch on the program, not on data
  uint32_t r[L];
  Fd::zer(r);
  int last = -1;
  MAB_NOUNROLL
  for (int pc = 0; pc < P.ncode; pc++) {
    const mab_insn I = P.code[pc];
    uint32_t a[L], b[L];
    const bool two = (I.op == MAB_OP_ADD || I.op == MAB_OP_SUB || I.op == MAB_OP_MUL);
    const bool one = !(I.op == MAB_OP_ZER || I.op == MAB_OP_ONE || I.op == MAB_OP_INT);
    if (one) { if ((int)I.a == last) Fd::cpy(a, r); else ld(a, I.a); }
    if (two) { if ((int)I.b == last) Fd::cpy(b, r); else ld(b, I.b); }
    switch (I.op) {
      case MAB_OP_ADD: F::add(r, a, b); break;
      case MAB_OP_SUB: F::sub(r, a, b); break;
      case MAB_OP_MUL: F::mul(r, a, b); break;
      case MAB_OP_NEG: F::neg(r, a); break;
      case MAB_OP_SQR: F::sqr(r, a); break;
      case MAB_OP_MLI: F::mli(r, a, I.imm); break;
      case MAB_OP_CPY: Fd::cpy(r, a); break;
      case MAB_OP_NSQR: Fd::cpy(r, a); Fd::nsqr(r, (int)I.imm); break;
      case MAB_OP_PRO: F::pro(r, a); break;
      case MAB_OP_INV: Fd::template inv<false>(r, a, a); break;
      case MAB_OP_SQRT: Fd::template sqrt<false>(r, a, a); break;
      case MAB_OP_ZER: Fd::zer(r); break;
      case MAB_OP_ONE: Fd::one(r); break;
      case MAB_OP_INT: Fd::from_int(r, I.imm); break;
      case MAB_OP_HAF: Fd::cpy(r, a); Fd::haf(r); break;
      default: Fd::zer(r); break;
    }
    st(I.dst, r);
    last = I.dst;
  }
  if (!live) return;
  for (int k = 0; k < P.nout; k++) {
    uint32_t x[L];
    ld(x, P.out_reg[k]);
#pragma unroll
    for (int w = 0; w < L; w++) P.out[k][w * P.stride + i] = x[w];
  }
}

// ---- byte strings ------------------------------------------------------------------
// NW words of element i from an AoS byte array (element size 4*NW bytes), widest aligned access.
template <int NW> static __device__ __forceinline__ void aos_ld(uint32_t (&w)[NW], const uint8_t* base, size_t i, unsigned align) {
  const uint8_t* e = base + i * (size_t)(4 * NW);
  if (NW % 4 == 0 && align >= 16) {
#pragma unroll
    for (int j = 0; j < NW / 4; j++) {
      uint4 v = reinterpret_cast<const uint4*>(e)[j];
      w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
    }
  } else if (NW % 2 == 0 && align >= 8) {
#pragma unroll
    for (int j = 0; j < NW / 2; j++) {
      uint2 v = reinterpret_cast<const uint2*>(e)[j];
      w[2 * j] = v.x; w[2 * j + 1] = v.y;
    }
  } else if (align >= 4) {
#pragma unroll
    for (int j = 0; j < NW; j++) w[j] = reinterpret_cast<const uint32_t*>(e)[j];
  } else {
#pragma unroll
    for (int j = 0; j < NW; j++)
      w[j] = (uint32_t)e[4 * j] | ((uint32_t)e[4 * j + 1] << 8) | ((uint32_t)e[4 * j + 2] << 16) | ((uint32_t)e[4 * j + 3] << 24);
  }
}
template <int NW> static __device__ __forceinline__ void aos_st(uint8_t* base, size_t i, unsigned align, const uint32_t (&w)[NW]) {
  uint8_t* e = base + i * (size_t)(4 * NW);
  if (NW % 4 == 0 && align >= 16) {
#pragma unroll
    for (int j = 0; j < NW / 4; j++) reinterpret_cast<uint4*>(e)[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
  } else if (NW % 2 == 0 && align >= 8) {
#pragma unroll
    for (int j = 0; j < NW / 2; j++) reinterpret_cast<uint2*>(e)[j] = make_uint2(w[2 * j], w[2 * j + 1]);
  } else if (align >= 4) {
#pragma unroll
    for (int j = 0; j < NW; j++) reinterpret_cast<uint32_t*>(e)[j] = w[j];
  } else {
#pragma unroll
    for (int j = 0; j < NW; j++) {
      e[4 * j] = (uint8_t)w[j]; e[4 * j + 1] = (uint8_t)(w[j] >> 8); e[4 * j + 2] = (uint8_t)(w[j] >> 16); e[4 * j + 3] = (uint8_t)(w[j] >> 24);
    }
  }
}

// modimp (pseudo.py:1130-1146): big-endian Nbytes -> planes; status[i] = 1 iff value < p
template <class F> __global__ void __launch_bounds__(128) k_imp(const uint8_t* bytes, uint32_t* r, int* status, size_t n, size_t stride, unsigned align) {
  constexpr int L = F::L;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[L], a[L];
  if constexpr (F::NBYTES == 4 * L) {
    uint32_t raw[L];
    aos_ld<L>(raw, bytes, i, align);
#pragma unroll
    for (int j = 0; j < L; j++) w[j] = mab_bswap(raw[L - 1 - j]);
  } else {
    // Nbytes is not a whole number of words (e.g. 521-bit moduli): byte-wise, most significant first
    const uint8_t* e = bytes + i * (size_t)F::NBYTES;
#pragma unroll
    for (int j = 0; j < L; j++) w[j] = 0;
#pragma unroll
    for (int b = 0; b < F::NBYTES; b++) {
      const int pos = F::NBYTES - 1 - b;
      w[pos >> 2] |= (uint32_t)e[b] << (8 * (pos & 3));
    }
  }
  uint32_t lt = Field<F>::from_words(a, w);
  plane_st<L>(r, stride, i, a);
  if (status) status[i] = (int)lt;
}
// modexp (pseudo.py:1115-1127): planes -> canonical big-endian Nbytes
template <class F> __global__ void __launch_bounds__(128) k_exp(const uint32_t* ap, uint8_t* bytes, size_t n, size_t stride, unsigned align) {
  constexpr int L = F::L;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[L], w[L];
  plane_ld<L>(a, ap, stride, i);
  Field<F>::to_words(w, a);
  if constexpr (F::NBYTES == 4 * L) {
    uint32_t raw[L];
#pragma unroll
    for (int j = 0; j < L; j++) raw[L - 1 - j] = mab_bswap(w[j]);
    aos_st<L>(bytes, i, align, raw);
  } else {
    uint8_t* e = bytes + i * (size_t)F::NBYTES;
#pragma unroll
    for (int b = 0; b < F::NBYTES; b++) {
      const int pos = F::NBYTES - 1 - b;
      e[b] = (uint8_t)(w[pos >> 2] >> (8 * (pos & 3)));
    }
  }
}

// rfc7748 (rfc7748.c:156): bv[i] = clamp(bk[i]) * bu[i], little-endian Nbytes strings
#ifndef MAB_LADDER_THREADS
#define MAB_LADDER_THREADS 128
#endif
// resident CTAs per SM the ladder is compiled for: chosen per modulus by the generator
// (F::LADDER_MINBLOCKS: 3 for X25519 = 146 registers in k_rfc7748_rounds, no spills) unless overridden for experiments
#ifdef MAB_LADDER_MINBLOCKS
#define MAB_LADDER_BOUNDS(F) __launch_bounds__(MAB_LADDER_THREADS, MAB_LADDER_MINBLOCKS)
#else
#define MAB_LADDER_BOUNDS(F) __launch_bounds__(MAB_LADDER_THREADS, F::LADDER_MINBLOCKS)
#endif
template <class F, bool VALIDATE = false> __global__ void MAB_LADDER_BOUNDS(F) k_rfc7748(const uint8_t* bk, const uint8_t* bu, uint8_t* bv, size_t n, unsigned align) {
  constexpr int L = F::L;
  static_assert(F::NBYTES == 4 * L, "byte strings are whole words for the supported curves");
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[L], u[L], out[L];
  aos_ld<L>(k, bk, i, align);
  aos_ld<L>(u, bu, i, align);
  if constexpr (F::LADDER_STASH) {
    // word j of this thread's column at stash[j][threadIdx.x]: consecutive threads, consecutive banks
    __shared__ uint32_t stash[2 * L][MAB_LADDER_THREADS];
    Rfc7748<F>::template scalarmult<VALIDATE>(out, k, u, &stash[0][threadIdx.x], MAB_LADDER_THREADS);
  } else {
    Rfc7748<F>::template scalarmult<VALIDATE>(out, k, u);
  }
  aos_st<L>(bv, i, align, out);
}

// rfc7748 with the inversions of up to four keys per thread shared (Rfc7748<F>::finish_batch).
// Persistent grid (one launch fills the GPU exactly); a warp repeatedly takes the next CHUNK of K
// consecutive 32-key groups and every thread runs K ladders back to back before one shared inversion.
//
// Work distribution: the G groups of the batch are cut into `nq` equal contiguous QUEUES.  With the full
// persistent grid nq = SMs x 4 and a warp serves the queue of the sub-partition it runs on (%smid, warp
// index: a 128-thread CTA puts warp w on sub-partition w), so every sub-partition is handed the same
// amount of work whatever the batch size -- the multiplier pipe of a sub-partition is the resource, and a
// batch of 1.7 rounds that lets some sub-partitions run two full rounds while others idle loses 13 %
// (round 1: 0.87 of the large-batch rate at 2^17 keys).  Inside a queue the resident warps draw chunks
// from an atomic counter: K = 4 chunks first, then K = 2, single groups last, so that the warps the
// scheduler favoured (the sub-partition arbiter is not fair) take more of them and all finish together.
// A warp whose queue is empty steals from the other queues, which also makes the result independent of
// where the CTAs were placed.  Small batches (grid below full residency) use one queue.
#ifndef MAB_LADDER_KMAX
#define MAB_LADDER_KMAX 4             // 4 or 2 (-DMAB_LADDER_KMAX=2: occupancy experiments; the host cuts no K = 4 chunks then)
#endif
struct MabQueues {
  unsigned long long* counter;     // nq counters, zeroed before the launch
  unsigned nq;                     // number of queues
  unsigned sms;                    // SMs the grid was sized for (nq == 4 * sms, or nq == 1)
  // The G groups are cut into nq contiguous queues of `base` groups, the first `rem` of them one longer.  The
  // chunk list of a queue (see mab_queue_chunks on the host side) depends on its length only, so the host works
  // it out for the two lengths that occur: index 0 = base groups, index 1 = base + 1.
  unsigned base, rem;
  unsigned c4s, c2s, c4l, c2l;     // K = 4 chunks, then K = 2 chunks (s: base groups, l: base + 1); single groups follow
};
// groups [lo, lo + g) of queue q; `longer` = 1 for the queues that hold base + 1 groups
static __device__ __forceinline__ void mab_queue_range(const MabQueues& Q, unsigned q, unsigned long long& lo, unsigned& g, unsigned& longer) {
  longer = q < Q.rem ? 1u : 0u;
  lo = (unsigned long long)q * Q.base + (q < Q.rem ? q : Q.rem);
  g = Q.base + longer;
}
// chunk ci of a queue -> K and the first group: mab_queue_chunk (mab_queue_plan.h, shared with the host and the CPU tests)
template <class F> __global__ void MAB_LADDER_BOUNDS(F) k_rfc7748_rounds(const uint8_t* bk, const uint8_t* bu, uint8_t* bv, size_t n, unsigned align, MabQueues Q) {
  constexpr int L = F::L;
  constexpr int T = MAB_LADDER_THREADS;
  extern __shared__ uint32_t mab_smem[];
  uint32_t* st = mab_smem + threadIdx.x;                               // K slots x 3 elements x L words
  uint32_t* stash = F::LADDER_STASH ? (mab_smem + MAB_LADDER_KMAX * 3 * L * T + threadIdx.x) : nullptr;
  const unsigned lane = threadIdx.x & 31;
  // The only scheduling state that lives across a ladder is `swept`: AT_HOME while the warp serves the queue of
  // its own sub-partition, afterwards the number of queues behind the home queue already seen empty (counters
  // only grow, so they stay empty).  The other queues are gone through 32 at a time, every lane peeking at one
  // counter: a sweep over all SMs x 4 queues is a handful of round trips (a lane-0-only scan at the end of the
  // kernel cost 7 % of a 2^17-key batch).
  constexpr unsigned AT_HOME = 0xffffffffu;
  unsigned swept = AT_HOME;
  for (;;) {
    unsigned home = 0;
    if (Q.nq > 1) {
      unsigned smid;
      asm("mov.u32 %0, %%smid;" : "=r"(smid));
      while (smid >= Q.sms) smid -= Q.sms;            // sparse SM numbering: fold (any mapping is correct)
      home = smid * 4 + ((threadIdx.x >> 5) & 3);
    }
    int K = 0;
    unsigned first = 0;
    unsigned long long glo = 0;
    for (;;) {
      unsigned q = home;
      if (swept != AT_HOME) {
        // the next queue behind home + swept that still holds chunks
        bool found = false;
        while (swept + 1 < Q.nq) {
          const unsigned cand = swept + 1 + lane;                       // offset from home, 1 .. nq-1
          int has = 0;
          if (cand < Q.nq) {
            unsigned qq = home + cand;
            if (qq >= Q.nq) qq -= Q.nq;
            unsigned long long lo2; unsigned g2, lg2;
            mab_queue_range(Q, qq, lo2, g2, lg2);
            const unsigned c4 = lg2 ? Q.c4l : Q.c4s, c2 = lg2 ? Q.c2l : Q.c2s;
            has = *(volatile unsigned long long*)(Q.counter + qq) < (unsigned long long)mab_queue_nchunks(g2, c4, c2);
          }
          const unsigned m = __ballot_sync(0xffffffffu, has);
          if (m) {
            swept += (unsigned)(__ffs((int)m) - 1);                     // everything before it is empty for good
            q = home + swept + 1;
            if (q >= Q.nq) q -= Q.nq;
            found = true;
            break;
          }
          swept += 32;
        }
        if (!found) break;
      }
      unsigned g, longer;
      mab_queue_range(Q, q, glo, g, longer);
      const unsigned c4 = longer ? Q.c4l : Q.c4s, c2 = longer ? Q.c2l : Q.c2s;
      const unsigned nchunks = mab_queue_nchunks(g, c4, c2);
      unsigned long long ci = 0;
      int ok = 0;
      if (lane == 0) {
        // a plain read first: an exhausted queue costs a load, not an atomic
        if (*(volatile unsigned long long*)(Q.counter + q) < (unsigned long long)nchunks) {
          ci = atomicAdd(Q.counter + q, 1ULL);
          ok = 1;
        }
      }
      ok = __shfl_sync(0xffffffffu, ok, 0);
      if (ok) {
        ci = __shfl_sync(0xffffffffu, ci, 0);
        K = mab_queue_chunk(g, c4, c2, ci, first);
        if (K) break;
      }
      // this queue is empty: the home queue is left for good; a stolen-from queue is swept past
      swept = (swept == AT_HOME) ? 0 : swept + 1;
      if (Q.nq == 1) break;
    }
    if (!K) break;
    const size_t start = (size_t)(glo + first) * 32;
    MAB_NOUNROLL
    for (int j = 0; j < K; j++) {
      const size_t idx = start + (size_t)j * 32 + lane;
      uint32_t k[L], u[L], x1[L], x2[L], z2[L];
      if (idx < n) {
        aos_ld<L>(k, bk, idx, align);
        aos_ld<L>(u, bu, idx, align);
      } else {
#pragma unroll
        for (int w = 0; w < L; w++) { k[w] = 0; u[w] = 0; }
      }
      Rfc7748<F>::ladder(x2, z2, x1, k, u, stash, T);
      Rfc7748<F>::st_(st, T, j, 0, x2);
      Rfc7748<F>::st_(st, T, j, 1, z2);
    }
    Rfc7748<F>::finish_batch(st, T, K);
    MAB_NOUNROLL
    for (int j = 0; j < K; j++) {
      const size_t idx = start + (size_t)j * 32 + lane;
      if (idx < n) {
        uint32_t out[L];
        Rfc7748<F>::ld(out, st, T, j, 0);
        aos_st<L>(bv, idx, align, out);
      }
    }
  }
}

// ecnXXXset + ecnXXXmul + ecnXXXget (weierstrass.c:415-427,494-542,333-349) for n independent points:
// (xo, yo) = affine(e * (x, y)); all strings big-endian Nbytes as the reference's char* arguments;
// a point that is not on the curve, a zero scalar or a multiple of the group order give (0, 1).
// Where the fixed-window table lives is a property of the group (G::ECN_GLOBAL_TABLE, G::ECN_MINBLOCKS;
// -DMAB_ECN_GLOBAL_TABLE / -DMAB_ECN_MINBLOCKS override both groups for experiments).  The table is
// 9 entries x 3 coordinates x L words = 864 B per point for L = 8:
//   shared memory   two 128-thread CTAs per SM at most; measured best for Ed25519, whose digit step is
//                   short enough that the lookups matter (36.1 vs 31.0 M/s);
//   global memory   each resident CTA owns one slice of a workspace that only resident CTAs touch
//                   (148 x 3 x 108 KB = 48 MB: L2 resident, re-used block after block); shared memory
//                   then only keeps the scalar and its recoding carries, and a third CTA fits --
//                   measured best for P-256, whose long reduction chains need the extra warps (11.3 vs
//                   9.6 M/s).
#ifdef MAB_ECN_GLOBAL_TABLE
#define MAB_ECN_GLOBAL(G) (MAB_ECN_GLOBAL_TABLE != 0)
#else
#define MAB_ECN_GLOBAL(G) (G::ECN_GLOBAL_TABLE)
#endif
#ifdef MAB_ECN_MINBLOCKS
#define MAB_ECN_MB(G) (MAB_ECN_MINBLOCKS)
#else
#define MAB_ECN_MB(G) (G::ECN_MINBLOCKS)
#endif
#ifdef MAB_ECN2_MINBLOCKS
#define MAB_ECN2_MB(G) (MAB_ECN2_MINBLOCKS)
#else
#define MAB_ECN2_MB(G) (G::ECN2_MINBLOCKS)
#endif
#define MAB_ECN_THREADS 128
// Persistent grid: CTA b handles the 128-point blocks b, b + gridDim.x, ...
template <class F, class G> __global__ void __launch_bounds__(MAB_ECN_THREADS, MAB_ECN_MB(G))
k_ecnmul(const uint8_t* e, const uint8_t* x, const uint8_t* y, uint8_t* xo, uint8_t* yo, size_t n, unsigned align,
         uint4* tabws) {
  constexpr int L = F::L;
  static_assert(F::NBYTES == 4 * L, "whole-word byte strings");
  extern __shared__ uint4 mab_smem4[];
  typedef EcnMul<G, MAB_ECN_THREADS> M;
  uint4* tab;
  uint32_t* scr;
  if (MAB_ECN_GLOBAL(G)) {
    tab = tabws + (size_t)blockIdx.x * (9 * 3 * (L / 4) * MAB_ECN_THREADS) + threadIdx.x;
    scr = reinterpret_cast<uint32_t*>(mab_smem4) + threadIdx.x;
  } else {
    tab = mab_smem4 + threadIdx.x;
    scr = nullptr;
  }
  for (size_t blk = blockIdx.x; blk * MAB_ECN_THREADS < n; blk += gridDim.x) {
    const size_t i = blk * MAB_ECN_THREADS + threadIdx.x;
    if (i >= n) break;
    uint32_t raw[L], ew[L], xw[L], yw[L];
    aos_ld<L>(raw, e, i, align);
#pragma unroll
    for (int j = 0; j < L; j++) ew[j] = mab_bswap(raw[L - 1 - j]);
    aos_ld<L>(raw, x, i, align);
#pragma unroll
    for (int j = 0; j < L; j++) xw[j] = mab_bswap(raw[L - 1 - j]);
    aos_ld<L>(raw, y, i, align);
#pragma unroll
    for (int j = 0; j < L; j++) yw[j] = mab_bswap(raw[L - 1 - j]);
    typename G::Pt P;
    G::set(P, xw, yw);
    M::mul(P, ew, tab, MAB_ECN_THREADS, scr, align >> 16);   // align <= 16: a zero ptxas cannot fold
    G::get(xw, yw, P);
#pragma unroll
    for (int j = 0; j < L; j++) raw[L - 1 - j] = mab_bswap(xw[j]);
    aos_st<L>(xo, i, align, raw);
#pragma unroll
    for (int j = 0; j < L; j++) raw[L - 1 - j] = mab_bswap(yw[j]);
    aos_st<L>(yo, i, align, raw);
  }
}

// ecnXXXset x2 + ecnXXXmul2 + ecnXXXget (weierstrass.c:545-572 / edwards.c:486-513) for n independent
// pairs: (xo, yo) = affine(e*(x1,y1) + f*(x2,y2)); same conventions and the same persistent grid /
// table placement as k_ecnmul (the table has five entries here, the scratch column 4(L+1) words).
template <class F, class G> __global__ void __launch_bounds__(MAB_ECN_THREADS, MAB_ECN2_MB(G))
k_ecnmul2(const uint8_t* e, const uint8_t* x1, const uint8_t* y1, const uint8_t* f, const uint8_t* x2, const uint8_t* y2,
          uint8_t* xo, uint8_t* yo, size_t n, unsigned align, uint4* tabws) {
  constexpr int L = F::L;
  extern __shared__ uint4 mab_smem4[];
  typedef EcnMul<G, MAB_ECN_THREADS> M;
  uint4* tab;
  uint32_t* scr;
  if (G::MUL2_WINDOW) {              // joint 2-bit windows: 16-entry table, always in the global workspace
    tab = tabws + (size_t)blockIdx.x * (M::NE2W * 3 * (L / 4) * MAB_ECN_THREADS) + threadIdx.x;
    scr = reinterpret_cast<uint32_t*>(mab_smem4) + threadIdx.x;
  } else if (MAB_ECN_GLOBAL(G)) {
    tab = tabws + (size_t)blockIdx.x * (9 * 3 * (L / 4) * MAB_ECN_THREADS) + threadIdx.x;
    scr = reinterpret_cast<uint32_t*>(mab_smem4) + threadIdx.x;
  } else {
    tab = mab_smem4 + threadIdx.x;
    scr = reinterpret_cast<uint32_t*>(mab_smem4 + 5 * 3 * (L / 4) * MAB_ECN_THREADS) + threadIdx.x;
  }
  for (size_t blk = blockIdx.x; blk * MAB_ECN_THREADS < n; blk += gridDim.x) {
    const size_t i = blk * MAB_ECN_THREADS + threadIdx.x;
    if (i >= n) break;
    uint32_t raw[L], ew[L], fw[L], xw[L], yw[L];
    typename G::Pt P, Q, R;
    aos_ld<L>(raw, x1, i, align);
#pragma unroll
    for (int j = 0; j < L; j++) xw[j] = mab_bswap(raw[L - 1 - j]);
    aos_ld<L>(raw, y1, i, align);
#pragma unroll
    for (int j = 0; j < L; j++) yw[j] = mab_bswap(raw[L - 1 - j]);
    G::set(P, xw, yw);
    aos_ld<L>(raw, x2, i, align);
#pragma unroll
    for (int j = 0; j < L; j++) xw[j] = mab_bswap(raw[L - 1 - j]);
    aos_ld<L>(raw, y2, i, align);
#pragma unroll
    for (int j = 0; j < L; j++) yw[j] = mab_bswap(raw[L - 1 - j]);
    G::set(Q, xw, yw);
    aos_ld<L>(raw, e, i, align);
#pragma unroll
    for (int j = 0; j < L; j++) ew[j] = mab_bswap(raw[L - 1 - j]);
    aos_ld<L>(raw, f, i, align);
#pragma unroll
    for (int j = 0; j < L; j++) fw[j] = mab_bswap(raw[L - 1 - j]);
    if (G::MUL2_WINDOW) M::mul2w(R, ew, P, fw, Q, tab, MAB_ECN_THREADS, scr, align >> 16);
    else M::mul2(R, ew, P, fw, Q, tab, MAB_ECN_THREADS, scr, align >> 16);
    G::get(xw, yw, R);
#pragma unroll
    for (int j = 0; j < L; j++) raw[L - 1 - j] = mab_bswap(xw[j]);
    aos_st<L>(xo, i, align, raw);
#pragma unroll
    for (int j = 0; j < L; j++) raw[L - 1 - j] = mab_bswap(yw[j]);
    aos_st<L>(yo, i, align, raw);
  }
}

