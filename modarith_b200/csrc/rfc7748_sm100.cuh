// rfc7748_sm100.cuh -- RFC 7748 Montgomery ladder, one key per thread, state in registers.
//
// Counterpart of rfc7748.c:156-256 (and of the <<<2,1>>> demo kernel
// simd/rfc7748_simt.cu:155-226, whose 13 field elements live in a 432-byte local-memory
// frame behind 522 calls).  Here the step keeps 4 persistent elements (x2,z2,x3,z3) plus
// x1 and 4 temporaries, all in registers, every field call inlined into straight-line
// IMAD.WIDE chains; the scalar is consumed by shifting it left one bit per step so no
// register is indexed dynamically; the two conditional swaps of the reference's step are
// replaced by a masked choice of the two operands of its doubling half (see the step).
#pragma once
#include "mab_field.cuh"

// experiment switch: -DMAB_STEP_UNROLL2 unrolls the step loop twice (lets ptxas overlap the ALU-only head of a step with
// the multiplications that end the step before it)
#ifdef MAB_STEP_UNROLL2
#define MAB_STEP_UNROLL _Pragma("unroll 2")
#else
#define MAB_STEP_UNROLL MAB_NOUNROLL
#endif
// experiment switch: -DMAB_NO_TAIL_DOUBLING runs all Nbits steps through the full ladder step
#ifdef MAB_NO_TAIL_DOUBLING
#define MAB_TAIL_DOUBLINGS(F) 0
#else
#define MAB_TAIL_DOUBLINGS(F) (F::COF)
#endif

template <class F> struct Rfc7748 {
  static constexpr int L = F::L;
  typedef Field<F> Fd;

  // clamp (rfc7748.c:135-141) on little-endian words
  static MAB_DEV void clamp(uint32_t (&k)[L]) {
    constexpr int s = (8 - (F::NBITS % 8)) % 8;
    k[0] &= ~((1u << F::COF) - 1u);
    constexpr uint32_t topmask = ((0xffu >> s) << 24) | 0x00ffffffu;
    k[L - 1] &= topmask;
    k[L - 1] |= (0x80u >> s) << 24;
  }

  // One scalar multiplication.  k, u: little-endian byte strings as words; out likewise.
  // `stash`: optional per-thread column of 2*L words (element j at stash[j*pitch]) in shared memory.
  // When given (F::LADDER_STASH), the scalar and x1 -- read once per 32 steps / once per step -- live
  // there instead of in 2*L registers, which buys the 14-limb X448 ladder a third warp per
  // sub-partition; nullptr keeps everything in registers.
  // VALIDATE = false: the TWIST_SECURE tail (rfc7748.c:225-227); true: the cheap point validation
  // of the #else branch (rfc7748.c:228-251, eprint 2020/1497) -- the result is forced to zero
  // when u is not the x-coordinate of a point on the curve.
  template <bool VALIDATE = false>
  static MAB_DEV void scalarmult(uint32_t (&out)[L], uint32_t (&k)[L], uint32_t (&u)[L],
                                 uint32_t* stash = nullptr, int pitch = 0) {
    uint32_t x1[L], x2[L], z2[L];
    ladder(x2, z2, x1, k, u, stash, pitch);
    tail<VALIDATE>(out, x2, z2, x1, stash, pitch);
  }

  // clamp, import, and the Nbits ladder steps (rfc7748.c:166-223): leaves the projective result
  // (x2 : z2) and x1 = u in internal form (x1 also in the stash when one is given).
  static MAB_DEV void ladder(uint32_t (&x2)[L], uint32_t (&z2)[L], uint32_t (&x1)[L], uint32_t (&k)[L],
                             uint32_t (&u)[L], uint32_t* stash = nullptr, int pitch = 0) {
    // mask() (rfc7748.c:148-152,172): drop the bits above Nbits in the top byte of u
    constexpr int rbits = (F::NBITS % 8) ? (F::NBITS % 8) : 8;
    u[L - 1] &= ((((1u << rbits) - 1u) << 24) | 0x00ffffffu);
    clamp(k);

    uint32_t x3[L], z3[L];
    (void)Fd::from_words(x1, u);                 // modimp (rfc7748.c:178)
    Fd::one(x2);
    Fd::zer(z2);
    Fd::cpy(x3, x1);
    Fd::one(z3);

    // The scalar is consumed from the top: the current word sits in k[L-1] and is shifted left one
    // bit per step; after its last bit the words rotate up by one.  Two ALU instructions per step and
    // no dynamically indexed register (bit(), rfc7748.c:144-146, indexes a byte array instead).
    constexpr int topbits = F::NBITS - 32 * (L - 1);        // bits of the scalar held by the top word
    if (topbits < 32) k[L - 1] <<= (32 - topbits);
    if (stash) {
#pragma unroll
      for (int j = 0; j < L; j++) { stash[j * pitch] = k[j]; stash[(L + j) * pitch] = x1[j]; }
    }

    // Every addition and subtraction of the step takes two products (or the initial 1, 0, u < 2^Nbits):
    // where the field keeps products below 2^Nbits + c*2^13 (F::TIGHT) the _tt forms apply.
    uint32_t swap = 0;
    MAB_NOUNROLL
    for (int w = L - 1; w >= 0; w--) {           // rfc7748.c:186-221, bits Nbits-1 .. 0
      uint32_t kw = stash ? stash[w * pitch] : k[L - 1];
      // the lowest COF bits of a clamped scalar are zero (rfc7748.c:137): those steps are plain
      // doublings of (x2:z2) and are done after the loop without the differential-addition half
      const int nb = (w == L - 1) ? topbits : (w == 0 ? 32 - MAB_TAIL_DOUBLINGS(F) : 32);
      MAB_STEP_UNROLL
      for (int bi = 0; bi < nb; bi++) {
        uint32_t kt = kw >> 31;
        kw <<= 1;
        swap ^= kt;
        // The two cswaps of rfc7748.c:192-193 exchange A<->C and B<->D below.  DA and CB only
        // trade places under that exchange and are consumed symmetrically ((DA+CB)^2,
        // (DA-CB)^2), so they are formed from the unswapped values and only the operands of
        // the doubling half, AA = (swap ? C : A)^2 and BB = (swap ? D : B)^2, are selected:
        // 2L word selects per step instead of 4L word swaps.
        const uint32_t m = 0u - swap;
        swap = kt;

        uint32_t A[L], B[L], C[L], D[L], AA[L], BB[L];
        F::add_tt(A, x2, z2);                         // A = x2+z2
        F::sub_tt(B, x2, z2);                         // B = x2-z2
        F::add_tt(C, x3, z3);                         // C = x3+z3
        F::sub_tt(D, x3, z3);                         // D = x3-z3
#pragma unroll
        for (int j = 0; j < L; j++) AA[j] = (A[j] & ~m) | (C[j] & m);
#pragma unroll
        for (int j = 0; j < L; j++) BB[j] = (B[j] & ~m) | (D[j] & m);
        F::mul(D, D, A);                           // DA (or CB of the swapped frame)
        F::mul(C, C, B);                           // CB (or DA)
        F::sqr(A, AA);                             // AA
        F::sqr(B, BB);                             // BB
        F::add_tt(x3, D, C);
        F::sub_tt(z3, D, C);
        F::sqr(x3, x3);                            // x3 = (DA+CB)^2
        F::sqr(z3, z3);
        if (stash) {
          uint32_t t1[L];
#pragma unroll
          for (int j = 0; j < L; j++) t1[j] = stash[(L + j) * pitch];
          F::mul(z3, z3, t1);                      // z3 = x1*(DA-CB)^2
        } else {
          F::mul(z3, z3, x1);
        }
        F::mul(x2, A, B);                          // x2 = AA*BB
        F::sub_tt(B, A, B);                           // E = AA-BB
        F::mla(z2, B, F::A24, A);                  // a24*E + AA  (modmli + modadd fused)
        F::mul(z2, z2, B);                         // z2 = E*(AA+a24*E)
      }
      if (!stash) {
#pragma unroll
        for (int j = L - 1; j > 0; j--) k[j] = k[j - 1];
      }
    }
    Fd::csw(swap, x2, x3);
    Fd::csw(swap, z2, z3);
#pragma unroll 1
    for (int i = 0; i < MAB_TAIL_DOUBLINGS(F); i++) {   // bits COF-1..0 are 0: x2,z2 <- double(x2,z2)
      uint32_t A[L], B[L];
      F::add_tt(A, x2, z2);
      F::sub_tt(B, x2, z2);
      F::sqr(A, A);                              // AA
      F::sqr(B, B);                              // BB
      F::mul(x2, A, B);                          // x2 = AA*BB
      F::sub_tt(B, A, B);                           // E
      F::mla(z2, B, F::A24, A);
      F::mul(z2, z2, B);                         // z2 = E*(AA+a24*E)
    }
  }

  // the part after the ladder (rfc7748.c:225-255): x2/z2 (0 -> 0), canonical little-endian words
  template <bool VALIDATE>
  static MAB_DEV void tail(uint32_t (&out)[L], uint32_t (&x2)[L], uint32_t (&z2)[L], const uint32_t (&x1)[L],
                           uint32_t* stash, int pitch) {
    if (!VALIDATE) {
      // TWIST_SECURE branch (rfc7748.c:225-227,252): x2/z2 with 0 -> 0
      uint32_t h[L];
      F::pro(h, z2);
      Fd::template inv<true>(z2, z2, h);
    } else {
      // rfc7748.c:228-251: one progenitor gives both 1/z2 and the quadratic character of the
      // curve equation at u; D ends as 1 (valid point) or 0 and multiplies the result
      uint32_t A[L], B[L], C[L], D[L], E[L], w[L];
      if (stash) {
#pragma unroll
        for (int j = 0; j < L; j++) w[j] = stash[(L + j) * pitch];
      } else {
        Fd::cpy(w, x1);
      }
      F::mul(B, w, z2);                          // wZ
      F::mul(A, B, z2);                          // wZ^2
      F::pro(E, A);                              // y
      Fd::cpy(C, A);
      F::mul(D, E, z2);                          // y.Z2
      F::sqr(D, D);
      F::mul(D, D, w);                           // w.(y.z2)^2
#pragma unroll
      for (int i = 0; i < F::COF - 2; i++) { F::sqr(C, C); F::mul(C, C, A); }
#pragma unroll
      for (int i = 0; i < F::COF; i++) F::sqr(E, E);
      F::mul(C, C, E);
      F::mul(z2, C, B);
#pragma unroll
      for (int i = 0; i < F::COF - 2; i++) F::sqr(D, D);
      Fd::one(A);
      F::add(D, D, A);
      (void)Fd::fsb(D);
      (void)Fd::shr(D, 1);                       // 1 for QR, else 0
      F::mul(x2, x2, D);                         // zero for a bad input point
    }
    F::mul(x2, x2, z2);
    Fd::to_words(out, x2);                       // modexp (rfc7748.c:254)
  }

  // ---- one inversion for K keys (Montgomery's simultaneous-inversion trick) ---------------------------
  // A thread that has run K ladders back to back finishes them together: the K progenitor chains of
  // rfc7748.c:226-227 (each ~Nbits squarings, a tenth of a scalar multiplication) become ONE chain plus
  // 3(K-1) multiplications.  Results are identical: x2_j * z2_j^-1 is what every key gets, and a key
  // whose z2 is zero (low-order input) still gives zero as modinv(0)=0 does (pseudo.py:788-812): its z2
  // is replaced by 1 inside the product and its result forced to zero, all branch-free.
  // st: this thread's column of K slots x 3 elements x L words; element e of slot j, word w at
  // st[((j*3+e)*L + w)*pitch].  In: e0 = x2, e1 = z2.  Out: e0 = canonical result words.
  static MAB_DEV void finish_batch(uint32_t* st, int pitch, int K) {
    uint32_t acc[L], z[L], t[L], one[L];
    Fd::one(one);
    uint32_t flags = 0;
    MAB_NOUNROLL
    for (int j = 0; j < K; j++) {
      ld(z, st, pitch, j, 1);
      uint32_t f = Fd::is0_stored(z);
      flags |= f << j;
      Fd::cmv(f, one, z);                        // z' = 1 where z2 == 0
      st_(st, pitch, j, 1, z);
      if (j == 0) Fd::cpy(acc, z); else F::mul(acc, acc, z);
      st_(st, pitch, j, 2, acc);                 // prefix product P_j
    }
    {
      uint32_t h[L];
      F::pro(h, acc);
      Fd::template inv<true>(acc, acc, h);       // 1 / (z'_0 ... z'_{K-1})
    }
    MAB_NOUNROLL
    for (int j = K - 1; j >= 0; j--) {
      if (j > 0) {
        ld(t, st, pitch, j - 1, 2);
        F::mul(t, acc, t);                       // 1 / z'_j
        ld(z, st, pitch, j, 1);
        F::mul(acc, acc, z);                     // 1 / (z'_0 ... z'_{j-1})
      } else {
        Fd::cpy(t, acc);
      }
      ld(z, st, pitch, j, 0);
      F::mul(z, z, t);                           // x2_j / z2_j
      uint32_t zero[L];
      Fd::zer(zero);
      Fd::cmv((flags >> j) & 1u, zero, z);
      Fd::to_words(t, z);
      st_(st, pitch, j, 0, t);
    }
  }
  static MAB_DEV void ld(uint32_t (&r)[L], const uint32_t* st, int pitch, int j, int e) {
#pragma unroll
    for (int w = 0; w < L; w++) r[w] = st[((j * 3 + e) * L + w) * pitch];
  }
  static MAB_DEV void st_(uint32_t* st, int pitch, int j, int e, const uint32_t (&r)[L]) {
#pragma unroll
    for (int w = 0; w < L; w++) st[((j * 3 + e) * L + w) * pitch] = r[w];
  }
};
