// ecnmul_sm100.cuh -- the reference's constant-time fixed-window scalar multiplication ecnXXXmul
// (weierstrass.c:441-542 and, identically, edwards.c:382-484) over any group G that provides
// Pt{x,y,z}, inf, cpy, neg, add (P <- P+Q) and dbl: table O,P,..,8P in shared memory or in a global
// workspace slice (the kernel decides per group), nibbles recoded
// to signed digits in [-8,7], every lookup a masked scan of all nine entries, 4 doublings + 1 addition
// per digit.  One point per thread; nothing depends on the scalar except through masks.
#pragma once
#include "mab_field.cuh"

// PITCH > 0 fixes the stride between consecutive words of one thread's table / scratch column at compile
// time (every access becomes base + immediate); PITCH == 0 takes it from the `pitch` argument.
template <class G, int PITCH = 0> struct EcnMul {
  static constexpr int L = G::L;
  typedef typename G::Pt Pt;
  static MAB_DEV int stride(int pitch) { return PITCH ? PITCH : pitch; }

  // tab: this thread's column of the table W[0..8] in 16-byte chunks: words 4q..4q+3 of coordinate c
  // of entry e form chunk ((e*3+c)*L/4 + q), stored at tab[chunk*pitch] -- one vector load / store per
  // chunk, and a warp's 32 chunks are one contiguous 512-byte run (no bank conflicts, full lines).
  static_assert(L % 4 == 0, "table chunks are four words");
  static constexpr int CH = L / 4;                 // chunks per coordinate
  static MAB_DEV void chunk_st(uint4* tab, int sp, int chunk, const uint32_t* v) {
    tab[(size_t)chunk * sp] = make_uint4(v[0], v[1], v[2], v[3]);
  }
  static MAB_DEV void tab_st(uint4* tab, int pitch, int e, const Pt& P) {
    const int sp = stride(pitch);
#pragma unroll
    for (int q = 0; q < CH; q++) {
      chunk_st(tab, sp, (e * 3 + 0) * CH + q, P.x + 4 * q);
      chunk_st(tab, sp, (e * 3 + 1) * CH + q, P.y + 4 * q);
      chunk_st(tab, sp, (e * 3 + 2) * CH + q, P.z + 4 * q);
    }
  }
  // constant-time lookup of digit d in [-8,8]: every one of the nine entries is loaded and merged under
  // a mask (one LOP3 per word; the mask is made opaque so that the compiler can neither turn the scan
  // into a branch around the loads nor into a select plus an OR), then the result is negated if d < 0
  static MAB_DEV void pick(uint32_t* r, const uint4 v, uint32_t mask) {
    r[0] |= v.x & mask;  r[1] |= v.y & mask;  r[2] |= v.z & mask;  r[3] |= v.w & mask;
  }
  template <int NE = 9> static MAB_DEV void select(Pt& R, const uint4* tab, int pitch, int d) {
    const int sp = stride(pitch);
    const int m = d >> 31;
    const uint32_t dabs = (uint32_t)((d ^ m) - m);
#pragma unroll
    for (int w = 0; w < L; w++) { R.x[w] = 0; R.y[w] = 0; R.z[w] = 0; }
    MAB_NOUNROLL
    for (uint32_t e = 0; e < NE; e++) {
      uint32_t hit = 0u - (uint32_t)(e == dabs);
#ifndef MAB_HOSTSIM
      asm volatile("" : "+r"(hit));
#endif
      const uint4* te = tab + (size_t)e * (3 * CH) * sp;
#pragma unroll
      for (int q = 0; q < CH; q++) {
        pick(R.x + 4 * q, te[(size_t)(0 * CH + q) * sp], hit);
        pick(R.y + 4 * q, te[(size_t)(1 * CH + q) * sp], hit);
        pick(R.z + 4 * q, te[(size_t)(2 * CH + q) * sp], hit);
      }
    }
    Pt N;
    G::cpy(N, R);
    G::neg(N);
    G::cmv((uint32_t)m & 1u, N, R);
  }

  // P <- e*P; e = plain scalar as little-endian words (the reference takes Nbytes big-endian bytes).
  // scr, when given, is a 2L-word column (same stride as the table) that holds the scalar and the
  // recoding carries during the loop instead of 2L registers; it may be indexed dynamically.
  // z: a zero the compiler cannot see, handed to the group law (Weierstrass::Seq); 0 when not needed.
  static MAB_DEV void mul(Pt& P, const uint32_t (&e)[L], uint4* tab, int pitch, uint32_t* scr = nullptr, uint32_t z = 0) {
    const int sp = stride(pitch);
    {
      Pt Q, T;
      G::inf(Q);                        tab_st(tab, pitch, 0, Q);
      tab_st(tab, pitch, 1, P);
      G::cpy(Q, P); G::dbl(Q);          tab_st(tab, pitch, 2, Q);     // 2P
      G::cpy(T, Q); G::add(T, P);       tab_st(tab, pitch, 3, T);     // 3P
      G::dbl(Q);                        tab_st(tab, pitch, 4, Q);     // 4P
      { Pt U; G::cpy(U, Q); G::add(U, P); tab_st(tab, pitch, 5, U); } // 5P
      G::dbl(T);                        tab_st(tab, pitch, 6, T);     // 6P
      G::add(T, P);                     tab_st(tab, pitch, 7, T);     // 7P
      G::dbl(Q);                        tab_st(tab, pitch, 8, Q);     // 8P
    }

    // signed digits (weierstrass.c:513-526): digit j = nibble j + carry_in - 16*carry_out with
    // carry_out = (nibble + carry_in > 7).  The loop runs from the top digit down, so the carries are
    // produced first, bottom up, one bit per nibble.
    constexpr int ND = 8 * L;
    uint32_t carries[L];
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < L; w++) {
      uint32_t cw = 0;
#pragma unroll
      for (int n = 0; n < 8; n++) {
        const uint32_t v = ((e[w] >> (4 * n)) & 0xfu) + c;
        c = (v > 7u) ? 1u : 0u;
        cw |= c << n;
      }
      carries[w] = cw;
    }
    if (scr) {
#pragma unroll
      for (int w = 0; w < L; w++) { scr[w * sp] = e[w]; scr[(L + w) * sp] = carries[w]; }
    }
    select(P, tab, pitch, (int)c);                // top digit = final carry
    typename G::Seq q = G::seq(z);
    MAB_NOUNROLL
    for (int j = ND - 1; j >= 0; j--) {
      const int w = j >> 3, n = j & 7;
      uint32_t ew = 0, cwd = 0, cprev = 0;
      if (scr) {
        ew = scr[w * sp];
        cwd = scr[(L + w) * sp];
        cprev = scr[(L + (w > 0 ? w - 1 : 0)) * sp];
      } else {
#pragma unroll
        for (int q = 0; q < L; q++) {             // pick word w with masks: no dynamically indexed register
          const uint32_t mk = 0u - (uint32_t)(q == w);
          ew |= e[q] & mk;
          cwd |= carries[q] & mk;
          if (q > 0) cprev |= carries[q - 1] & mk;
        }
      }
      const uint32_t cin = (n == 0) ? ((w == 0) ? 0u : (cprev >> 7) & 1u) : ((cwd >> (n - 1)) & 1u);
      const uint32_t cout = (cwd >> n) & 1u;
      const int d = (int)(((ew >> (4 * n)) & 0xfu) + cin) - (int)(cout << 4);
      G::dbl4(P, q);                  // P <- 16 P: four doublings, in whatever form the group does them fastest
      Pt Q;
      select(Q, tab, pitch, d);       // after the doublings: Q's 3L registers are not live across them
      G::add(P, Q, q);
    }
  }

  // R <- e*P + f*Q  (ecnXXXmul2, weierstrass.c:545-572 / edwards.c:486-513; ECDSA / EdDSA verification).
  // Joint signed digits w_i = (bit_i(3e) - bit_i(e)) + 3 (bit_i(3f) - bit_i(f)) in [-4,4] (dnaf,
  // weierstrass.c:463-492), table O, P, Q-P, Q, Q+P, digits i = 8*Nbytes+7 .. 1 from the top: one
  // doubling, then +W[w_i] or -W[-w_i].  The reference skips leading and zero digits (it is variable
  // time by design: the inputs of a verification are public); a warp would execute every branch anyway,
  // so here every digit does the doubling and one addition of a masked-scan-selected entry (O for a zero
  // digit) -- same result, constant work per point, no divergence.
  // scr: 4(L+1)-word column: the +1 and -1 digit positions of e and of f as bit masks.
  static constexpr int SCR2 = 4 * (L + 1);
  static MAB_DEV void naf_masks(uint32_t* plus, uint32_t* minus, int sp, const uint32_t (&e)[L]) {
    uint32_t c = 0, prev = 0;
#pragma unroll
    for (int w = 0; w <= L; w++) {
      const uint32_t ew = (w < L) ? e[w] : 0u;
      const uint32_t two = (ew << 1) | (prev >> 31);             // word w of 2e
      prev = ew;
      const uint64_t t = (uint64_t)ew + two + c;                 // word w of 3e
      c = (uint32_t)(t >> 32);
      const uint32_t t3 = (uint32_t)t;
      plus[w * sp] = t3 & ~ew;
      minus[w * sp] = ew & ~t3;
    }
  }
  static MAB_DEV void mul2(Pt& R, const uint32_t (&e)[L], const Pt& P, const uint32_t (&f)[L], const Pt& Q, uint4* tab,
                           int pitch, uint32_t* scr, uint32_t z = 0) {
    const int sp = stride(pitch);
    {
      Pt T, N;
      G::inf(T);                                     tab_st(tab, pitch, 0, T);     // O
      tab_st(tab, pitch, 1, P);                                                    // P
      tab_st(tab, pitch, 3, Q);                                                    // Q
      G::cpy(N, P); G::neg(N); G::cpy(T, Q); G::add(T, N);  tab_st(tab, pitch, 2, T);     // Q-P
      G::cpy(T, Q); G::add(T, P);                    tab_st(tab, pitch, 4, T);     // Q+P
    }
    uint32_t* pe = scr;
    uint32_t* me = scr + (size_t)(L + 1) * sp;
    uint32_t* pf = scr + (size_t)2 * (L + 1) * sp;
    uint32_t* mf = scr + (size_t)3 * (L + 1) * sp;
    naf_masks(pe, me, sp, e);
    naf_masks(pf, mf, sp, f);
    G::inf(R);
    typename G::Seq q = G::seq(z);
    MAB_NOUNROLL
    for (int i = 32 * L + 7; i >= 1; i--) {
      const int w = i >> 5, b = i & 31;
      const int de = (int)((pe[w * sp] >> b) & 1u) - (int)((me[w * sp] >> b) & 1u);
      const int df = (int)((pf[w * sp] >> b) & 1u) - (int)((mf[w * sp] >> b) & 1u);
      G::dbl(R, q);
      Pt T;
      select<5>(T, tab, pitch, de + 3 * df);
      G::add(R, T, q);
    }
  }

  // R <- e*P + f*Q with joint 2-bit windows: table T[a + 4b] = a*P + b*Q for a, b in 0..3 (2 doublings and 13
  // additions to build), then for each of the 16L digit pairs from the top: R <- 4R + T[e_i + 4 f_i].  256 doublings
  // and 128 + 13 additions for 256-bit scalars where the one-bit joint-digit form above (the reference's schedule,
  // minus its variable-time skipping) does 263 + 263: the same point, a quarter fewer products.  Every digit does
  // the same work (masked scan of all sixteen entries).  scr: 2L-word column holding e and f.
  static constexpr int NE2W = 16;
  static constexpr int SCR2W = 2 * L;
  static MAB_DEV void mul2w(Pt& R, const uint32_t (&e)[L], const Pt& P, const uint32_t (&f)[L], const Pt& Q, uint4* tab,
                            int pitch, uint32_t* scr, uint32_t z = 0) {
    const int sp = stride(pitch);
    typename G::Seq q = G::seq(z);
    {
      Pt T, U;
      G::inf(T);                                   tab_st(tab, pitch, 0, T);
      tab_st(tab, pitch, 1, P);
      G::cpy(T, P); G::dbl(T, q);                  tab_st(tab, pitch, 2, T);       // 2P
      G::add(T, P, q);                             tab_st(tab, pitch, 3, T);       // 3P
      tab_st(tab, pitch, 4, Q);
      G::cpy(U, Q); G::dbl(U, q);                  tab_st(tab, pitch, 8, U);       // 2Q
      G::add(U, Q, q);                             tab_st(tab, pitch, 12, U);      // 3Q
      MAB_NOUNROLL
      for (int b = 1; b < 4; b++) {
        MAB_NOUNROLL
        for (int a = 1; a < 4; a++) {
          tab_ld_dyn(T, tab, sp, a);               // loop counters: public indices, plain loads
          tab_ld_dyn(U, tab, sp, 4 * b);
          G::add(T, U, q);
          tab_st_dyn(tab, sp, a + 4 * b, T);
        }
      }
    }
#pragma unroll
    for (int w = 0; w < L; w++) { scr[w * sp] = e[w]; scr[(L + w) * sp] = f[w]; }
    G::inf(R);
    MAB_NOUNROLL
    for (int i = 16 * L - 1; i >= 0; i--) {
      const int w = i >> 4, sh = (i & 15) * 2;
      const int d = (int)((scr[w * sp] >> sh) & 3u) + 4 * (int)((scr[(L + w) * sp] >> sh) & 3u);
      G::template dbln<2>(R, q);
      Pt T;
      select<NE2W>(T, tab, pitch, d);
      G::add(R, T, q);
    }
  }
  static MAB_DEV void tab_ld_dyn(Pt& P, const uint4* tab, int sp, int e) {
#pragma unroll
    for (int qd = 0; qd < CH; qd++) {
      const uint4 vx = tab[(size_t)((e * 3 + 0) * CH + qd) * sp], vy = tab[(size_t)((e * 3 + 1) * CH + qd) * sp],
                  vz = tab[(size_t)((e * 3 + 2) * CH + qd) * sp];
      P.x[4 * qd] = vx.x; P.x[4 * qd + 1] = vx.y; P.x[4 * qd + 2] = vx.z; P.x[4 * qd + 3] = vx.w;
      P.y[4 * qd] = vy.x; P.y[4 * qd + 1] = vy.y; P.y[4 * qd + 2] = vy.z; P.y[4 * qd + 3] = vy.w;
      P.z[4 * qd] = vz.x; P.z[4 * qd + 1] = vz.y; P.z[4 * qd + 2] = vz.z; P.z[4 * qd + 3] = vz.w;
    }
  }
  // table store at an entry index that is not a compile-time constant (same layout as tab_st)
  static MAB_DEV void tab_st_dyn(uint4* tab, int sp, int e, const Pt& P) {
#pragma unroll
    for (int qd = 0; qd < CH; qd++) {
      chunk_st(tab, sp, (e * 3 + 0) * CH + qd, P.x + 4 * qd);
      chunk_st(tab, sp, (e * 3 + 1) * CH + qd, P.y + 4 * qd);
      chunk_st(tab, sp, (e * 3 + 2) * CH + qd, P.z + 4 * qd);
    }
  }
};
