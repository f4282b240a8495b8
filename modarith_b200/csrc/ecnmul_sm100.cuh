// ecnmul_sm100.cuh -- the reference's constant-time fixed-window scalar multiplication ecnXXXmul
// (weierstrass.c:441-542 and, identically, edwards.c:382-484) over any group G that provides
// Pt{x,y,z}, inf, cpy, neg, add (P <- P+Q) and dbl: table O,P,..,8P in shared memory, nibbles recoded
// to signed digits in [-8,7], every lookup a masked scan of all nine entries, 4 doublings + 1 addition
// per digit.  One point per thread; nothing depends on the scalar except through masks.
#pragma once
#include "mab_field.cuh"

template <class G> struct EcnMul {
  static constexpr int L = G::L;
  typedef typename G::Pt Pt;

  // tab: this thread's column of the table W[0..8]: coordinate c, word w of entry e at
  // tab[((e*3+c)*L + w)*pitch]
  static MAB_DEV void tab_st(uint32_t* tab, int pitch, int e, const Pt& P) {
#pragma unroll
    for (int w = 0; w < L; w++) {
      tab[((e * 3 + 0) * L + w) * pitch] = P.x[w];
      tab[((e * 3 + 1) * L + w) * pitch] = P.y[w];
      tab[((e * 3 + 2) * L + w) * pitch] = P.z[w];
    }
  }
  // constant-time lookup of digit d in [-8,8]: scan all nine entries, then negate if d < 0
  static MAB_DEV void select(Pt& R, const uint32_t* tab, int pitch, int d) {
    const int m = d >> 31;
    const uint32_t dabs = (uint32_t)((d ^ m) - m);
#pragma unroll
    for (int w = 0; w < L; w++) { R.x[w] = 0; R.y[w] = 0; R.z[w] = 0; }
    MAB_NOUNROLL
    for (uint32_t e = 0; e < 9; e++) {
      const uint32_t mask = 0u - (uint32_t)(e == dabs);
#pragma unroll
      for (int w = 0; w < L; w++) {
        R.x[w] |= tab[((e * 3 + 0) * L + w) * pitch] & mask;
        R.y[w] |= tab[((e * 3 + 1) * L + w) * pitch] & mask;
        R.z[w] |= tab[((e * 3 + 2) * L + w) * pitch] & mask;
      }
    }
    Pt N;
    G::cpy(N, R);
    G::neg(N);
    G::cmv((uint32_t)m & 1u, N, R);
  }

  // P <- e*P; e = plain scalar as little-endian words (the reference takes Nbytes big-endian bytes)
  static MAB_DEV void mul(Pt& P, const uint32_t (&e)[L], uint32_t* tab, int pitch) {
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
    select(P, tab, pitch, (int)c);                // top digit = final carry
    MAB_NOUNROLL
    for (int j = ND - 1; j >= 0; j--) {
      const int w = j >> 3, n = j & 7;
      uint32_t ew = 0, cwd = 0, cprev = 0;
#pragma unroll
      for (int q = 0; q < L; q++) {               // pick word w with masks: no dynamically indexed register
        const uint32_t mk = 0u - (uint32_t)(q == w);
        ew |= e[q] & mk;
        cwd |= carries[q] & mk;
        if (q > 0) cprev |= carries[q - 1] & mk;
      }
      const uint32_t cin = (n == 0) ? ((w == 0) ? 0u : (cprev >> 7) & 1u) : ((cwd >> (n - 1)) & 1u);
      const uint32_t cout = (cwd >> n) & 1u;
      const int d = (int)(((ew >> (4 * n)) & 0xfu) + cin) - (int)(cout << 4);
      G::dbl(P); G::dbl(P); G::dbl(P); G::dbl(P);
      select(Q, tab, pitch, d);       // after the doublings: Q's 3L registers are not live across them
      G::add(P, Q);
    }
  }
};
