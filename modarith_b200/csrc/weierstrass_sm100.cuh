// weierstrass_sm100.cuh -- short-Weierstrass group law (A = -3) and constant-time scalar multiplication
// on the batched field, one point per thread.  SURVEY.md section 8(f) row 1: the direct consumer of the
// P-256 field work.
//
// Counterpart of weierstrass.c: complete projective addition and doubling from eprint 2015/1060
// (Renes, Costello, Batina; Algorithms 4 and 6 for a = -3, which weierstrass.c:69-176,187-282 also
// transcribes), point set/validate (weierstrass.c:364-427), affine/get (:297-349) and the fixed-window
// signed-digit multiplication ecnXXXmul (:441-542): table O,P,..,8P, nibbles recoded to digits in
// [-8,7], every lookup a scan of the whole table with masks.  Results are compared as affine (x,y) byte
// strings, which are unique, so only the values -- not the order of internal operations -- are pinned.
#pragma once
#include "mab_field.cuh"

template <class F> struct Weierstrass {
  static constexpr int L = F::L;
  typedef Field<F> Fd;
  struct Pt { uint32_t x[L], y[L], z[L]; };

  static MAB_DEV void inf(Pt& P) { Fd::zer(P.x); Fd::one(P.y); Fd::zer(P.z); }          // weierstrass.c:283-288
  static MAB_DEV void cpy(Pt& R, const Pt& P) { Fd::cpy(R.x, P.x); Fd::cpy(R.y, P.y); Fd::cpy(R.z, P.z); }
  static MAB_DEV void cmv(uint32_t d, const Pt& Q, Pt& P) { Fd::cmv(d, Q.x, P.x); Fd::cmv(d, Q.y, P.y); Fd::cmv(d, Q.z, P.z); }

  // P <- P + Q, complete (eprint 2015/1060 Algorithm 4, a = -3): 12M + 2 mul-by-b + 29 add/sub
  static MAB_DEV void add(Pt& P, const Pt& Q) {
    uint32_t b[L], t0[L], t1[L], t2[L], t3[L], t4[L], x3[L], y3[L], z3[L];
    F::set_b(b);
    F::mul(t0, P.x, Q.x);  F::mul(t1, P.y, Q.y);  F::mul(t2, P.z, Q.z);
    F::add(t3, P.x, P.y);  F::add(t4, Q.x, Q.y);  F::mul(t3, t3, t4);
    F::add(t4, t0, t1);    F::sub(t3, t3, t4);    F::add(t4, P.y, P.z);
    F::add(x3, Q.y, Q.z);  F::mul(t4, t4, x3);    F::add(x3, t1, t2);
    F::sub(t4, t4, x3);    F::add(x3, P.x, P.z);  F::add(y3, Q.x, Q.z);
    F::mul(x3, x3, y3);    F::add(y3, t0, t2);    F::sub(y3, x3, y3);
    F::mul(z3, b, t2);     F::sub(x3, y3, z3);    F::add(z3, x3, x3);
    F::add(x3, x3, z3);    F::sub(z3, t1, x3);    F::add(x3, t1, x3);
    F::mul(y3, b, y3);     F::add(t1, t2, t2);    F::add(t2, t1, t2);
    F::sub(y3, y3, t2);    F::sub(y3, y3, t0);    F::add(t1, y3, y3);
    F::add(y3, t1, y3);    F::add(t1, t0, t0);    F::add(t0, t1, t0);
    F::sub(t0, t0, t2);    F::mul(t1, t4, y3);    F::mul(t2, t0, y3);
    F::mul(y3, x3, z3);    F::add(y3, y3, t2);    F::mul(x3, t3, x3);
    F::sub(x3, x3, t1);    F::mul(z3, t4, z3);    F::mul(t1, t3, t0);
    F::add(z3, z3, t1);
    Fd::cpy(P.x, x3); Fd::cpy(P.y, y3); Fd::cpy(P.z, z3);
  }

  // P <- 2P, complete (Algorithm 6, a = -3): 8M + 3S + 2 mul-by-b + 21 add/sub
  static MAB_DEV void dbl(Pt& P) {
    uint32_t b[L], t0[L], t1[L], t2[L], t3[L], x3[L], y3[L], z3[L];
    F::set_b(b);
    F::sqr(t0, P.x);       F::sqr(t1, P.y);       F::sqr(t2, P.z);
    F::mul(t3, P.x, P.y);  F::add(t3, t3, t3);    F::mul(z3, P.x, P.z);
    F::add(z3, z3, z3);    F::mul(y3, b, t2);     F::sub(y3, y3, z3);
    F::add(x3, y3, y3);    F::add(y3, x3, y3);    F::sub(x3, t1, y3);
    F::add(y3, t1, y3);    F::mul(y3, x3, y3);    F::mul(x3, x3, t3);
    F::add(t3, t2, t2);    F::add(t2, t2, t3);    F::mul(z3, b, z3);
    F::sub(z3, z3, t2);    F::sub(z3, z3, t0);    F::add(t3, z3, z3);
    F::add(z3, z3, t3);    F::add(t3, t0, t0);    F::add(t0, t3, t0);
    F::sub(t0, t0, t2);    F::mul(t0, t0, z3);    F::add(y3, y3, t0);
    F::mul(t0, P.y, P.z);  F::add(t0, t0, t0);    F::mul(z3, t0, z3);
    F::sub(x3, x3, z3);    F::mul(z3, t0, t1);    F::add(z3, z3, z3);
    F::add(z3, z3, z3);
    Fd::cpy(P.x, x3); Fd::cpy(P.y, y3); Fd::cpy(P.z, z3);
  }

  // ecnXXXset with both coordinates (weierstrass.c:364-396,415-427): (x,y) if y^2 = x^3 - 3x + b, else O.
  // xw, yw: plain values as little-endian words (any value < 2^(32L): modimp semantics)
  static MAB_DEV void set(Pt& P, const uint32_t (&xw)[L], const uint32_t (&yw)[L]) {
    uint32_t v[L], t[L], b[L];
    (void)Fd::from_words(P.x, xw);
    (void)Fd::from_words(P.y, yw);
    F::sqr(v, P.x);
    F::mul(v, v, P.x);
    F::sub(v, v, P.x); F::sub(v, v, P.x); F::sub(v, v, P.x);
    F::set_b(b);
    F::add(v, v, b);
    F::sqr(t, P.y);
    const uint32_t bad = 1u - Fd::cmp(t, v);
    Fd::one(P.z);
    Pt O;
    inf(O);
    cmv(bad, O, P);
  }

  // ecnXXXget (weierstrass.c:297-349): affine coordinates as canonical plain words; O -> (0, 1)
  static MAB_DEV void get(uint32_t (&xw)[L], uint32_t (&yw)[L], const Pt& P) {
    uint32_t i[L], x[L], y[L], one[L];
    Fd::template inv<false>(i, P.z, P.z);          // 0 -> 0
    F::mul(x, P.x, i);
    F::mul(y, P.y, i);
    Fd::one(one);
    Fd::cmv(Fd::is0_stored(P.z), one, y);
    Fd::to_words(xw, x);
    Fd::to_words(yw, y);
  }

  // ---- ecnXXXmul (weierstrass.c:441-542) -------------------------------------------------------------
  // tab: this thread's column of the table W[0..8] = O,P,..,8P in shared memory: coordinate c, word w of
  // entry e at tab[((e*3+c)*L + w)*pitch]
  static MAB_DEV void tab_st(uint32_t* tab, int pitch, int e, const Pt& P) {
#pragma unroll
    for (int w = 0; w < L; w++) {
      tab[((e * 3 + 0) * L + w) * pitch] = P.x[w];
      tab[((e * 3 + 1) * L + w) * pitch] = P.y[w];
      tab[((e * 3 + 2) * L + w) * pitch] = P.z[w];
    }
  }
  // constant-time lookup of digit d in [-8,8]: scan all nine entries, then negate y if d < 0
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
    uint32_t ny[L];
    F::neg(ny, R.y);
    Fd::cmv((uint32_t)m & 1u, ny, R.y);
  }

  // P <- e*P; e = plain scalar as little-endian words (the reference takes Nbytes big-endian bytes)
  static MAB_DEV void mul(Pt& P, const uint32_t (&e)[L], uint32_t* tab, int pitch) {
    Pt Q;
    inf(Q);          tab_st(tab, pitch, 0, Q);
    tab_st(tab, pitch, 1, P);
    cpy(Q, P); dbl(Q);            tab_st(tab, pitch, 2, Q);     // 2P
    Pt T;
    cpy(T, Q); add(T, P);         tab_st(tab, pitch, 3, T);     // 3P
    dbl(Q);                       tab_st(tab, pitch, 4, Q);     // 4P
    { Pt U; cpy(U, Q); add(U, P); tab_st(tab, pitch, 5, U); }   // 5P
    dbl(T);                       tab_st(tab, pitch, 6, T);     // 6P
    add(T, P);                    tab_st(tab, pitch, 7, T);     // 7P
    dbl(Q);                       tab_st(tab, pitch, 8, Q);     // 8P

    // signed digits: nibble j plus the carry of nibble j-1, minus 16 when it exceeds 7 (weierstrass.c:513-526);
    // processed from the top, so the carries are produced by a first pass from the bottom
    constexpr int ND = 8 * L;                   // nibbles
    uint32_t carries[L];                        // bit j of word j/32... one carry bit per nibble
#pragma unroll
    for (int w = 0; w < L; w++) carries[w] = 0;
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < L; w++) {
      uint32_t cw = 0;
#pragma unroll
      for (int n = 0; n < 8; n++) {
        const uint32_t v = ((e[w] >> (4 * n)) & 0xfu) + c;      // 0..16
        c = (v > 7u) ? 1u : 0u;
        cw |= c << n;
      }
      carries[w] = cw;                           // carry OUT of each nibble of this word
    }
    // top digit = final carry
    select(P, tab, pitch, (int)c);
    MAB_NOUNROLL
    for (int j = ND - 1; j >= 0; j--) {
      const int w = j >> 3, n = j & 7;
      // digit j = nibble + carry_in - 16*carry_out
      uint32_t ew = 0, cwd = 0, cprev = 0;
#pragma unroll
      for (int q = 0; q < L; q++) {              // static indexing only: pick word w with masks
        const uint32_t mk = 0u - (uint32_t)(q == w);
        ew |= e[q] & mk;
        cwd |= carries[q] & mk;
        if (q > 0) cprev |= carries[q - 1] & (0u - (uint32_t)(q == w));
      }
      const uint32_t cin = (n == 0) ? ((w == 0) ? 0u : (cprev >> 7) & 1u) : ((cwd >> (n - 1)) & 1u);
      const uint32_t cout = (cwd >> n) & 1u;
      const int d = (int)(((ew >> (4 * n)) & 0xfu) + cin) - (int)(cout << 4);
      select(Q, tab, pitch, d);
      dbl(P); dbl(P); dbl(P); dbl(P);
      add(P, Q);
    }
  }
};
