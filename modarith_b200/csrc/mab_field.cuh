// mab_field.cuh -- the prime-independent half of the generated-code API, as templates
// over a generated field struct F (gen/emit.py).  One thread = one field element held in
// F::L registers; every function is branch-free in the data ("constant time": no
// secret-dependent branch or address, cf. pseudo.py:979-1048 and README.md:104-108).
//
// Function-by-function counterpart of the reference's generated C (file:line in each
// comment).  Argument ORDER follows this library's convention (destination first); the
// C ABI in mab_capi.cu restores the reference's order.
#pragma once
#include "mab_common.cuh"

template <class F> struct Field {
  static constexpr int L = F::L;
  typedef uint32_t fe[F::L];

  static MAB_DEV void cpy(uint32_t (&r)[L], const uint32_t (&a)[L]) {          // modcpy, pseudo.py:730-743
#pragma unroll
    for (int i = 0; i < L; i++) r[i] = a[i];
  }
  static MAB_DEV void zer(uint32_t (&r)[L]) {                                  // modzer, pseudo.py:909-919
#pragma unroll
    for (int i = 0; i < L; i++) r[i] = 0;
  }
  static MAB_DEV void one(uint32_t (&r)[L]) { F::set_one(r); }                 // modone, pseudo.py:922-934
  static MAB_DEV void from_int(uint32_t (&r)[L], uint32_t x) {                 // modint, pseudo.py:937-949
    uint32_t t[L];
    zer(t);
    t[0] = x;
    F::nres(r, t);
  }
  static MAB_DEV void nsqr(uint32_t (&a)[L], int n) {                          // modnsqr, pseudo.py:745-755
    MAB_NOUNROLL
    for (int i = 0; i < n; i++) F::sqr_w(a, a);      // weakly reduced inside the chain where the plan has such a form
    if (F::WEAK) (void)F::canon(a, a);
  }
  // modfsb, pseudo.py:272-283: canonicalise in place, return 1 iff the stored value was < p
  static MAB_DEV uint32_t fsb(uint32_t (&a)[L]) { return F::canon(a, a); }

  // modinv, pseudo.py:788-812.  h = progenitor x^PE if HAS_H, else computed here.  0 -> 0.
  template <bool HAS_H>
  static MAB_DEV void inv(uint32_t (&z)[L], const uint32_t (&x)[L], const uint32_t (&h)[L]) {
    uint32_t s[L], t[L];
    if (HAS_H) cpy(t, h); else F::pro(t, x);
    cpy(s, x);
#pragma unroll
    for (int i = 0; i < F::PM1D2 - 1; i++) { F::sqr(s, s); F::mul(s, s, x); }
    nsqr(t, F::PM1D2 + 1);
    F::mul(z, s, t);
  }

  // modis1 / modis0, pseudo.py:877-906 (on the canonical plain value)
  static MAB_DEV uint32_t is1(const uint32_t (&a)[L]) {
    uint32_t c[L];
    F::redc(c, a);
    (void)F::canon(c, c);
    uint32_t d = c[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < L; i++) d |= c[i];
    return d == 0 ? 1u : 0u;
  }
  static MAB_DEV uint32_t is0(const uint32_t (&a)[L]) {
    uint32_t c[L];
    F::redc(c, a);
    (void)F::canon(c, c);
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < L; i++) d |= c[i];
    return d == 0 ? 1u : 0u;
  }

  // zero test on the STORED value (no redc: zero is zero in Montgomery form too); used where the
  // value is only needed as a flag (shared inversions)
  static MAB_DEV uint32_t is0_stored(const uint32_t (&a)[L]) {
    uint32_t c[L];
    (void)F::canon(c, a);
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < L; i++) d |= c[i];
    return d == 0 ? 1u : 0u;
  }

  // modcsw / modcmv, mask form of pseudo.py:1006-1013,1041-1047 (PSCR=False): b in {0,1}
  static MAB_DEV void csw(uint32_t b, uint32_t (&g)[L], uint32_t (&f)[L]) {
    uint32_t m = 0u - b;
#pragma unroll
    for (int i = 0; i < L; i++) {
      uint32_t gi = g[i], fi = f[i];
      g[i] = (gi & ~m) | (fi & m);
      f[i] = (fi & ~m) | (gi & m);
    }
  }
  static MAB_DEV void cmv(uint32_t b, const uint32_t (&g)[L], uint32_t (&f)[L]) {   // f <- g iff b
    uint32_t m = 0u - b;
#pragma unroll
    for (int i = 0; i < L; i++) f[i] = (f[i] & ~m) | (g[i] & m);
  }

  // modqr, pseudo.py:815-831 (argument order of the reference: h first)
  template <bool HAS_H>
  static MAB_DEV uint32_t qr(const uint32_t (&h)[L], const uint32_t (&x)[L]) {
    uint32_t r[L];
    if (HAS_H) F::sqr(r, h); else { F::pro(r, x); F::sqr(r, r); }
    F::mul(r, r, x);
    if (F::PM1D2 > 1) nsqr(r, F::PM1D2 - 1);
    return is1(r) | is0(x);
  }

  // modsqrt, pseudo.py:834-874: x*x^PE when p = 3 mod 4, constant-time Tonelli-Shanks otherwise
  template <bool HAS_H>
  static MAB_DEV void sqrt(uint32_t (&r)[L], const uint32_t (&x)[L], const uint32_t (&h)[L]) {
    uint32_t s[L], y[L];
    if (HAS_H) cpy(y, h); else F::pro(y, x);
    F::mul(s, y, x);
    if (F::PM1D2 > 1) {
      uint32_t t[L], b[L], v[L], z[L];
      F::mul(t, s, y);
      F::set_roi(z);
      MAB_NOUNROLL
      for (int k = F::PM1D2; k > 1; k--) {
        cpy(b, t);
        nsqr(b, k - 2);
        uint32_t d = 1u - is1(b);
        F::mul(v, s, z);
        cmv(d, v, s);
        F::sqr(z, z);
        F::mul(v, t, z);
        cmv(d, v, t);
      }
    }
    cpy(r, s);
  }

  // modhaf, pseudo.py:1084-1100: a/2 mod p
  static MAB_DEV void haf(uint32_t (&a)[L]) {
    uint32_t c[L], pw[L];
    (void)F::canon(c, a);
    F::set_p(pw);
    uint32_t m = 0u - (c[0] & 1u);          // odd: add p first (sum < 2^(32L+1), keep the carry)
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
      uint64_t t = (uint64_t)c[i] + (pw[i] & m) + carry;
      c[i] = (uint32_t)t;
      carry = (uint32_t)(t >> 32);
    }
#pragma unroll
    for (int i = 0; i < L - 1; i++) a[i] = mab_shf_r(c[i], c[i + 1], 1);
    a[L - 1] = mab_shf_r(c[L - 1], carry, 1);
  }

  // a/2 mod p for a value that is already in [0, p) (what the fully reduced plans store): no canonicalisation first
  static MAB_DEV void haf_reduced(uint32_t (&a)[L]) {
    uint32_t pw[L];
    F::set_p(pw);
    const uint32_t m = 0u - (a[0] & 1u);
    uint32_t c[L], carry = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
      uint64_t t = (uint64_t)a[i] + (pw[i] & m) + carry;
      c[i] = (uint32_t)t;
      carry = (uint32_t)(t >> 32);
    }
#pragma unroll
    for (int i = 0; i < L - 1; i++) a[i] = mab_shf_r(c[i], c[i + 1], 1);
    a[L - 1] = mab_shf_r(c[L - 1], carry, 1);
  }

  // modshl, pseudo.py:1052-1065.  The reference shifts raw limbs (no reduction; it has spare
  // bits).  Saturated limbs have none, so this is a * 2^n as a field element -- identical
  // wherever the reference's result is a legal (non-overflowing) element.
  static MAB_DEV void shl(uint32_t (&a)[L], unsigned n) {
    MAB_NOUNROLL
    for (unsigned i = 0; i < n; i++) F::add(a, a, a);
  }
  // modshr, pseudo.py:1068-1081: floor(v / 2^n) of the CANONICAL stored value v, returning the
  // n low bits (n < 32).  The reference shifts whatever representative its limbs hold; the two
  // agree on canonical input, which is the only case its own callers use (modexp, modhaf).
  static MAB_DEV uint32_t shr(uint32_t (&a)[L], unsigned n) {
    uint32_t c[L];
    (void)F::canon(c, a);
    uint32_t out = c[0] & ((1u << n) - 1u);
#pragma unroll
    for (int i = 0; i < L - 1; i++) a[i] = mab_shf_r(c[i], c[i + 1], n);
    a[L - 1] = c[L - 1] >> n;
    return out;
  }
  // mod2r, pseudo.py:1102-1112 / monty.py:1568-1577: a = 2^r (0 when r >= 8*Nbytes)
  static MAB_DEV void pow2(uint32_t (&a)[L], unsigned r) {
    uint32_t t[L];
#pragma unroll
    for (int i = 0; i < L; i++) t[i] = (r < 8u * F::NBYTES && (r >> 5) == (unsigned)i) ? (1u << (r & 31)) : 0u;
    import_raw(a, t);
  }
  // raw words (any value below 2^(32L)) -> stored form; returns 1 iff the value was < p.  Plain-residue plans
  // canonicalise (their canon covers the whole raw range by construction).  Montgomery plans multiply the RAW
  // words by R^2: (w*R2 + m*p)/R < 2p for any w < R, so the product's own final subtraction is enough whatever
  // the ratio 2^(32L)/p is -- moduli with spare bits in their top word included (canon alone, one conditional
  // subtraction, would be short there); canon still supplies the w < p flag.
  static MAB_DEV uint32_t import_raw(uint32_t (&a)[L], const uint32_t (&w)[L]) {
    uint32_t t[L];
    const uint32_t lt = F::canon(t, w);
    if (F::MONTGOMERY) F::nres(a, w); else F::nres(a, t);
    return lt;
  }

  // modexp, pseudo.py:1115-1127: canonical plain value as words, least significant first
  static MAB_DEV void to_words(uint32_t (&w)[L], const uint32_t (&a)[L]) {
    F::redc(w, a);
    (void)F::canon(w, w);
  }
  // modimp, pseudo.py:1130-1146: raw words (value < 2^(32L)) -> stored form; returns 1 iff < p
  static MAB_DEV uint32_t from_words(uint32_t (&a)[L], const uint32_t (&w)[L]) { return import_raw(a, w); }
  // modsign, pseudo.py:1149-1158; modcmp, pseudo.py:1161-1174
  static MAB_DEV uint32_t sign(const uint32_t (&a)[L]) {
    uint32_t c[L];
    to_words(c, a);
    return c[0] & 1u;
  }
  static MAB_DEV uint32_t cmp(const uint32_t (&a)[L], const uint32_t (&b)[L]) {
    uint32_t c[L], d[L];
    to_words(c, a);
    to_words(d, b);
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < L; i++) x |= c[i] ^ d[i];
    return x == 0 ? 1u : 0u;
  }
};
