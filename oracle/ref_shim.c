/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Appended (by oracle/build_ref.py) after the reference's generated field.c and,
 * for the Montgomery curves, the reference's rfc7748.c, inside one scratch
 * translation unit.  The generated functions are `static` (pseudo.py:1861,
 * 1910-1961), so this file re-exports them under a ref_ prefix with the exact
 * signatures of SURVEY.md section 8(a), plus a multi-threaded batch loop over
 * rfc7748() (rfc7748.c:156) that bench.py times as the CPU baseline.
 * Nothing here is product code and no reference source is copied into the repo.
 */
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REF_EXPORT __attribute__((visibility("default")))

REF_EXPORT int ref_wordlength(void) { return Wordlength; }
REF_EXPORT int ref_nlimbs(void) { return Nlimbs; }
REF_EXPORT int ref_radix(void) { return Radix; }
REF_EXPORT int ref_nbits(void) { return Nbits; }
REF_EXPORT int ref_nbytes(void) { return Nbytes; }

REF_EXPORT int ref_modfsb(spint *n) { return (int)modfsb(n); }
REF_EXPORT void ref_modadd(const spint *a, const spint *b, spint *n) { modadd(a, b, n); }
REF_EXPORT void ref_modsub(const spint *a, const spint *b, spint *n) { modsub(a, b, n); }
REF_EXPORT void ref_modneg(const spint *b, spint *n) { modneg(b, n); }
REF_EXPORT void ref_modmul(const spint *a, const spint *b, spint *c) { modmul(a, b, c); }
REF_EXPORT void ref_modsqr(const spint *a, spint *c) { modsqr(a, c); }
REF_EXPORT void ref_modmli(const spint *a, int b, spint *c) { modmli(a, b, c); }
REF_EXPORT void ref_modcpy(const spint *a, spint *c) { modcpy(a, c); }
REF_EXPORT void ref_modnsqr(spint *a, int n) { modnsqr(a, n); }
REF_EXPORT void ref_modpro(const spint *w, spint *z) { modpro(w, z); }
REF_EXPORT void ref_modinv(const spint *x, const spint *h, spint *z) { modinv(x, h, z); }
REF_EXPORT int ref_modqr(const spint *h, const spint *x) { return modqr(h, x); }
REF_EXPORT void ref_modsqrt(const spint *x, const spint *h, spint *r) { modsqrt(x, h, r); }
REF_EXPORT int ref_modis1(const spint *a) { return modis1(a); }
REF_EXPORT int ref_modis0(const spint *a) { return modis0(a); }
REF_EXPORT void ref_modzer(spint *a) { modzer(a); }
REF_EXPORT void ref_modone(spint *a) { modone(a); }
REF_EXPORT void ref_modint(int x, spint *a) { modint(x, a); }
REF_EXPORT void ref_nres(const spint *m, spint *n) { nres(m, n); }
REF_EXPORT void ref_redc(const spint *n, spint *m) { redc(n, m); }
REF_EXPORT void ref_modcsw(int b, spint *g, spint *f) { modcsw(b, g, f); }
REF_EXPORT void ref_modcmv(int b, const spint *g, spint *f) { modcmv(b, g, f); }
REF_EXPORT void ref_modshl(unsigned int n, spint *a) { modshl(n, a); }
REF_EXPORT int ref_modshr(unsigned int n, spint *a) { return modshr(n, a); }
REF_EXPORT void ref_modhaf(spint *a) { modhaf(a); }
REF_EXPORT void ref_mod2r(unsigned int r, spint *a) { mod2r(r, a); }
REF_EXPORT void ref_modexp(const spint *a, char *b) { modexp(a, b); }
REF_EXPORT int ref_modimp(const char *b, spint *a) { return modimp(b, a); }
REF_EXPORT int ref_modsign(const spint *a) { return modsign(a); }
REF_EXPORT int ref_modcmp(const spint *a, const spint *b) { return modcmp(a, b); }

/* Batched byte-level drivers: every element goes modimp -> op -> modexp through the
 * reference's own functions, so callers never touch the reference's limb layout.
 * op: 0 mul, 1 sqr, 2 inv, 3 sqrt, 4 add, 5 sub, 6 neg, 7 pro, 8 identity(imp->exp),
 *     9 mli (b = small int taken from ib), 10 haf, 11 qr (writes 0/1 into out[0]) */
REF_EXPORT void ref_field_batch(int op, const char *a, const char *b, int ib, char *out,
                                int *status, size_t n, int nthreads) {
    long i;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (i = 0; i < (long)n; i++) {
        spint x[Nlimbs], y[Nlimbs], z[Nlimbs];
        int st = modimp(a + (size_t)i * Nbytes, x);
        if (b) (void)modimp(b + (size_t)i * Nbytes, y);
        switch (op) {
            case 0: modmul(x, y, z); break;
            case 1: modsqr(x, z); break;
            case 2: modinv(x, NULL, z); break;
            case 3: modsqrt(x, NULL, z); break;
            case 4: modadd(x, y, z); break;
            case 5: modsub(x, y, z); break;
            case 6: modneg(x, z); break;
            case 7: modpro(x, z); break;
            case 9: modmli(x, ib, z); break;
            case 10: modcpy(x, z); modhaf(z); break;
            case 11: modzer(z); st = modqr(NULL, x); break;
            default: modcpy(x, z); break;
        }
        modexp(z, out + (size_t)i * Nbytes);
        if (status) status[i] = st;
    }
}

#ifdef REF_HAS_RFC7748
/* rfc7748.c:156 over n independent keys, AoS little-endian byte strings. */
REF_EXPORT void ref_rfc7748(const char *bk, const char *bu, char *bv) { rfc7748(bk, bu, bv); }

REF_EXPORT void ref_rfc7748_batch(const char *bk, const char *bu, char *bv, size_t n, int nthreads) {
    long i;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (i = 0; i < (long)n; i++)
        rfc7748(bk + (size_t)i * Nbytes, bu + (size_t)i * Nbytes, bv + (size_t)i * Nbytes);
}
#endif

REF_EXPORT int ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
