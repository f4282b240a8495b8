/* oracle/ref_curve_shim.c -- TEST INFRASTRUCTURE ONLY.  Appended after the reference's weierstrass.c
 * (patched by its own curve.py for NIST256) in a scratch translation unit: a batch driver around
 * ecnXXXset / ecnXXXmul / ecnXXXget (weierstrass.c:415-427, 494-542, 333-349). */
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif
__attribute__((visibility("default")))
void ref_ecnmul_batch(const char *e, const char *x, const char *y, char *xo, char *yo, size_t n, int nthreads) {
    long i;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (i = 0; i < (long)n; i++) {
        point P;
        ecn_nist256_set(0, x + (size_t)i * BYTES, y + (size_t)i * BYTES, &P);
        ecn_nist256_mul(e + (size_t)i * BYTES, &P);
        ecn_nist256_get(&P, xo + (size_t)i * BYTES, yo + (size_t)i * BYTES);
    }
}

/* ecnXXXset x2 / ecnXXXmul2 / ecnXXXget (weierstrass.c:545-572, edwards.c:486-513): R = e*P + f*Q */
__attribute__((visibility("default")))
void ref_ecnmul2_batch(const char *e, const char *x1, const char *y1, const char *f, const char *x2, const char *y2,
                       char *xo, char *yo, size_t n, int nthreads) {
    long i;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (i = 0; i < (long)n; i++) {
        point P, Q, R;
        size_t o = (size_t)i * BYTES;
        ecn_nist256_set(0, x1 + o, y1 + o, &P);
        ecn_nist256_set(0, x2 + o, y2 + o, &Q);
        ecn_nist256_mul2(e + o, &P, f + o, &Q, &R);
        ecn_nist256_get(&R, xo + o, yo + o);
    }
}

/* The flow of the reference's testcurve.c main (testcurve.c:213-300) with a selectable iteration count:
 * P = Q = generator; `iters` times P = a*P; then `iters` times P = a*P + b*Q.  Writes the generator, the
 * point after the first loop and the final point (affine, big-endian). */
__attribute__((visibility("default")))
void ref_testcurve_chain(const char *a, const char *b, int iters, char *gx, char *gy, char *x1, char *y1, char *x2, char *y2) {
    point P, Q;
    int i;
    ecn_nist256_gen(&P);
    ecn_nist256_get(&P, gx, gy);
    ecn_nist256_cpy(&P, &Q);
    for (i = 0; i < iters; i++) ecn_nist256_mul(a, &P);
    ecn_nist256_get(&P, x1, y1);
    for (i = 0; i < iters; i++) ecn_nist256_mul2(a, &P, b, &Q, &P);
    ecn_nist256_get(&P, x2, y2);
}
