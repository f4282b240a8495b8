/* oracle/oracle.c -- plain-C restatement of the reference's batch hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and the cpu_baseline leg of bench.py may build, load or call
 * this file; the product (modarith_b200/) never does.
 *
 * What it restates, with the reference file:line each function follows:
 *   - the value semantics of the generated field functions (pseudo.py:223-1174, monty.py:352-1619)
 *     for the three moduli of the hot path, on canonical residues held as little-endian 32-bit words;
 *   - the RFC 7748 ladder driver rfc7748.c:156-256 (clamp :135-141, bit :144-146, mask :148-152,
 *     step :186-221, TWIST_SECURE tail :225-227,252, export :254-255), byte for byte.
 * It deliberately shares nothing with the product: no saturated-limb tricks, no Montgomery form, no
 * special squaring -- schoolbook product, then "x = (x mod 2^n) + (x >> n) * (2^n - p)" until it fits,
 * then conditional subtraction.  The limb layout, R factor and lazily reduced "< 2p" representatives of
 * the reference are invisible after redc/modexp, which is the level this oracle works at.
 *
 * Pinned (tests/test_oracle_pinned.py): RFC 7748 vectors of rfc7748.c:271,274 and
 * simd/rfc7748_simt.cu:245,249, the deterministic outputs of rfc7748.c:main, golden vectors generated
 * by the reference's own C (tests/golden), and oracle/_ref on random inputs.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define MAXW 14          /* 448 bits */
typedef struct {
    int nbits, nw, nbytes;          /* Nbits, words, Nbytes (pseudo.py:1403-1407) */
    uint32_t p[MAXW];               /* modulus */
    uint32_t delta[MAXW];           /* 2^nbits - p */
    int dw;                         /* words of delta */
    uint32_t pe[MAXW];              /* progenitor exponent (p-1-2^k)/2^(k+1)  (pseudo.py:1574-1581) */
    int k;                          /* PM1D2 */
    uint32_t roi[MAXW];             /* 2^k-th root of unity (pseudo.py:1616-1630) */
    uint32_t a24; int cof; uint32_t gen;   /* rfc7748.c:120-132 */
} fld;

static fld F25519, F448, FP256, FSECP, FORD;
static int inited = 0;

/* ---- tiny multiword helpers (little-endian 32-bit words) ------------------------------------- */
static int cmpw(const uint32_t *a, const uint32_t *b, int n) {
    for (int i = n - 1; i >= 0; i--) { if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1; }
    return 0;
}
static uint32_t addw(uint32_t *r, const uint32_t *a, const uint32_t *b, int n) {
    uint64_t c = 0;
    for (int i = 0; i < n; i++) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
static uint32_t subw(uint32_t *r, const uint32_t *a, const uint32_t *b, int n) {
    int64_t c = 0;
    for (int i = 0; i < n; i++) { c += (int64_t)a[i] - b[i]; r[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)(c & 1);
}
static void mulw(uint32_t *r, const uint32_t *a, int na, const uint32_t *b, int nb) {
    memset(r, 0, 4 * (size_t)(na + nb));
    for (int i = 0; i < na; i++) {
        uint64_t c = 0;
        for (int j = 0; j < nb; j++) { c += (uint64_t)a[i] * b[j] + r[i + j]; r[i + j] = (uint32_t)c; c >>= 32; }
        r[i + nb] = (uint32_t)c;
    }
}
static int bitw(const uint32_t *a, int i) { return (a[i >> 5] >> (i & 31)) & 1; }

/* x (nx words) -> canonical residue mod p in f->nw words */
static void reduce(const fld *f, uint32_t *out, const uint32_t *x, int nx) {
    uint32_t cur[2 * MAXW + 2], lo[MAXW + 1], hi[2 * MAXW + 2], t[3 * MAXW + 4];
    int n = nx, nw = f->nw, sh = f->nbits & 31, w0 = f->nbits >> 5;
    memset(cur, 0, sizeof cur);
    memcpy(cur, x, 4 * (size_t)nx);
    for (;;) {
        /* hi = cur >> nbits, lo = cur mod 2^nbits */
        int nh = n - w0; if (nh < 0) nh = 0;
        int any = 0;
        for (int i = 0; i < nh; i++) {
            uint64_t v = cur[w0 + i] >> sh;
            if (sh && w0 + i + 1 < n) v |= (uint64_t)cur[w0 + i + 1] << (32 - sh);
            hi[i] = (uint32_t)v; any |= hi[i] != 0;
        }
        memset(lo, 0, sizeof lo);
        for (int i = 0; i < nw; i++) lo[i] = i < n ? cur[i] : 0;
        if (sh) lo[w0] &= (1u << sh) - 1u;
        if (!any) { memcpy(cur, lo, 4 * (size_t)nw); break; }
        mulw(t, hi, nh, f->delta, f->dw);                      /* hi * (2^n - p) */
        int nt = nh + f->dw; if (nt < nw + 1) { memset(t + nt, 0, 4 * (size_t)(nw + 1 - nt)); nt = nw + 1; }
        uint64_t c = 0;
        for (int i = 0; i < nt; i++) { c += (uint64_t)t[i] + (i < nw ? lo[i] : 0); cur[i] = (uint32_t)c; c >>= 32; }
        cur[nt] = (uint32_t)c; n = nt + 1;
        for (int i = n; i < 2 * MAXW + 2; i++) cur[i] = 0;
    }
    while (cmpw(cur, f->p, nw) >= 0) subw(cur, cur, f->p, nw);
    memcpy(out, cur, 4 * (size_t)nw);
}

/* ---- generated-code API, value level ------------------------------------------------------------ */
static void f_mul(const fld *f, uint32_t *r, const uint32_t *a, const uint32_t *b) {   /* modmul pseudo.py:616-659 */
    uint32_t t[2 * MAXW]; mulw(t, a, f->nw, b, f->nw); reduce(f, r, t, 2 * f->nw);
}
static void f_sqr(const fld *f, uint32_t *r, const uint32_t *a) { f_mul(f, r, a, a); }   /* modsqr pseudo.py:663-702 */
static void f_add(const fld *f, uint32_t *r, const uint32_t *a, const uint32_t *b) {   /* modadd pseudo.py:286-304 */
    uint32_t t[MAXW + 1]; t[f->nw] = addw(t, a, b, f->nw); reduce(f, r, t, f->nw + 1);
}
static void f_sub(const fld *f, uint32_t *r, const uint32_t *a, const uint32_t *b) {   /* modsub pseudo.py:307-326 */
    uint32_t t[MAXW];
    if (subw(t, a, b, f->nw)) addw(t, t, f->p, f->nw);           /* a,b canonical: one correction */
    memcpy(r, t, 4 * (size_t)f->nw);
}
static void f_mli(const fld *f, uint32_t *r, const uint32_t *a, uint32_t b) {          /* modmli pseudo.py:705-728 */
    uint32_t t[MAXW + 1]; mulw(t, a, f->nw, &b, 1); reduce(f, r, t, f->nw + 1);
}
static void f_set(const fld *f, uint32_t *r, uint32_t v) { memset(r, 0, 4 * (size_t)f->nw); r[0] = v; }
static int f_is(const fld *f, const uint32_t *a, uint32_t v) {                          /* modis1/modis0 pseudo.py:877-906 */
    if (a[0] != v) return 0;
    for (int i = 1; i < f->nw; i++) if (a[i]) return 0;
    return 1;
}
static void f_pro(const fld *f, uint32_t *z, const uint32_t *w) {                       /* modpro pseudo.py:758-785: w^PE */
    uint32_t acc[MAXW]; f_set(f, acc, 1);
    int top = 32 * f->nw - 1; while (top > 0 && !bitw(f->pe, top)) top--;
    for (int i = top; i >= 0; i--) { f_sqr(f, acc, acc); if (bitw(f->pe, i)) f_mul(f, acc, acc, w); }
    memcpy(z, acc, 4 * (size_t)f->nw);
}
static void f_inv(const fld *f, uint32_t *z, const uint32_t *x, const uint32_t *h) {   /* modinv pseudo.py:788-812 */
    uint32_t s[MAXW], t[MAXW];
    if (h) memcpy(t, h, 4 * (size_t)f->nw); else f_pro(f, t, x);
    memcpy(s, x, 4 * (size_t)f->nw);
    for (int i = 0; i < f->k - 1; i++) { f_sqr(f, s, s); f_mul(f, s, s, x); }
    for (int i = 0; i < f->k + 1; i++) f_sqr(f, t, t);
    f_mul(f, z, s, t);
}
static void f_sqrt(const fld *f, uint32_t *r, const uint32_t *x, const uint32_t *h) {  /* modsqrt pseudo.py:834-874 */
    uint32_t s[MAXW], y[MAXW], t[MAXW], b[MAXW], v[MAXW], z[MAXW];
    if (h) memcpy(y, h, 4 * (size_t)f->nw); else f_pro(f, y, x);
    f_mul(f, s, y, x);
    if (f->k > 1) {
        f_mul(f, t, s, y);
        memcpy(z, f->roi, 4 * (size_t)f->nw);
        for (int k = f->k; k > 1; k--) {
            memcpy(b, t, 4 * (size_t)f->nw);
            for (int i = 0; i < k - 2; i++) f_sqr(f, b, b);
            int d = 1 - f_is(f, b, 1);
            f_mul(f, v, s, z); if (d) memcpy(s, v, 4 * (size_t)f->nw);      /* modcmv(d,v,s) */
            f_sqr(f, z, z);
            f_mul(f, v, t, z); if (d) memcpy(t, v, 4 * (size_t)f->nw);
        }
    }
    memcpy(r, s, 4 * (size_t)f->nw);
}
static int f_qr(const fld *f, const uint32_t *x) {                                      /* modqr pseudo.py:815-831 */
    uint32_t r[MAXW];
    f_pro(f, r, x); f_sqr(f, r, r); f_mul(f, r, r, x);
    for (int i = 0; i < f->k - 1; i++) f_sqr(f, r, r);
    return f_is(f, r, 1) | f_is(f, x, 0);
}
static void f_haf(const fld *f, uint32_t *r, const uint32_t *a) {                       /* modhaf pseudo.py:1084-1100 */
    uint32_t t[MAXW + 1]; memcpy(t, a, 4 * (size_t)f->nw); t[f->nw] = 0;
    if (t[0] & 1) t[f->nw] = addw(t, t, f->p, f->nw);
    for (int i = 0; i < f->nw; i++) r[i] = (t[i] >> 1) | (t[i + 1] << 31);
}
/* modimp pseudo.py:1130-1146: big-endian Nbytes -> value; returns 1 iff the integer was < p */
static int f_imp(const fld *f, uint32_t *a, const unsigned char *b) {
    uint32_t t[MAXW]; memset(t, 0, sizeof t);
    for (int i = 0; i < f->nbytes; i++) { int pos = f->nbytes - 1 - i; t[pos >> 2] |= (uint32_t)b[i] << (8 * (pos & 3)); }
    int lt = cmpw(t, f->p, f->nw) < 0;
    reduce(f, a, t, f->nw);
    return lt;
}
static void f_exp(const fld *f, unsigned char *b, const uint32_t *a) {                  /* modexp pseudo.py:1115-1127 */
    for (int i = 0; i < f->nbytes; i++) { int pos = f->nbytes - 1 - i; b[i] = (unsigned char)(a[pos >> 2] >> (8 * (pos & 3))); }
}

/* ---- constants ------------------------------------------------------------------------------------- */
static void shr_words(uint32_t *a, int n, int s) { for (int i = 0; i < n; i++) a[i] = (a[i] >> s) | (i + 1 < n ? a[i + 1] << (32 - s) : 0); }
static void finish(fld *f) {
    uint32_t pow2[MAXW + 1]; memset(pow2, 0, sizeof pow2);
    f->nw = (f->nbits + 31) / 32; f->nbytes = (f->nbits + 7) / 8;
    pow2[f->nbits >> 5] = 1u << (f->nbits & 31);
    uint32_t pp[MAXW + 1]; memcpy(pp, f->p, 4 * MAXW); pp[MAXW] = 0;
    uint32_t d[MAXW + 1]; subw(d, pow2, pp, MAXW + 1);
    memcpy(f->delta, d, 4 * MAXW); f->dw = f->nw; while (f->dw > 1 && !f->delta[f->dw - 1]) f->dw--;
    /* k = 2-adicity of p-1; PE = (p-1-2^k) / 2^(k+1) */
    uint32_t pm1[MAXW]; memcpy(pm1, f->p, 4 * MAXW); pm1[0] -= 1;
    f->k = 0; while (!bitw(pm1, f->k)) f->k++;
    uint32_t e[MAXW]; memset(e, 0, sizeof e); e[0] = 1u << f->k;
    subw(f->pe, pm1, e, MAXW); shr_words(f->pe, MAXW, f->k + 1);
    /* root of unity (pseudo.py:1616-1627): p-1 (k=1), 2^((p-1)/4) (k=2), qnr^((p-1)/2^k) (k>2) */
    if (f->k == 1) memcpy(f->roi, pm1, 4 * MAXW);
    else {
        uint32_t base = 2;
        if (f->k > 2) {                       /* smallest quadratic non-residue */
            uint32_t half[MAXW]; memcpy(half, pm1, 4 * MAXW); shr_words(half, MAXW, 1);
            for (;; base++) {
                uint32_t acc[MAXW], b[MAXW]; f_set(f, acc, 1); f_set(f, b, base);
                for (int i = 32 * f->nw - 1; i >= 0; i--) { f_sqr(f, acc, acc); if (bitw(half, i)) f_mul(f, acc, acc, b); }
                if (!f_is(f, acc, 1)) break;
            }
        }
        uint32_t ex[MAXW]; memcpy(ex, pm1, 4 * MAXW); shr_words(ex, MAXW, f->k);
        uint32_t acc[MAXW], b[MAXW]; f_set(f, acc, 1); f_set(f, b, base);
        for (int i = 32 * f->nw - 1; i >= 0; i--) { f_sqr(f, acc, acc); if (bitw(ex, i)) f_mul(f, acc, acc, b); }
        memcpy(f->roi, acc, 4 * MAXW);
    }
}
static void init(void) {
    if (inited) return;
    memset(&F25519, 0, sizeof F25519); memset(&F448, 0, sizeof F448); memset(&FP256, 0, sizeof FP256);
    memset(&FSECP, 0, sizeof FSECP); memset(&FORD, 0, sizeof FORD);
    /* 2^255-19 (pseudo.py:1523-1524) */
    F25519.nbits = 255; for (int i = 0; i < 8; i++) F25519.p[i] = 0xffffffffu; F25519.p[0] = 0xffffffedu; F25519.p[7] = 0x7fffffffu;
    F25519.a24 = 121665; F25519.cof = 3; F25519.gen = 9;
    /* 2^448-2^224-1 (monty.py:1986-1987) */
    F448.nbits = 448; for (int i = 0; i < 14; i++) F448.p[i] = 0xffffffffu; F448.p[7] = 0xfffffffeu;
    F448.a24 = 39081; F448.cof = 2; F448.gen = 5;
    /* NIST P-256 (monty.py:1966-1967) */
    FP256.nbits = 256;
    { const uint32_t w[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0, 1, 0xffffffffu}; memcpy(FP256.p, w, sizeof w); }
    /* secp256k1 field prime 2^256-2^32-977 (monty.py:2066-2067); order of the P-256 group (monty.py:2110-2127) */
    FSECP.nbits = 256; for (int i = 0; i < 8; i++) FSECP.p[i] = 0xffffffffu; FSECP.p[0] = 0xfffffc2fu; FSECP.p[1] = 0xfffffffeu;
    FORD.nbits = 256;
    { const uint32_t w[8] = {0xfc632551u, 0xf3b9cac2u, 0xa7179e84u, 0xbce6faadu, 0xffffffffu, 0xffffffffu, 0, 0xffffffffu}; memcpy(FORD.p, w, sizeof w); }
    finish(&F25519); finish(&F448); finish(&FP256); finish(&FSECP); finish(&FORD);
    inited = 1;
}
static const fld *field(int id) {
    init();
    return id == 0 ? &F25519 : id == 1 ? &F448 : id == 2 ? &FP256 : id == 3 ? &FSECP : &FORD;
}

/* ---- exported: batched byte-level drivers (same shape as oracle/ref_shim.c) ------------------------- */
#define EXPORT __attribute__((visibility("default")))

/* op: 0 mul 1 sqr 2 inv 3 sqrt 4 add 5 sub 6 neg 7 pro 8 id 9 mli 10 haf 11 qr ; big-endian Nbytes strings */
EXPORT void oracle_field_batch(int prime, int op, const unsigned char *a, const unsigned char *b, int ib,
                               unsigned char *out, int *status, size_t n) {
    const fld *f = field(prime);
    for (size_t i = 0; i < n; i++) {
        uint32_t x[MAXW], y[MAXW], z[MAXW];
        memset(z, 0, sizeof z);
        int st = f_imp(f, x, a + i * (size_t)f->nbytes);
        if (b) (void)f_imp(f, y, b + i * (size_t)f->nbytes);
        switch (op) {
            case 0: f_mul(f, z, x, y); break;
            case 1: f_sqr(f, z, x); break;
            case 2: f_inv(f, z, x, NULL); break;
            case 3: f_sqrt(f, z, x, NULL); break;
            case 4: f_add(f, z, x, y); break;
            case 5: f_sub(f, z, x, y); break;
            case 6: f_set(f, y, 0); f_sub(f, z, y, x); break;
            case 7: f_pro(f, z, x); break;
            case 9: f_mli(f, z, x, (uint32_t)ib); break;
            case 10: f_haf(f, z, x); break;
            case 11: st = f_qr(f, x); break;
            default: memcpy(z, x, sizeof z); break;
        }
        f_exp(f, out + i * (size_t)f->nbytes, z);
        if (status) status[i] = st;
    }
}

/* rfc7748.c:156-256, one key; prime 0 = X25519, 1 = X448 */
static void rfc7748_one(const fld *f, const unsigned char *bk, const unsigned char *bu, unsigned char *bv) {
    unsigned char ck[64], cu[64], ob[64];
    int nb = f->nbytes;
    for (int i = 0; i < nb; i++) { ck[i] = bk[i]; cu[i] = bu[nb - 1 - i]; }      /* copy; reverse(cu) :166-171 */
    int r = f->nbits % 8; if (r == 0) r = 8;
    cu[0] &= (unsigned char)((1 << r) - 1);                                       /* mask() :148-152,172 */
    int s = (8 - (f->nbits % 8)) % 8;                                             /* clamp() :135-141 */
    ck[0] &= (unsigned char)(-(1 << f->cof));
    ck[nb - 1] &= (unsigned char)(0xffu >> s);
    ck[nb - 1] |= (unsigned char)(0x80u >> s);
    uint32_t u[MAXW], x1[MAXW], x2[MAXW], z2[MAXW], x3[MAXW], z3[MAXW];
    uint32_t A[MAXW], B[MAXW], AA[MAXW], BB[MAXW], C[MAXW], D[MAXW], E[MAXW], t[MAXW];
    (void)f_imp(f, u, cu);                                                        /* :178 */
    size_t W = 4 * (size_t)f->nw;
    memcpy(x1, u, W); f_set(f, x2, 1); f_set(f, z2, 0); memcpy(x3, u, W); f_set(f, z3, 1);
    int swap = 0;
    for (int i = f->nbits - 1; i >= 0; i--) {                                     /* :186-221 */
        int kt = (ck[i / 8] >> (i % 8)) & 1;
        swap ^= kt;
        if (swap) { memcpy(t, x2, W); memcpy(x2, x3, W); memcpy(x3, t, W); memcpy(t, z2, W); memcpy(z2, z3, W); memcpy(z3, t, W); }
        swap = kt;
        f_add(f, A, x2, z2); f_add(f, C, x3, z3);
        f_sub(f, B, x2, z2); f_sub(f, D, x3, z3);
        f_sqr(f, AA, A); f_sqr(f, BB, B);
        f_mul(f, D, D, A); f_mul(f, C, C, B);
        f_sub(f, z3, D, C); f_sub(f, E, AA, BB);
        f_mli(f, z2, E, f->a24);
        f_add(f, x3, D, C); f_add(f, z2, z2, AA);
        f_mul(f, z2, z2, E);
        f_sqr(f, x3, x3); f_sqr(f, z3, z3);
        f_mul(f, z3, z3, x1); f_mul(f, x2, AA, BB);
    }
    if (swap) { memcpy(t, x2, W); memcpy(x2, x3, W); memcpy(x3, t, W); memcpy(t, z2, W); memcpy(z2, z3, W); memcpy(z3, t, W); }
    f_pro(f, A, z2); f_inv(f, z2, z2, A);                                         /* :226-227 */
    f_mul(f, x2, x2, z2);                                                         /* :252 */
    f_exp(f, ob, x2);                                                             /* :254 */
    for (int i = 0; i < nb; i++) bv[i] = ob[nb - 1 - i];                          /* reverse :255 */
}

EXPORT void oracle_rfc7748_batch(int prime, const unsigned char *bk, const unsigned char *bu, unsigned char *bv, size_t n) {
    const fld *f = field(prime);
    long i;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (i = 0; i < (long)n; i++)
        rfc7748_one(f, bk + (size_t)i * f->nbytes, bu + (size_t)i * f->nbytes, bv + (size_t)i * f->nbytes);
}
