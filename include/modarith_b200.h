/* modarith_b200.h -- C ABI of the B200-native batched finite-field / Montgomery-ladder engine.
 *
 * Drop-in boundary for the reference's ONLY FFI on this path: the generators build their
 * functions non-static into test.so and bind them through ctypes (pseudo.py:1689-1750,
 * monty.py:2260-2325); consumers otherwise paste the same functions as text (rfc7748.c:24-28).
 * Every entry point below is the batched form of one generated function:
 * same name, same argument order, same aliasing freedom (outputs may alias inputs, cf.
 * pseudo.py:1832-1845), with three trailing arguments appended:
 *
 *     size_t n        number of independent elements / keys
 *     size_t stride   plane pitch in elements (>= n); limb j of element i is p[j*stride + i]
 *     void  *stream   cudaStream_t to launch on (NULL = default stream); calls are asynchronous
 *
 * and decorated  mab_<PRIME>_<function>  (the reference's own decoration scheme appends
 * _<prime>_ct, pseudo.py:1940-1944).  <PRIME> is X25519, X448 or NIST256.
 *
 * All pointers are DEVICE pointers unless the name ends in _host.  A field element is
 * MAB_<PRIME>_LIMBS planes of uint32_t in this library's internal form (saturated radix 2^32;
 * Montgomery form for NIST256) -- like the reference's spint[Nlimbs] it is opaque: convert with
 * modimp/modexp (big-endian bytes) or nres/redc (canonical little-endian 32-bit words).
 * Byte strings keep the reference's layout: element i occupies bytes [i*Nbytes, (i+1)*Nbytes).
 *
 * Return value: 0 on success, otherwise a cudaError_t value (or MAB_ERR_*).  As in the
 * reference no input is ever rejected; per-element status/predicate results go to int arrays.
 * There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef MODARITH_B200_H
#define MODARITH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MAB_API __attribute__((visibility("default")))
#else
#define MAB_API
#endif

#define MAB_X25519_LIMBS 8   /* Nlimbs  (pseudo.py:1404); Nbytes 32, Nbits 255 */
#define MAB_X448_LIMBS 14    /*                           Nbytes 56, Nbits 448 */
#define MAB_NIST256_LIMBS 8  /*                           Nbytes 32, Nbits 256 */

#define MAB_ERR_BADARG 100001
#define MAB_ERR_NODEVICE 100002
#define MAB_ERR_NOJIT 100003 /* mab_<P>_modprog_jit: the run-time compiler (NVRTC) could not be loaded */
#define MAB_ERR_JIT 100004   /* mab_<P>_modprog_jit: compilation failed; mab_jit_log() has the compiler's output */

/* ---- library-wide ---------------------------------------------------------------------- */
MAB_API const char *mab_version(void);
MAB_API const char *mab_error_string(int code);
MAB_API int mab_device_count(void);
/* The *_host entry points keep three streams and staging buffers per device between calls;
 * this frees them (optional; they are also reclaimed at process exit). */
MAB_API void mab_release_workspaces(void);
/* Output of the run-time compiler for the calling thread's last mab_<P>_modprog_jit / _modprog_cubin call. */
MAB_API const char *mab_jit_log(void);
/* Parameters of a modulus by name ("X25519", ...): Wordlength/Nlimbs/Radix/Nbits/Nbytes macros
 * of the generated header (pseudo.py:1403-1407).  Returns 0, or MAB_ERR_BADARG. */
MAB_API int mab_params(const char *prime, int *wordlength, int *nlimbs, int *radix, int *nbits, int *nbytes);
/* Algorithmic 32x32->64 limb products of one call, for roofline arithmetic:
 * what = "modmul" | "modsqr" | "modmli" | "modpro" | "modinv" | "modsqrt" | "rfc7748". */
MAB_API long long mab_products(const char *prime, const char *what);

/* INT32 multiplier-pipe microbenchmark: dependent-free streams of IMAD-class instructions.
 * variant 0: IMAD.WIDE.U32  1: IMAD.LO  2: IMAD.HI  3: IMAD.WIDE.U32.X carry chains
 *         4: IMAD.WIDE + 1 ALU op each  5: IMAD.WIDE + 2 ALU ops each  6: IADD3 only
 * Writes elapsed milliseconds and the number of multiply instructions executed by all threads. */
MAB_API int mab_imad_peak(int variant, int iters, int blocks, int threads, float *ms, double *instructions, void *stream);

/* Instruction-mix probe (generated, csrc/mab_probe.cuh): times one PTX block mixing IMAD.WIDE and
 * ALU-pipe instructions; returns the mix through name/nwide/nalu.  Variants 0..13; see tools/run_probe.py. */
MAB_API int mab_pipe_probe(int variant, int iters, int blocks, int threads, float *ms, const char **name,
                           int *nwide, int *nalu, void *stream);

/* Comparison kernel (csrc/mab_unsat29.cuh), not on the product path: c = a * b^iters mod 2^255-19 on the
 * REFERENCE's WL=32 limb plan (unsaturated radix 2^29, 9 limb planes, pseudo.py:124-140), so that the
 * choice of saturated limbs is backed by a measurement.  Limbs in, limbs out (value = sum c[k]*2^(29k)). */
MAB_API int mab_probe_unsat29_modmul(const uint32_t *a, const uint32_t *b, uint32_t *c, unsigned int iters,
                                     size_t n, size_t stride, void *stream);

/* ---- straight-line programs of field operations (mab_<P>_modprog) ------------------------------ */
/* What a consumer of the generated functions actually runs is a SEQUENCE of calls on a few variables (the
 * complete point addition of weierstrass.c:69-160 is 14 modmul + 23 modadd/modsub on 10 spint arrays).  One
 * launch per call moves 3 x 4 x Nlimbs bytes per element through HBM each time; mab_<P>_modprog executes the
 * whole sequence in ONE launch with the variables held on chip.  A program is an array of mab_insn over
 * MAB_PROG_NREG field registers: registers 0 .. nin-1 are loaded from the limb planes in[0..nin-1] before the
 * first instruction, and register out_reg[k] is stored to the limb planes out[k] after the last (k < nout).
 * Semantics of every operation = the generated function of the same name (same results as calling
 * mab_<P>_mod<op> with the same operands); dst may equal a or b, as the reference allows aliasing.
 * `code`, `in`, `out`, `out_reg` are HOST arrays (of device plane pointers); at most MAB_PROG_MAX instructions. */
#define MAB_PROG_NREG 16
#define MAB_PROG_MAX 320
enum mab_opcode {
  MAB_OP_ADD = 0, /* dst = a + b        modadd */
  MAB_OP_SUB,     /* dst = a - b        modsub */
  MAB_OP_NEG,     /* dst = -a           modneg */
  MAB_OP_MUL,     /* dst = a * b        modmul */
  MAB_OP_SQR,     /* dst = a^2          modsqr */
  MAB_OP_MLI,     /* dst = a * imm      modmli, 0 <= imm < 2^31 */
  MAB_OP_CPY,     /* dst = a            modcpy */
  MAB_OP_NSQR,    /* dst = a^(2^imm)    modcpy + modnsqr */
  MAB_OP_PRO,     /* dst = a^PE         modpro */
  MAB_OP_INV,     /* dst = 1/a, 0 -> 0  modinv(a, NULL, dst) */
  MAB_OP_SQRT,    /* dst = sqrt(a)      modsqrt(a, NULL, dst) */
  MAB_OP_ZER,     /* dst = 0            modzer */
  MAB_OP_ONE,     /* dst = 1            modone */
  MAB_OP_INT,     /* dst = imm          modint */
  MAB_OP_HAF,     /* dst = a / 2        modcpy + modhaf */
  MAB_OP_COUNT
};
typedef struct mab_insn {
  unsigned char op, dst, a, b; /* enum mab_opcode, then register numbers 0 .. MAB_PROG_NREG-1 */
  uint32_t imm;                /* small-integer operand of MLI / NSQR / INT */
} mab_insn;
/* Compiled programs (mab_<P>_modprog_jit).  The reference's model is textual: the generator prints functions, the
 * consumer pastes them into its source and compiles (pseudo.py:1694-1702; rfc7748.c:24-28, weierstrass.c:16-20).
 * mab_<P>_modprog_jit does that for a program handed over at run time: it prints a kernel that calls the generated
 * functions on variables held in MACHINE REGISTERS, compiles it with NVRTC for sm_100a against the headers the
 * library itself was built from (embedded as text) and caches the loaded kernel, keyed by the program.  The first
 * call of a new program costs a compilation (about a second); every later call is one launch.  Same instruction
 * set, arguments and bit-identical results as mab_<P>_modprog, which interprets the program with its variables in
 * shared memory and needs no compiler.  NVRTC is loaded on first use (libnvrtc.so.12, or the path in the
 * environment variable MAB_NVRTC); without it the call returns MAB_ERR_NOJIT -- there is no silent fall-back.
 * mab_<P>_modprog_cubin returns the compiled cubin of a program for inspection (needs no device): *size is the
 * capacity of buf on entry and the cubin's size on return; buf may be NULL to ask for the size. */

/* ---- per-modulus API (P = X25519, X448, NIST256) ------------------------------------------ */
#define MAB_DECLARE_FIELD(P)                                                                              \
  /* macros of the generated header (pseudo.py:1403-1407) plus the modpro chain cost; any pointer may be NULL */ \
  MAB_API int mab_##P##_info(int *nlimbs, int *nbits, int *nbytes, int *pm1d2, int *pro_sqr, int *pro_mul, int *montgomery, int *has_curve); \
  /* modfsb   pseudo.py:272-283   canonicalise in place; was_lt[i]=1 iff stored value was < p (may be NULL) */ \
  MAB_API int mab_##P##_modfsb(uint32_t *n_, int *was_lt, size_t n, size_t stride, void *stream);                 \
  /* modadd   pseudo.py:286-304 */                                                                        \
  MAB_API int mab_##P##_modadd(const uint32_t *a, const uint32_t *b, uint32_t *n_, size_t n, size_t stride, void *stream); \
  /* modsub   pseudo.py:307-326 */                                                                        \
  MAB_API int mab_##P##_modsub(const uint32_t *a, const uint32_t *b, uint32_t *n_, size_t n, size_t stride, void *stream); \
  /* modneg   pseudo.py:329-348 */                                                                        \
  MAB_API int mab_##P##_modneg(const uint32_t *b, uint32_t *n_, size_t n, size_t stride, void *stream);           \
  /* modmul   pseudo.py:616-659, monty.py:663-872 */                                                      \
  MAB_API int mab_##P##_modmul(const uint32_t *a, const uint32_t *b, uint32_t *c, size_t n, size_t stride, void *stream); \
  /* measurement helper, not part of the reference API: c = a * b^iters, product kept in registers */  \
  MAB_API int mab_##P##_bench_modmul(const uint32_t *a, const uint32_t *b, uint32_t *c, unsigned int iters, size_t n, size_t stride, void *stream); \
  /* modsqr   pseudo.py:663-702, monty.py:982-1165 */                                                     \
  MAB_API int mab_##P##_modsqr(const uint32_t *a, uint32_t *c, size_t n, size_t stride, void *stream);            \
  /* modmli   pseudo.py:705-728, monty.py:876-978   0 <= b < 2^31 */                                      \
  MAB_API int mab_##P##_modmli(const uint32_t *a, int b, uint32_t *c, size_t n, size_t stride, void *stream);     \
  /* modcpy   pseudo.py:730-743 */                                                                        \
  MAB_API int mab_##P##_modcpy(const uint32_t *a, uint32_t *c, size_t n, size_t stride, void *stream);            \
  /* modnsqr  pseudo.py:745-755   square in place k times */                                              \
  MAB_API int mab_##P##_modnsqr(uint32_t *a, int k, size_t n, size_t stride, void *stream);                       \
  /* modpro   pseudo.py:758-785   z = w^((p-1-2^k)/2^(k+1)) */                                            \
  MAB_API int mab_##P##_modpro(const uint32_t *w, uint32_t *z, size_t n, size_t stride, void *stream);            \
  /* modinv   pseudo.py:788-812   h = progenitor planes or NULL; 0 -> 0 */                                \
  MAB_API int mab_##P##_modinv(const uint32_t *x, const uint32_t *h, uint32_t *z, size_t n, size_t stride, void *stream); \
  /* modinv with one progenitor chain per element, as the reference spends it (modinv itself shares one chain among 8 elements when h == NULL) */ \
  MAB_API int mab_##P##_modinv_perelement(const uint32_t *x, uint32_t *z, size_t n, size_t stride, void *stream); \
  /* modqr    pseudo.py:815-831   note (h, x) order; out[i] = 1 iff x is a QR or 0 */                     \
  MAB_API int mab_##P##_modqr(const uint32_t *h, const uint32_t *x, int *out, size_t n, size_t stride, void *stream); \
  /* modsqrt  pseudo.py:834-874 */                                                                        \
  MAB_API int mab_##P##_modsqrt(const uint32_t *x, const uint32_t *h, uint32_t *r, size_t n, size_t stride, void *stream); \
  /* modis1 / modis0  pseudo.py:877-906 */                                                                \
  MAB_API int mab_##P##_modis1(const uint32_t *a, int *out, size_t n, size_t stride, void *stream);               \
  MAB_API int mab_##P##_modis0(const uint32_t *a, int *out, size_t n, size_t stride, void *stream);               \
  /* modzer / modone / modint  pseudo.py:909-949 */                                                       \
  MAB_API int mab_##P##_modzer(uint32_t *a, size_t n, size_t stride, void *stream);                               \
  MAB_API int mab_##P##_modone(uint32_t *a, size_t n, size_t stride, void *stream);                               \
  MAB_API int mab_##P##_modint(int x, uint32_t *a, size_t n, size_t stride, void *stream);                        \
  /* nres / redc  pseudo.py:952-976, monty.py:1386-1416.  Plain side = canonical value as           */    \
  /* MAB_<P>_LIMBS little-endian 32-bit words (planes).                                              */    \
  MAB_API int mab_##P##_nres(const uint32_t *m, uint32_t *n_, size_t n, size_t stride, void *stream);             \
  MAB_API int mab_##P##_redc(const uint32_t *n_, uint32_t *m, size_t n, size_t stride, void *stream);             \
  /* modcsw / modcmv  pseudo.py:979-1048   b[i] in {0,1} per element */                                   \
  MAB_API int mab_##P##_modcsw(const int *b, uint32_t *g, uint32_t *f, size_t n, size_t stride, void *stream);    \
  MAB_API int mab_##P##_modcmv(const int *b, const uint32_t *g, uint32_t *f, size_t n, size_t stride, void *stream); \
  /* modshl / modshr  pseudo.py:1052-1081   k < 32; see DESIGN.md for the saturated-limb semantics */     \
  MAB_API int mab_##P##_modshl(unsigned int k, uint32_t *a, size_t n, size_t stride, void *stream);               \
  MAB_API int mab_##P##_modshr(unsigned int k, uint32_t *a, int *out, size_t n, size_t stride, void *stream);     \
  /* modhaf   pseudo.py:1084-1100;  mod2r  pseudo.py:1102-1112 */                                         \
  MAB_API int mab_##P##_modhaf(uint32_t *a, size_t n, size_t stride, void *stream);                               \
  MAB_API int mab_##P##_mod2r(unsigned int r, uint32_t *a, size_t n, size_t stride, void *stream);                \
  /* modexp   pseudo.py:1115-1127   canonical big-endian Nbytes per element */                            \
  MAB_API int mab_##P##_modexp(const uint32_t *a, char *b, size_t n, size_t stride, void *stream);                \
  /* modimp   pseudo.py:1130-1146   status[i] = 1 iff the integer was < p (status may be NULL) */         \
  MAB_API int mab_##P##_modimp(const char *b, uint32_t *a, int *status, size_t n, size_t stride, void *stream);   \
  /* modsign / modcmp  pseudo.py:1149-1174 */                                                             \
  MAB_API int mab_##P##_modsign(const uint32_t *a, int *out, size_t n, size_t stride, void *stream);              \
  MAB_API int mab_##P##_modcmp(const uint32_t *a, const uint32_t *b, int *out, size_t n, size_t stride, void *stream); \
  /* a sequence of the calls above in one launch, see "straight-line programs" */                       \
  MAB_API int mab_##P##_modprog(const mab_insn *code, size_t ncode, const uint32_t *const *in, int nin, uint32_t *const *out, \
                                const unsigned char *out_reg, int nout, size_t n, size_t stride, void *stream); \
  /* the same program COMPILED: see "compiled programs" above.  Same arguments, same results */         \
  MAB_API int mab_##P##_modprog_jit(const mab_insn *code, size_t ncode, const uint32_t *const *in, int nin, uint32_t *const *out, \
                                    const unsigned char *out_reg, int nout, size_t n, size_t stride, void *stream); \
  MAB_API int mab_##P##_modprog_cubin(const mab_insn *code, size_t ncode, int nin, const unsigned char *out_reg, int nout, \
                                      void *buf, size_t *size);

MAB_DECLARE_FIELD(X25519)
MAB_DECLARE_FIELD(X448)
MAB_DECLARE_FIELD(NIST256)
/* Two moduli outside the three hot-path ones: the secp256k1 field prime 2^256 - 2^32 - 977 (monty.py:2066-2067;
 * here a plain-residue plan that folds 2^256 == 2^32 + 977) and the order of the P-256 group (the reference's
 * "00<decimal>" mode, monty.py:2110-2127; the generator's fall-back plan: full Montgomery, any odd modulus,
 * monty.py:2237-2244).  Further moduli: `python -m modarith_b200.build --prime NAME[=<expression>]` builds an add-on
 * library libmodarith_b200_<NAME>.so that exports MAB_DECLARE_FIELD(NAME) ("Textual inclusion and
 * further moduli"). */
MAB_DECLARE_FIELD(SECP256K1)
MAB_DECLARE_FIELD(NIST256ORDER)

/* ---- short-Weierstrass scalar multiplication (weierstrass.c) ------------------- */
/* For each of n independent points: ecnXXXset(0, x, y, &P); ecnXXXmul(e, &P); ecnXXXget(&P, xo, yo)
 * (weierstrass.c:415-427, 494-542, 333-349; curve constants curve.py:157-166).  All five arrays hold
 * big-endian 32-byte strings, element i at offset 32*i, device pointers.  A point that is not on the
 * curve, a zero scalar or a multiple of the group order give the point at infinity, reported as
 * (0, 1) exactly as ecnXXXget does.  Constant time: fixed-window signed digits, masked table scans. */
MAB_API int mab_NIST256_ecnmul(const char *e, const char *x, const char *y, char *xo, char *yo, size_t n, void *stream);
/* The same three calls on the twisted Edwards curve Ed25519 (edwards.c:347-356, 435-484, 219-241; constants
 * curve.py:85-94) over the 2^255-19 field code; the identity is reported as (0, 1). */
MAB_API int mab_ED25519_ecnmul(const char *e, const char *x, const char *y, char *xo, char *yo, size_t n, void *stream);
/* Double multiplication, the verification building block (nist256.c:251):
 * ecnXXXset(0, x1, y1, &P); ecnXXXset(0, x2, y2, &Q); ecnXXXmul2(e, &P, f, &Q, &R); ecnXXXget(&R, xo, yo)
 * (weierstrass.c:545-572 / edwards.c:486-513): (xo, yo) = e*(x1,y1) + f*(x2,y2).  Same string conventions.
 * The reference's routine is variable time (it skips zero digits); this one does the same work for every
 * point.  e = f = 0 gives (0, 1) (the reference reads outside its digit array in that case). */
MAB_API int mab_NIST256_ecnmul2(const char *e, const char *x1, const char *y1, const char *f, const char *x2, const char *y2,
                                char *xo, char *yo, size_t n, void *stream);
MAB_API int mab_ED25519_ecnmul2(const char *e, const char *x1, const char *y1, const char *f, const char *x2, const char *y2,
                                char *xo, char *yo, size_t n, void *stream);

/* ---- RFC 7748 (rfc7748.c:156  void rfc7748(const char *bk, const char *bu, char *bv)) ---- */
/* The ladder entry points of one Montgomery curve, described one by one below for X25519 and X448.  An add-on
 * library built with `python -m modarith_b200.build --prime NAME[=<expression>] --a24 N --cof K [--generator G]`
 * exports the same set for a user-defined curve -- the counterpart of adding an A24 / COF / GENERATOR block to
 * "Describe Montgomery Curve parameters" in rfc7748.c:117-132. */
#define MAB_DECLARE_CURVE(P)                                                                                  \
  MAB_API int mab_##P##_rfc7748(const char *bk, const char *bu, char *bv, size_t n, void *stream);            \
  MAB_API int mab_##P##_rfc7748_perkey(const char *bk, const char *bu, char *bv, size_t n, void *stream);     \
  MAB_API int mab_##P##_rfc7748_validate(const char *bk, const char *bu, char *bv, size_t n, void *stream);   \
  MAB_API int mab_##P##_rfc7748_host(const char *bk, const char *bu, char *bv, size_t n, int device);         \
  MAB_API int mab_##P##_rfc7748_host_multi(const char *bk, const char *bu, char *bv, size_t n, int ndev);

/* bv[i] = clamp(bk[i]) * bu[i]; little-endian Nbytes strings, n keys, device pointers. */
MAB_API int mab_X25519_rfc7748(const char *bk, const char *bu, char *bv, size_t n, void *stream);
MAB_API int mab_X448_rfc7748(const char *bk, const char *bu, char *bv, size_t n, void *stream);
/* Same results from the plain kernel (one key per thread, one inversion per key); the default entry
 * points above share one inversion among up to four keys per thread.  Kept for comparison. */
MAB_API int mab_X25519_rfc7748_perkey(const char *bk, const char *bu, char *bv, size_t n, void *stream);
MAB_API int mab_X448_rfc7748_perkey(const char *bk, const char *bu, char *bv, size_t n, void *stream);
/* The same driver compiled WITHOUT `#define TWIST_SECURE` (rfc7748.c:228-251): cheap point validation
 * (eprint 2020/1497); bv[i] is all zero when bu[i] is not the x-coordinate of a point on the curve. */
MAB_API int mab_X25519_rfc7748_validate(const char *bk, const char *bu, char *bv, size_t n, void *stream);
MAB_API int mab_X448_rfc7748_validate(const char *bk, const char *bu, char *bv, size_t n, void *stream);
/* Same with HOST pointers: copies in, runs on `device`, copies out, returns when bv is complete.
 * Chunked over three streams so transfers overlap the ladder; pinned memory is used at full
 * speed, pageable memory works but serialises the copies. */
MAB_API int mab_X25519_rfc7748_host(const char *bk, const char *bu, char *bv, size_t n, int device);
MAB_API int mab_X448_rfc7748_host(const char *bk, const char *bu, char *bv, size_t n, int device);
/* The whole box: what replaces a consumer's loop over keys (the timing loop of the reference's driver,
 * rfc7748.c:301-304, calls rfc7748() once per key) when more than one GPU is present.  Device g of `ndev`
 * takes the contiguous key range [g*n/ndev, (g+1)*n/ndev); one host thread and one set of streams per
 * device, no exchange between devices.  ndev <= 0 uses every visible device; ndev larger
 * than the device count is MAB_ERR_BADARG.  Page-locked buffers (cudaHostAlloc / cudaHostRegister) are read
 * and written in place by every GPU; pageable buffers are staged per device.  Returns when bv is complete. */
MAB_API int mab_X25519_rfc7748_host_multi(const char *bk, const char *bu, char *bv, size_t n, int ndev);
MAB_API int mab_X448_rfc7748_host_multi(const char *bk, const char *bu, char *bv, size_t n, int ndev);

#ifdef __cplusplus
}
#endif
#endif /* MODARITH_B200_H */
