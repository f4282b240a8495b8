"""Print one modulus' field code as a CUDA header (`field_<NAME>.cuh`).

Counterpart of `functions()` in the reference generators (pseudo.py:1413-1445,
monty.py:1885-1918), which print ~32 C functions for one prime.  Here the
prime-specific arithmetic (mul/sqr/mli/add/sub/neg/canon as inline-PTX blocks, the
modpro addition chain, the constants) is printed as a struct `F_<NAME>` of static
device functions, and the prime-independent functions of the API (modinv, modsqrt,
modqr, modimp/modexp, cswap, the RFC 7748 ladder, ...) are hand-written templates
over that struct in csrc/mab_field.cuh and csrc/rfc7748_sm100.cuh -- the same split
as the reference's generated field.c + hand-written rfc7748.c.
"""
from __future__ import annotations

from .. import addchain
from .plan import Plan, words


def _arr(name, L, const=False):
    return "%suint32_t (&%s)[%d]" % ("const " if const else "", name, L)


def _block_fn(plan: Plan, fname, asm, sig, ret=None):
    s = "  static MAB_DEV %s %s(%s) {\n" % ("uint32_t" if ret else "void", fname, sig)
    if ret:
        s += "    uint32_t %s;\n" % ret
    s += "#ifndef MAB_HOSTSIM\n"
    s += asm.emit_cuda("    ")
    s += "#else\n"
    s += asm.emit_sim("    ")
    s += "#endif\n"
    if ret:
        s += "    return %s;\n" % ret
    s += "  }\n\n"
    return s


def _const_fn(fname, ws, L):
    body = " ".join("r[%d] = 0x%08xu;" % (i, w) for i, w in enumerate(ws))
    return "  static MAB_DEV void %s(%s) { %s }\n" % (fname, _arr("r", L), body)


def emit_field_header(plan: Plan) -> str:
    P, L = plan.P, plan.L
    blocks = plan.build()
    plan.self_check(trials=120)
    prog = addchain.find_chain(P.pe)
    assert addchain.evaluate(prog, 3, P.p) == pow(3, P.pe, P.p)
    nsq, nmu = addchain.cost(prog)
    S = "F_" + P.name
    out = []
    out.append("// Automatically generated field arithmetic for sm_100a -- do not edit.\n")
    out.append("// Command line : python -m modarith_b200.gen.%s_sm100 %s\n" %
               ("pseudo" if P.family == "pseudo" else "monty", P.name))
    out.append("// modulus %s = 0x%x\n" % (P.name, P.p))
    out.append("// plan %s: %d saturated 32-bit limbs; stored values < %s; R = 2^%d\n" % (
        type(plan).__name__, L, "p" if plan.bound == P.p else "2^%d" % (32 * L),
        plan.R.bit_length() - 1))
    weak = "mul_w" in blocks
    for k in ("mul", "sqr", "mli", "mla", "add", "sub", "canon") + (("add_tt", "sub_tt") if plan.tight else ()) + \
            (("mul_w", "sqr_w") if weak else ()):
        w, i, a = blocks[k].stats()
        out.append("//   %-5s : %3d IMAD.WIDE  %2d IMAD  ~%3d ALU-pipe ops\n" % (k, w, i, a))
    out.append("//   modpro: %d squarings + %d multiplies (exponent (p-1-2^k)/2^(k+1), k=%d)\n" %
               (nsq, nmu, P.pm1d2))
    out.append("#pragma once\n#include \"mab_common.cuh\"\n\n")
    out.append("struct %s {\n" % S)
    out.append("  static constexpr int L = %d;\n" % L)
    out.append("  static constexpr int NBITS = %d;\n" % P.nbits)
    out.append("  static constexpr int NBYTES = %d;\n" % P.nbytes)
    out.append("  static constexpr int PM1D2 = %d;\n" % P.pm1d2)
    out.append("  static constexpr bool MONTGOMERY = %s;\n" % ("true" if plan.R != 1 else "false"))
    out.append("  static constexpr int PRO_SQR = %d, PRO_MUL = %d;\n" % (nsq, nmu))
    out.append("  static constexpr int LADDER_MINBLOCKS = %d;   // resident 128-thread CTAs per SM for k_rfc7748\n"
               % plan.ladder_minblocks)
    out.append("  static constexpr bool LADDER_STASH = %s;   // scalar and x1 in shared memory (see rfc7748_sm100.cuh)\n"
               % ("true" if plan.ladder_stash else "false"))
    if P.a24 is not None:
        out.append("  static constexpr bool HAS_CURVE = true;\n")
        out.append("  static constexpr uint32_t A24 = %d;\n" % P.a24)
        out.append("  static constexpr int COF = %d;\n" % P.cof)
        out.append("  static constexpr uint32_t GENERATOR = %d;\n" % P.generator)
    else:
        out.append("  static constexpr bool HAS_CURVE = false;\n")
        out.append("  static constexpr uint32_t A24 = 0;\n  static constexpr int COF = 0;\n"
                   "  static constexpr uint32_t GENERATOR = 0;\n")
    out.append("  static const char* name() { return \"%s\"; }\n\n" % P.name)

    a, b, r = _arr("a", L, True), _arr("b", L, True), _arr("r", L)
    out.append("  // c = a*b (pseudo.py:616-659 / monty.py:663-872)\n")
    out.append(_block_fn(plan, "mul", blocks["mul"], "%s, %s, %s" % (r, a, b)))
    out.append("  // c = a*a (pseudo.py:663-702 / monty.py:982-1165)\n")
    out.append(_block_fn(plan, "sqr", blocks["sqr"], "%s, %s" % (r, a)))
    out.append("  // c = a*b for a small integer b (pseudo.py:705-728 / monty.py:876-978)\n")
    out.append(_block_fn(plan, "mli", blocks["mli"], "%s, %s, uint32_t b" % (r, a)))
    out.append("  // r = a*b + c, small integer b: modmli + modadd fused (rfc7748.c:209,212)\n")
    out.append(_block_fn(plan, "mla", blocks["mla"], "%s, %s, uint32_t b, %s" % (r, a, _arr("c", L, True))))
    out.append("  // n = a+b (pseudo.py:286-304)\n")
    out.append(_block_fn(plan, "add", blocks["add"], "%s, %s, %s" % (r, a, b)))
    out.append("  // n = a-b (pseudo.py:307-326)\n")
    out.append(_block_fn(plan, "sub", blocks["sub"], "%s, %s, %s" % (r, a, b)))
    if plan.tight:
        out.append("  // Products (mul, sqr) of this field stay below 2^%d + %d*2^13 whatever their operands; add_tt / sub_tt\n"
                   "  // are modadd / modsub for two such values (no second wrap to handle); their results are ordinary\n"
                   "  // stored values (< 2^%d), which every function accepts.\n" % (P.nbits, (1 << P.nbits) - P.p, 32 * L))
        out.append("  static constexpr bool TIGHT = true;\n")
        tw = words(plan.tight, L)
        out.append("#ifdef MAB_HOSTSIM\n  static inline bool sim_tight(%s) {   // a < 2^%d + %d*2^13 ?\n"
                   "    static const uint32_t t[%d] = {%s};\n"
                   "    for (int i = %d; i >= 0; i--) { if (a[i] != t[i]) return a[i] < t[i]; }\n    return false;\n  }\n#endif\n"
                   % (_arr("a", L, True), P.nbits, (1 << P.nbits) - P.p, L, ", ".join("0x%08xu" % w for w in tw), L - 1))
        chk = ('#ifdef MAB_HOSTSIM\n    MAB_SIM_REQUIRE(sim_tight(a) && sim_tight(b), "%s.%s operands below the product bound");\n#endif\n')
        for fn in ("add_tt", "sub_tt"):
            body = _block_fn(plan, fn, blocks[fn], "%s, %s, %s" % (r, a, b))
            head, rest = body.split("{\n", 1)
            out.append(head + "{\n" + chk % (P.name, fn) + rest)
    else:
        out.append("  // no spare bit above Nbits in this plan: the product-operand forms are the general ones\n")
        out.append("  static constexpr bool TIGHT = false;\n")
        out.append("  static MAB_DEV void add_tt(%s, %s, %s) { add(r, a, b); }\n" % (r, a, b))
        out.append("  static MAB_DEV void sub_tt(%s, %s, %s) { sub(r, a, b); }\n\n" % (r, a, b))
    if weak:
        out.append("  // Weakly reduced products for chains (modpro, modnsqr): operands and results are any representative\n"
                   "  // below 2^%d; the last step looks at the carry word only.  A chain ends with canon(), which restores\n"
                   "  // the [0, p) invariant that every other function of this field keeps.\n" % (32 * L))
        out.append("  static constexpr bool WEAK = true;\n")
        out.append(_block_fn(plan, "mul_w", blocks["mul_w"], "%s, %s, %s" % (r, a, b)))
        out.append(_block_fn(plan, "sqr_w", blocks["sqr_w"], "%s, %s" % (r, a)))
    else:
        out.append("  // no separate weakly-reduced products in this plan: chains use the ordinary ones\n")
        out.append("  static constexpr bool WEAK = false;\n")
        out.append("  static MAB_DEV void mul_w(%s, %s, %s) { mul(r, a, b); }\n" % (r, a, b))
        out.append("  static MAB_DEV void sqr_w(%s, %s) { sqr(r, a); }\n\n" % (r, a))
    out.append("  // n = -b (pseudo.py:329-348)\n")
    out.append(_block_fn(plan, "neg", blocks["neg"], "%s, %s" % (r, b)))
    out.append("  // canonical residue of a stored value; returns 1 iff it was already < p\n"
               "  // (flatten/modfsb, pseudo.py:255-283)\n")
    out.append(_block_fn(plan, "canon", blocks["canon"], "%s, %s" % (r, a), ret="lt"))

    out.append(_const_fn("set_p", words(P.p, L), L))
    out.append(_const_fn("set_one", words(plan.to_internal(1), L), L))
    out.append(_const_fn("set_roi", words(plan.to_internal(P.roi), L), L))
    out.append(_const_fn("set_r2", words(plan.R * plan.R % P.p, L), L))
    if P.ed_d is not None:
        out.append("  // twisted Edwards curve -x^2 + y^2 = 1 + d x^2 y^2 (curve.py:85-94): d in stored form\n")
        out.append(_const_fn("set_ed_d", words(plan.to_internal(P.ed_d), L), L))
    if P.wb is not None:
        out.append("  // short-Weierstrass curve y^2 = x^3 - 3x + b (curve.py:157-166): b in stored form\n")
        out.append("  static constexpr bool HAS_WEIERSTRASS = true;\n")
        out.append(_const_fn("set_b", words(plan.to_internal(P.wb), L), L))
    else:
        out.append("  static constexpr bool HAS_WEIERSTRASS = false;\n")
    out.append("\n")
    if plan.R != 1:
        out.append("  // nres: multiply by R^2 mod p (monty.py:1386-1399); redc: multiply by 1 (monty.py:1402-1416)\n")
        out.append("  static MAB_DEV void nres(%s, %s) { uint32_t c[L]; set_r2(c); mul(r, a, c); }\n" % (r, a))
        out.append("  static MAB_DEV void redc(%s, %s) { uint32_t c[L]; c[0] = 1;\n"
                   "    for (int i = 1; i < L; i++) c[i] = 0;\n    mul(r, a, c); }\n\n" % (r, a))
    else:
        out.append("  // nres: copy (pseudo.py:952-962); redc: copy + final subtract (pseudo.py:965-976)\n")
        out.append("  static MAB_DEV void nres(%s, %s) { for (int i = 0; i < L; i++) r[i] = a[i]; }\n" % (r, a))
        out.append("  static MAB_DEV void redc(%s, %s) { (void)canon(r, a); }\n\n" % (r, a))

    # progenitor chain (pseudo.py:758-785)
    tmps = addchain.temporaries(prog)
    out.append("  // z = w^PE, straight-line addition chain (pseudo.py:758-785; our own chain finder)\n")
    out.append("  static MAB_DEV void pro(%s, %s) {\n" % (_arr("z", L), _arr("w", L, True)))
    out.append("    uint32_t x[L];\n    for (int i = 0; i < L; i++) x[i] = w[i];\n")
    for t in tmps:
        out.append("    uint32_t %s[L];\n" % t)
    for op in prog:
        if op[0] == "mul":
            out.append("    mul_w(%s, %s, %s);\n" % (op[1], op[2], op[3]))
        else:
            dst, src, n = op[1], op[2], op[3]
            if n == 0:
                out.append("    for (int i = 0; i < L; i++) %s[i] = %s[i];\n" % (dst, src))
            else:
                out.append("    sqr_w(%s, %s);\n" % (dst, src))
                if n == 2:
                    out.append("    sqr_w(%s, %s);\n" % (dst, dst))
                elif n > 2:
                    out.append("    MAB_NOUNROLL\n    for (int i = 1; i < %d; i++) sqr_w(%s, %s);\n" % (n, dst, dst))
    out.append("    if (WEAK) (void)canon(z, z);\n")
    out.append("  }\n")
    out.append("};\n")
    return "".join(out)
