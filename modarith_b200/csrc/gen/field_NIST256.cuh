// Automatically generated field arithmetic for sm_100a -- do not edit.
// Command line : python -m modarith_b200.gen.monty_sm100 NIST256
// modulus NIST256 = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
// plan Montgomery: 8 saturated 32-bit limbs; stored values < p; R = 2^256
//   mul   :  64 IMAD.WIDE   0 IMAD  ~107 ALU-pipe ops
//   sqr   :  36 IMAD.WIDE   0 IMAD  ~119 ALU-pipe ops
//   mli   :   8 IMAD.WIDE   0 IMAD  ~102 ALU-pipe ops
//   mla   :   8 IMAD.WIDE   0 IMAD  ~103 ALU-pipe ops
//   add   :   0 IMAD.WIDE   0 IMAD  ~ 43 ALU-pipe ops
//   sub   :   0 IMAD.WIDE   0 IMAD  ~ 18 ALU-pipe ops
//   canon :   0 IMAD.WIDE   0 IMAD  ~ 34 ALU-pipe ops
//   mul_w :  64 IMAD.WIDE   0 IMAD  ~ 83 ALU-pipe ops
//   sqr_w :  36 IMAD.WIDE   0 IMAD  ~ 95 ALU-pipe ops
//   modpro: 253 squarings + 12 multiplies (exponent (p-1-2^k)/2^(k+1), k=1)
#pragma once
#include "mab_common.cuh"

struct F_NIST256 {
  static constexpr int L = 8;
  static constexpr int NBITS = 256;
  static constexpr int NBYTES = 32;
  static constexpr int PM1D2 = 1;
  static constexpr bool MONTGOMERY = true;
  static constexpr int PRO_SQR = 253, PRO_MUL = 12;
  static constexpr int LADDER_MINBLOCKS = 3;   // resident 128-thread CTAs per SM for k_rfc7748
  static constexpr bool LADDER_STASH = false;   // scalar and x1 in shared memory (see rfc7748_sm100.cuh)
  static constexpr bool HAS_CURVE = false;
  static constexpr uint32_t A24 = 0;
  static constexpr int COF = 0;
  static constexpr uint32_t GENERATOR = 0;
  static const char* name() { return "NIST256"; }

  // c = a*b (pseudo.py:616-659 / monty.py:663-872)
  static MAB_DEV void mul(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<132>;\n\t"
        "mul.lo.u32 t0, %8, %16;\n\t"
        "mul.hi.u32 t1, %8, %16;\n\t"
        "mul.lo.u32 t2, %10, %16;\n\t"
        "mul.hi.u32 t3, %10, %16;\n\t"
        "mul.lo.u32 t4, %12, %16;\n\t"
        "mul.hi.u32 t5, %12, %16;\n\t"
        "mul.lo.u32 t6, %14, %16;\n\t"
        "mul.hi.u32 t7, %14, %16;\n\t"
        "mul.lo.u32 t17, %9, %16;\n\t"
        "mul.hi.u32 t18, %9, %16;\n\t"
        "mul.lo.u32 t19, %11, %16;\n\t"
        "mul.hi.u32 t20, %11, %16;\n\t"
        "mul.lo.u32 t21, %13, %16;\n\t"
        "mul.hi.u32 t22, %13, %16;\n\t"
        "mul.lo.u32 t23, %15, %16;\n\t"
        "mul.hi.u32 t24, %15, %16;\n\t"
        "mad.lo.cc.u32 t2, %9, %17, t2;\n\t"
        "madc.hi.cc.u32 t3, %9, %17, t3;\n\t"
        "madc.lo.cc.u32 t4, %11, %17, t4;\n\t"
        "madc.hi.cc.u32 t5, %11, %17, t5;\n\t"
        "madc.lo.cc.u32 t6, %13, %17, t6;\n\t"
        "madc.hi.cc.u32 t7, %13, %17, t7;\n\t"
        "madc.lo.cc.u32 t8, %15, %17, 0x0;\n\t"
        "madc.hi.u32 t9, %15, %17, 0x0;\n\t"
        "mad.lo.cc.u32 t17, %8, %17, t17;\n\t"
        "madc.hi.cc.u32 t18, %8, %17, t18;\n\t"
        "madc.lo.cc.u32 t19, %10, %17, t19;\n\t"
        "madc.hi.cc.u32 t20, %10, %17, t20;\n\t"
        "madc.lo.cc.u32 t21, %12, %17, t21;\n\t"
        "madc.hi.cc.u32 t22, %12, %17, t22;\n\t"
        "madc.lo.cc.u32 t23, %14, %17, t23;\n\t"
        "madc.hi.cc.u32 t24, %14, %17, t24;\n\t"
        "madc.lo.u32 t25, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t2, %8, %18, t2;\n\t"
        "madc.hi.cc.u32 t3, %8, %18, t3;\n\t"
        "madc.lo.cc.u32 t4, %10, %18, t4;\n\t"
        "madc.hi.cc.u32 t5, %10, %18, t5;\n\t"
        "madc.lo.cc.u32 t6, %12, %18, t6;\n\t"
        "madc.hi.cc.u32 t7, %12, %18, t7;\n\t"
        "madc.lo.cc.u32 t8, %14, %18, t8;\n\t"
        "madc.hi.cc.u32 t9, %14, %18, t9;\n\t"
        "madc.lo.u32 t10, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t19, %9, %18, t19;\n\t"
        "madc.hi.cc.u32 t20, %9, %18, t20;\n\t"
        "madc.lo.cc.u32 t21, %11, %18, t21;\n\t"
        "madc.hi.cc.u32 t22, %11, %18, t22;\n\t"
        "madc.lo.cc.u32 t23, %13, %18, t23;\n\t"
        "madc.hi.cc.u32 t24, %13, %18, t24;\n\t"
        "madc.lo.cc.u32 t25, %15, %18, t25;\n\t"
        "madc.hi.u32 t26, %15, %18, 0x0;\n\t"
        "mad.lo.cc.u32 t4, %9, %19, t4;\n\t"
        "madc.hi.cc.u32 t5, %9, %19, t5;\n\t"
        "madc.lo.cc.u32 t6, %11, %19, t6;\n\t"
        "madc.hi.cc.u32 t7, %11, %19, t7;\n\t"
        "madc.lo.cc.u32 t8, %13, %19, t8;\n\t"
        "madc.hi.cc.u32 t9, %13, %19, t9;\n\t"
        "madc.lo.cc.u32 t10, %15, %19, t10;\n\t"
        "madc.hi.u32 t11, %15, %19, 0x0;\n\t"
        "mad.lo.cc.u32 t19, %8, %19, t19;\n\t"
        "madc.hi.cc.u32 t20, %8, %19, t20;\n\t"
        "madc.lo.cc.u32 t21, %10, %19, t21;\n\t"
        "madc.hi.cc.u32 t22, %10, %19, t22;\n\t"
        "madc.lo.cc.u32 t23, %12, %19, t23;\n\t"
        "madc.hi.cc.u32 t24, %12, %19, t24;\n\t"
        "madc.lo.cc.u32 t25, %14, %19, t25;\n\t"
        "madc.hi.cc.u32 t26, %14, %19, t26;\n\t"
        "madc.lo.u32 t27, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t4, %8, %20, t4;\n\t"
        "madc.hi.cc.u32 t5, %8, %20, t5;\n\t"
        "madc.lo.cc.u32 t6, %10, %20, t6;\n\t"
        "madc.hi.cc.u32 t7, %10, %20, t7;\n\t"
        "madc.lo.cc.u32 t8, %12, %20, t8;\n\t"
        "madc.hi.cc.u32 t9, %12, %20, t9;\n\t"
        "madc.lo.cc.u32 t10, %14, %20, t10;\n\t"
        "madc.hi.cc.u32 t11, %14, %20, t11;\n\t"
        "madc.lo.u32 t12, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t21, %9, %20, t21;\n\t"
        "madc.hi.cc.u32 t22, %9, %20, t22;\n\t"
        "madc.lo.cc.u32 t23, %11, %20, t23;\n\t"
        "madc.hi.cc.u32 t24, %11, %20, t24;\n\t"
        "madc.lo.cc.u32 t25, %13, %20, t25;\n\t"
        "madc.hi.cc.u32 t26, %13, %20, t26;\n\t"
        "madc.lo.cc.u32 t27, %15, %20, t27;\n\t"
        "madc.hi.u32 t28, %15, %20, 0x0;\n\t"
        "mad.lo.cc.u32 t6, %9, %21, t6;\n\t"
        "madc.hi.cc.u32 t7, %9, %21, t7;\n\t"
        "madc.lo.cc.u32 t8, %11, %21, t8;\n\t"
        "madc.hi.cc.u32 t9, %11, %21, t9;\n\t"
        "madc.lo.cc.u32 t10, %13, %21, t10;\n\t"
        "madc.hi.cc.u32 t11, %13, %21, t11;\n\t"
        "madc.lo.cc.u32 t12, %15, %21, t12;\n\t"
        "madc.hi.u32 t13, %15, %21, 0x0;\n\t"
        "mad.lo.cc.u32 t21, %8, %21, t21;\n\t"
        "madc.hi.cc.u32 t22, %8, %21, t22;\n\t"
        "madc.lo.cc.u32 t23, %10, %21, t23;\n\t"
        "madc.hi.cc.u32 t24, %10, %21, t24;\n\t"
        "madc.lo.cc.u32 t25, %12, %21, t25;\n\t"
        "madc.hi.cc.u32 t26, %12, %21, t26;\n\t"
        "madc.lo.cc.u32 t27, %14, %21, t27;\n\t"
        "madc.hi.cc.u32 t28, %14, %21, t28;\n\t"
        "madc.lo.u32 t29, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t6, %8, %22, t6;\n\t"
        "madc.hi.cc.u32 t7, %8, %22, t7;\n\t"
        "madc.lo.cc.u32 t8, %10, %22, t8;\n\t"
        "madc.hi.cc.u32 t9, %10, %22, t9;\n\t"
        "madc.lo.cc.u32 t10, %12, %22, t10;\n\t"
        "madc.hi.cc.u32 t11, %12, %22, t11;\n\t"
        "madc.lo.cc.u32 t12, %14, %22, t12;\n\t"
        "madc.hi.cc.u32 t13, %14, %22, t13;\n\t"
        "madc.lo.u32 t14, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t23, %9, %22, t23;\n\t"
        "madc.hi.cc.u32 t24, %9, %22, t24;\n\t"
        "madc.lo.cc.u32 t25, %11, %22, t25;\n\t"
        "madc.hi.cc.u32 t26, %11, %22, t26;\n\t"
        "madc.lo.cc.u32 t27, %13, %22, t27;\n\t"
        "madc.hi.cc.u32 t28, %13, %22, t28;\n\t"
        "madc.lo.cc.u32 t29, %15, %22, t29;\n\t"
        "madc.hi.u32 t30, %15, %22, 0x0;\n\t"
        "mad.lo.cc.u32 t8, %9, %23, t8;\n\t"
        "madc.hi.cc.u32 t9, %9, %23, t9;\n\t"
        "madc.lo.cc.u32 t10, %11, %23, t10;\n\t"
        "madc.hi.cc.u32 t11, %11, %23, t11;\n\t"
        "madc.lo.cc.u32 t12, %13, %23, t12;\n\t"
        "madc.hi.cc.u32 t13, %13, %23, t13;\n\t"
        "madc.lo.cc.u32 t14, %15, %23, t14;\n\t"
        "madc.hi.u32 t15, %15, %23, 0x0;\n\t"
        "mad.lo.cc.u32 t23, %8, %23, t23;\n\t"
        "madc.hi.cc.u32 t24, %8, %23, t24;\n\t"
        "madc.lo.cc.u32 t25, %10, %23, t25;\n\t"
        "madc.hi.cc.u32 t26, %10, %23, t26;\n\t"
        "madc.lo.cc.u32 t27, %12, %23, t27;\n\t"
        "madc.hi.cc.u32 t28, %12, %23, t28;\n\t"
        "madc.lo.cc.u32 t29, %14, %23, t29;\n\t"
        "madc.hi.cc.u32 t30, %14, %23, t30;\n\t"
        "addc.cc.u32 t31, 0x0, 0x0;\n\t"
        "addc.cc.u32 t32, t1, t17;\n\t"
        "addc.cc.u32 t33, t2, t18;\n\t"
        "addc.cc.u32 t34, t3, t19;\n\t"
        "addc.cc.u32 t35, t4, t20;\n\t"
        "addc.cc.u32 t36, t5, t21;\n\t"
        "addc.cc.u32 t37, t6, t22;\n\t"
        "addc.cc.u32 t38, t7, t23;\n\t"
        "addc.cc.u32 t39, t8, t24;\n\t"
        "addc.cc.u32 t40, t9, t25;\n\t"
        "addc.cc.u32 t41, t10, t26;\n\t"
        "addc.cc.u32 t42, t11, t27;\n\t"
        "addc.cc.u32 t43, t12, t28;\n\t"
        "addc.cc.u32 t44, t13, t29;\n\t"
        "addc.cc.u32 t45, t14, t30;\n\t"
        "addc.u32 t46, t15, t31;\n\t"
        "add.cc.u32 t47, t34, t0;\n\t"
        "addc.cc.u32 t48, t35, t32;\n\t"
        "addc.cc.u32 t49, t36, t33;\n\t"
        "addc.cc.u32 t50, t37, t34;\n\t"
        "addc.u32 t51, t38, t35;\n\t"
        "shl.b32 t52, t0, 1;\n\t"
        "shf.l.wrap.b32 t53, t0, t32, 1;\n\t"
        "add.cc.u32 t54, t50, t52;\n\t"
        "addc.u32 t55, t51, t53;\n\t"
        "sub.u32 t56, t55, t0;\n\t"
        "add.cc.u32 t57, t47, t0;\n\t"
        "addc.cc.u32 t58, t48, t32;\n\t"
        "addc.cc.u32 t59, t49, t33;\n\t"
        "addc.cc.u32 t60, t54, t47;\n\t"
        "addc.cc.u32 t61, t56, t48;\n\t"
        "addc.cc.u32 t62, 0x0, t49;\n\t"
        "addc.cc.u32 t63, 0x0, t54;\n\t"
        "addc.cc.u32 t64, 0x0, t56;\n\t"
        "addc.cc.u32 t65, 0x0, 0x0;\n\t"
        "madc.lo.u32 t66, 0, 0, 0;\n\t"
        "add.cc.u32 t67, t59, t0;\n\t"
        "addc.cc.u32 t68, t60, t32;\n\t"
        "addc.cc.u32 t69, t61, t33;\n\t"
        "addc.cc.u32 t70, t62, t47;\n\t"
        "addc.cc.u32 t71, t63, t48;\n\t"
        "addc.cc.u32 t72, t64, t49;\n\t"
        "addc.cc.u32 t73, t65, t54;\n\t"
        "addc.u32 t74, t66, t56;\n\t"
        "sub.cc.u32 t75, t58, t0;\n\t"
        "subc.cc.u32 t76, t67, t32;\n\t"
        "subc.cc.u32 t77, t68, t33;\n\t"
        "subc.cc.u32 t78, t69, t47;\n\t"
        "subc.cc.u32 t79, t70, t48;\n\t"
        "subc.cc.u32 t80, t71, t49;\n\t"
        "subc.cc.u32 t81, t72, t54;\n\t"
        "subc.cc.u32 t82, t73, t56;\n\t"
        "subc.u32 t83, t74, 0x0;\n\t"
        "add.cc.u32 t84, t34, t0;\n\t"
        "addc.cc.u32 t85, t35, t32;\n\t"
        "addc.cc.u32 t86, t36, t33;\n\t"
        "addc.cc.u32 t87, t37, t57;\n\t"
        "addc.cc.u32 t88, t38, t75;\n\t"
        "addc.cc.u32 t89, t39, t76;\n\t"
        "addc.cc.u32 t90, t40, t77;\n\t"
        "addc.cc.u32 t91, t41, t78;\n\t"
        "addc.cc.u32 t92, t42, t79;\n\t"
        "addc.cc.u32 t93, t43, t80;\n\t"
        "addc.cc.u32 t94, t44, t81;\n\t"
        "addc.cc.u32 t95, t45, t82;\n\t"
        "addc.cc.u32 t96, t46, t83;\n\t"
        "madc.lo.u32 t97, 0, 0, 0;\n\t"
        "sub.cc.u32 t98, t89, 0xffffffff;\n\t"
        "subc.cc.u32 t99, t90, 0xffffffff;\n\t"
        "subc.cc.u32 t100, t91, 0xffffffff;\n\t"
        "subc.cc.u32 t101, t92, 0x0;\n\t"
        "subc.cc.u32 t102, t93, 0x0;\n\t"
        "subc.cc.u32 t103, t94, 0x0;\n\t"
        "subc.cc.u32 t104, t95, 0x1;\n\t"
        "subc.cc.u32 t105, t96, 0xffffffff;\n\t"
        "subc.cc.u32 t106, t97, 0x0;\n\t"
        "subc.u32 t107, 0x0, 0x0;\n\t"
        "xor.b32 t108, t98, t89;\n\t"
        "and.b32 t109, t108, t107;\n\t"
        "xor.b32 t110, t109, t98;\n\t"
        "xor.b32 t111, t99, t90;\n\t"
        "and.b32 t112, t111, t107;\n\t"
        "xor.b32 t113, t112, t99;\n\t"
        "xor.b32 t114, t100, t91;\n\t"
        "and.b32 t115, t114, t107;\n\t"
        "xor.b32 t116, t115, t100;\n\t"
        "xor.b32 t117, t101, t92;\n\t"
        "and.b32 t118, t117, t107;\n\t"
        "xor.b32 t119, t118, t101;\n\t"
        "xor.b32 t120, t102, t93;\n\t"
        "and.b32 t121, t120, t107;\n\t"
        "xor.b32 t122, t121, t102;\n\t"
        "xor.b32 t123, t103, t94;\n\t"
        "and.b32 t124, t123, t107;\n\t"
        "xor.b32 t125, t124, t103;\n\t"
        "xor.b32 t126, t104, t95;\n\t"
        "and.b32 t127, t126, t107;\n\t"
        "xor.b32 t128, t127, t104;\n\t"
        "xor.b32 t129, t105, t96;\n\t"
        "and.b32 t130, t129, t107;\n\t"
        "xor.b32 t131, t130, t105;\n\t"
        "mov.u32 %0, t110;\n\t"
        "mov.u32 %1, t113;\n\t"
        "mov.u32 %2, t116;\n\t"
        "mov.u32 %3, t119;\n\t"
        "mov.u32 %4, t122;\n\t"
        "mov.u32 %5, t125;\n\t"
        "mov.u32 %6, t128;\n\t"
        "mov.u32 %7, t131;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t b_0_i = b[0];
    const uint32_t b_1_i = b[1];
    const uint32_t b_2_i = b[2];
    const uint32_t b_3_i = b[3];
    const uint32_t b_4_i = b[4];
    const uint32_t b_5_i = b[5];
    const uint32_t b_6_i = b[6];
    const uint32_t b_7_i = b[7];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127, t128, t129, t130, t131;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t0 = (uint32_t)((uint32_t)(a_0_i * b_0_i));
    t1 = (uint32_t)(((uint64_t)a_0_i * b_0_i) >> 32);
    t2 = (uint32_t)((uint32_t)(a_2_i * b_0_i));
    t3 = (uint32_t)(((uint64_t)a_2_i * b_0_i) >> 32);
    t4 = (uint32_t)((uint32_t)(a_4_i * b_0_i));
    t5 = (uint32_t)(((uint64_t)a_4_i * b_0_i) >> 32);
    t6 = (uint32_t)((uint32_t)(a_6_i * b_0_i));
    t7 = (uint32_t)(((uint64_t)a_6_i * b_0_i) >> 32);
    t17 = (uint32_t)((uint32_t)(a_1_i * b_0_i));
    t18 = (uint32_t)(((uint64_t)a_1_i * b_0_i) >> 32);
    t19 = (uint32_t)((uint32_t)(a_3_i * b_0_i));
    t20 = (uint32_t)(((uint64_t)a_3_i * b_0_i) >> 32);
    t21 = (uint32_t)((uint32_t)(a_5_i * b_0_i));
    t22 = (uint32_t)(((uint64_t)a_5_i * b_0_i) >> 32);
    t23 = (uint32_t)((uint32_t)(a_7_i * b_0_i));
    t24 = (uint32_t)(((uint64_t)a_7_i * b_0_i) >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * b_1_i) + t2; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_1_i) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_1_i) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_1_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_1_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_1_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_1_i) + 0x0u + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_1_i) >> 32) + 0x0u + cf_; t9 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_1_i) + t17; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_1_i) >> 32) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_1_i) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_1_i) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_1_i) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_1_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_1_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_1_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t25 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_2_i) + t2; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_2_i) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_2_i) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_2_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_2_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_2_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_2_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_2_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_2_i) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_2_i) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_2_i) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_2_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_2_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_2_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_2_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_2_i) >> 32) + 0x0u + cf_; t26 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_3_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_3_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_3_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_3_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_3_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_3_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_3_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_3_i) >> 32) + 0x0u + cf_; t11 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_3_i) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_3_i) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_3_i) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_3_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_3_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_3_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_3_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_3_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_4_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_4_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_4_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_4_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_4_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_4_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_4_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_4_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_4_i) + t21; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_4_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_4_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_4_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_4_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_4_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_4_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_4_i) >> 32) + 0x0u + cf_; t28 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_5_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_5_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_5_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_5_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_5_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_5_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_5_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_5_i) >> 32) + 0x0u + cf_; t13 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_5_i) + t21; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_5_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_5_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_5_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_5_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_5_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_5_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_5_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_6_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_6_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_6_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_6_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_6_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_6_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_6_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_6_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t14 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_6_i) + t23; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_6_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_6_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_6_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_6_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_6_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_6_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_6_i) >> 32) + 0x0u + cf_; t30 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_7_i) + t8; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_7_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_7_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_7_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_7_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_7_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_7_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_7_i) >> 32) + 0x0u + cf_; t15 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_7_i) + t23; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_7_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_7_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_7_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_7_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_7_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_7_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_7_i) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t1 + t17 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t18 + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t19 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t20 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t21 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t22 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t23 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t8 + t24 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t25 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t26 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t27 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t28 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t29 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + t30 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + t31 + cf_; t46 = (uint32_t)w_;
    w_ = (uint64_t)t34 + t0; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t35 + t32 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t36 + t33 + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t37 + t34 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t38 + t35 + cf_; t51 = (uint32_t)w_;
    t52 = (uint32_t)(t0 << 1);
    t53 = (uint32_t)(((((uint64_t)t32 << 32) | t0) << 1) >> 32);
    w_ = (uint64_t)t50 + t52; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t51 + t53 + cf_; t55 = (uint32_t)w_;
    w_ = (uint64_t)t55 - t0; t56 = (uint32_t)w_;
    w_ = (uint64_t)t47 + t0; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t48 + t32 + cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t49 + t33 + cf_; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t54 + t47 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t56 + t48 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t49 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t54 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t56 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t66 = (uint32_t)w_;
    w_ = (uint64_t)t59 + t0; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t60 + t32 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t61 + t33 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t62 + t47 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + t48 + cf_; t71 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t64 + t49 + cf_; t72 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t65 + t54 + cf_; t73 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t66 + t56 + cf_; t74 = (uint32_t)w_;
    w_ = (uint64_t)t58 - t0; t75 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t67 - t32 - cf_; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t68 - t33 - cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t69 - t47 - cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t70 - t48 - cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t71 - t49 - cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t72 - t54 - cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t73 - t56 - cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t74 - 0x0u - cf_; t83 = (uint32_t)w_;
    w_ = (uint64_t)t34 + t0; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t35 + t32 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t36 + t33 + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t37 + t57 + cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t38 + t75 + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t39 + t76 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + t77 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t41 + t78 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t42 + t79 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t43 + t80 + cf_; t93 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t44 + t81 + cf_; t94 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t45 + t82 + cf_; t95 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t46 + t83 + cf_; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t97 = (uint32_t)w_;
    w_ = (uint64_t)t89 - 0xffffffffu; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t90 - 0xffffffffu - cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t91 - 0xffffffffu - cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t92 - 0x0u - cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t93 - 0x0u - cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t94 - 0x0u - cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t95 - 0x1u - cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t96 - 0xffffffffu - cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t97 - 0x0u - cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t107 = (uint32_t)w_;
    t108 = (uint32_t)(t98 ^ t89);
    t109 = (uint32_t)(t108 & t107);
    t110 = (uint32_t)(t109 ^ t98);
    t111 = (uint32_t)(t99 ^ t90);
    t112 = (uint32_t)(t111 & t107);
    t113 = (uint32_t)(t112 ^ t99);
    t114 = (uint32_t)(t100 ^ t91);
    t115 = (uint32_t)(t114 & t107);
    t116 = (uint32_t)(t115 ^ t100);
    t117 = (uint32_t)(t101 ^ t92);
    t118 = (uint32_t)(t117 & t107);
    t119 = (uint32_t)(t118 ^ t101);
    t120 = (uint32_t)(t102 ^ t93);
    t121 = (uint32_t)(t120 & t107);
    t122 = (uint32_t)(t121 ^ t102);
    t123 = (uint32_t)(t103 ^ t94);
    t124 = (uint32_t)(t123 & t107);
    t125 = (uint32_t)(t124 ^ t103);
    t126 = (uint32_t)(t104 ^ t95);
    t127 = (uint32_t)(t126 & t107);
    t128 = (uint32_t)(t127 ^ t104);
    t129 = (uint32_t)(t105 ^ t96);
    t130 = (uint32_t)(t129 & t107);
    t131 = (uint32_t)(t130 ^ t105);
    r[0] = t110;
    r[1] = t113;
    r[2] = t116;
    r[3] = t119;
    r[4] = t122;
    r[5] = t125;
    r[6] = t128;
    r[7] = t131;
#endif
  }

  // c = a*a (pseudo.py:663-702 / monty.py:982-1165)
  static MAB_DEV void sqr(uint32_t (&r)[8], const uint32_t (&a)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<162>;\n\t"
        "mul.lo.u32 t2, %8, %10;\n\t"
        "mul.hi.u32 t3, %8, %10;\n\t"
        "mul.lo.u32 t4, %8, %12;\n\t"
        "mul.hi.u32 t5, %8, %12;\n\t"
        "mul.lo.u32 t6, %8, %14;\n\t"
        "mul.hi.u32 t7, %8, %14;\n\t"
        "mul.lo.u32 t17, %8, %9;\n\t"
        "mul.hi.u32 t18, %8, %9;\n\t"
        "mul.lo.u32 t19, %8, %11;\n\t"
        "mul.hi.u32 t20, %8, %11;\n\t"
        "mul.lo.u32 t21, %8, %13;\n\t"
        "mul.hi.u32 t22, %8, %13;\n\t"
        "mul.lo.u32 t23, %8, %15;\n\t"
        "mul.hi.u32 t24, %8, %15;\n\t"
        "mad.lo.cc.u32 t4, %9, %11, t4;\n\t"
        "madc.hi.cc.u32 t5, %9, %11, t5;\n\t"
        "madc.lo.cc.u32 t6, %9, %13, t6;\n\t"
        "madc.hi.cc.u32 t7, %9, %13, t7;\n\t"
        "madc.lo.cc.u32 t8, %9, %15, 0x0;\n\t"
        "madc.hi.u32 t9, %9, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t19, %9, %10, t19;\n\t"
        "madc.hi.cc.u32 t20, %9, %10, t20;\n\t"
        "madc.lo.cc.u32 t21, %9, %12, t21;\n\t"
        "madc.hi.cc.u32 t22, %9, %12, t22;\n\t"
        "madc.lo.cc.u32 t23, %9, %14, t23;\n\t"
        "madc.hi.cc.u32 t24, %9, %14, t24;\n\t"
        "madc.lo.u32 t25, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t6, %10, %12, t6;\n\t"
        "madc.hi.cc.u32 t7, %10, %12, t7;\n\t"
        "madc.lo.cc.u32 t8, %10, %14, t8;\n\t"
        "madc.hi.cc.u32 t9, %10, %14, t9;\n\t"
        "madc.lo.u32 t10, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t21, %10, %11, t21;\n\t"
        "madc.hi.cc.u32 t22, %10, %11, t22;\n\t"
        "madc.lo.cc.u32 t23, %10, %13, t23;\n\t"
        "madc.hi.cc.u32 t24, %10, %13, t24;\n\t"
        "madc.lo.cc.u32 t25, %10, %15, t25;\n\t"
        "madc.hi.u32 t26, %10, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t8, %11, %13, t8;\n\t"
        "madc.hi.cc.u32 t9, %11, %13, t9;\n\t"
        "madc.lo.cc.u32 t10, %11, %15, t10;\n\t"
        "madc.hi.u32 t11, %11, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t23, %11, %12, t23;\n\t"
        "madc.hi.cc.u32 t24, %11, %12, t24;\n\t"
        "madc.lo.cc.u32 t25, %11, %14, t25;\n\t"
        "madc.hi.cc.u32 t26, %11, %14, t26;\n\t"
        "madc.lo.u32 t27, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t10, %12, %14, t10;\n\t"
        "madc.hi.cc.u32 t11, %12, %14, t11;\n\t"
        "madc.lo.u32 t12, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t25, %12, %13, t25;\n\t"
        "madc.hi.cc.u32 t26, %12, %13, t26;\n\t"
        "madc.lo.cc.u32 t27, %12, %15, t27;\n\t"
        "madc.hi.u32 t28, %12, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t12, %13, %15, t12;\n\t"
        "madc.hi.u32 t13, %13, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t27, %13, %14, t27;\n\t"
        "madc.hi.cc.u32 t28, %13, %14, t28;\n\t"
        "madc.lo.u32 t29, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t29, %14, %15, t29;\n\t"
        "madc.hi.cc.u32 t30, %14, %15, 0x0;\n\t"
        "addc.cc.u32 t32, t2, t18;\n\t"
        "addc.cc.u32 t33, t3, t19;\n\t"
        "addc.cc.u32 t34, t4, t20;\n\t"
        "addc.cc.u32 t35, t5, t21;\n\t"
        "addc.cc.u32 t36, t6, t22;\n\t"
        "addc.cc.u32 t37, t7, t23;\n\t"
        "addc.cc.u32 t38, t8, t24;\n\t"
        "addc.cc.u32 t39, t9, t25;\n\t"
        "addc.cc.u32 t40, t10, t26;\n\t"
        "addc.cc.u32 t41, t11, t27;\n\t"
        "addc.cc.u32 t42, t12, t28;\n\t"
        "addc.cc.u32 t43, t13, t29;\n\t"
        "addc.cc.u32 t44, 0x0, t30;\n\t"
        "madc.lo.u32 t45, 0, 0, 0;\n\t"
        "shl.b32 t46, t17, 1;\n\t"
        "shf.l.wrap.b32 t47, t17, t32, 1;\n\t"
        "shf.l.wrap.b32 t48, t32, t33, 1;\n\t"
        "shf.l.wrap.b32 t49, t33, t34, 1;\n\t"
        "shf.l.wrap.b32 t50, t34, t35, 1;\n\t"
        "shf.l.wrap.b32 t51, t35, t36, 1;\n\t"
        "shf.l.wrap.b32 t52, t36, t37, 1;\n\t"
        "shf.l.wrap.b32 t53, t37, t38, 1;\n\t"
        "shf.l.wrap.b32 t54, t38, t39, 1;\n\t"
        "shf.l.wrap.b32 t55, t39, t40, 1;\n\t"
        "shf.l.wrap.b32 t56, t40, t41, 1;\n\t"
        "shf.l.wrap.b32 t57, t41, t42, 1;\n\t"
        "shf.l.wrap.b32 t58, t42, t43, 1;\n\t"
        "shf.l.wrap.b32 t59, t43, t44, 1;\n\t"
        "shf.l.wrap.b32 t60, t44, t45, 1;\n\t"
        "mad.lo.cc.u32 t61, %8, %8, 0x0;\n\t"
        "madc.hi.cc.u32 t62, %8, %8, t46;\n\t"
        "madc.lo.cc.u32 t63, %9, %9, t47;\n\t"
        "madc.hi.cc.u32 t64, %9, %9, t48;\n\t"
        "madc.lo.cc.u32 t65, %10, %10, t49;\n\t"
        "madc.hi.cc.u32 t66, %10, %10, t50;\n\t"
        "madc.lo.cc.u32 t67, %11, %11, t51;\n\t"
        "madc.hi.cc.u32 t68, %11, %11, t52;\n\t"
        "madc.lo.cc.u32 t69, %12, %12, t53;\n\t"
        "madc.hi.cc.u32 t70, %12, %12, t54;\n\t"
        "madc.lo.cc.u32 t71, %13, %13, t55;\n\t"
        "madc.hi.cc.u32 t72, %13, %13, t56;\n\t"
        "madc.lo.cc.u32 t73, %14, %14, t57;\n\t"
        "madc.hi.cc.u32 t74, %14, %14, t58;\n\t"
        "madc.lo.cc.u32 t75, %15, %15, t59;\n\t"
        "madc.hi.u32 t76, %15, %15, t60;\n\t"
        "add.cc.u32 t77, t64, t61;\n\t"
        "addc.cc.u32 t78, t65, t62;\n\t"
        "addc.cc.u32 t79, t66, t63;\n\t"
        "addc.cc.u32 t80, t67, t64;\n\t"
        "addc.u32 t81, t68, t65;\n\t"
        "shl.b32 t82, t61, 1;\n\t"
        "shf.l.wrap.b32 t83, t61, t62, 1;\n\t"
        "add.cc.u32 t84, t80, t82;\n\t"
        "addc.u32 t85, t81, t83;\n\t"
        "sub.u32 t86, t85, t61;\n\t"
        "add.cc.u32 t87, t77, t61;\n\t"
        "addc.cc.u32 t88, t78, t62;\n\t"
        "addc.cc.u32 t89, t79, t63;\n\t"
        "addc.cc.u32 t90, t84, t77;\n\t"
        "addc.cc.u32 t91, t86, t78;\n\t"
        "addc.cc.u32 t92, 0x0, t79;\n\t"
        "addc.cc.u32 t93, 0x0, t84;\n\t"
        "addc.cc.u32 t94, 0x0, t86;\n\t"
        "addc.cc.u32 t95, 0x0, 0x0;\n\t"
        "madc.lo.u32 t96, 0, 0, 0;\n\t"
        "add.cc.u32 t97, t89, t61;\n\t"
        "addc.cc.u32 t98, t90, t62;\n\t"
        "addc.cc.u32 t99, t91, t63;\n\t"
        "addc.cc.u32 t100, t92, t77;\n\t"
        "addc.cc.u32 t101, t93, t78;\n\t"
        "addc.cc.u32 t102, t94, t79;\n\t"
        "addc.cc.u32 t103, t95, t84;\n\t"
        "addc.u32 t104, t96, t86;\n\t"
        "sub.cc.u32 t105, t88, t61;\n\t"
        "subc.cc.u32 t106, t97, t62;\n\t"
        "subc.cc.u32 t107, t98, t63;\n\t"
        "subc.cc.u32 t108, t99, t77;\n\t"
        "subc.cc.u32 t109, t100, t78;\n\t"
        "subc.cc.u32 t110, t101, t79;\n\t"
        "subc.cc.u32 t111, t102, t84;\n\t"
        "subc.cc.u32 t112, t103, t86;\n\t"
        "subc.u32 t113, t104, 0x0;\n\t"
        "add.cc.u32 t114, t64, t61;\n\t"
        "addc.cc.u32 t115, t65, t62;\n\t"
        "addc.cc.u32 t116, t66, t63;\n\t"
        "addc.cc.u32 t117, t67, t87;\n\t"
        "addc.cc.u32 t118, t68, t105;\n\t"
        "addc.cc.u32 t119, t69, t106;\n\t"
        "addc.cc.u32 t120, t70, t107;\n\t"
        "addc.cc.u32 t121, t71, t108;\n\t"
        "addc.cc.u32 t122, t72, t109;\n\t"
        "addc.cc.u32 t123, t73, t110;\n\t"
        "addc.cc.u32 t124, t74, t111;\n\t"
        "addc.cc.u32 t125, t75, t112;\n\t"
        "addc.cc.u32 t126, t76, t113;\n\t"
        "madc.lo.u32 t127, 0, 0, 0;\n\t"
        "sub.cc.u32 t128, t119, 0xffffffff;\n\t"
        "subc.cc.u32 t129, t120, 0xffffffff;\n\t"
        "subc.cc.u32 t130, t121, 0xffffffff;\n\t"
        "subc.cc.u32 t131, t122, 0x0;\n\t"
        "subc.cc.u32 t132, t123, 0x0;\n\t"
        "subc.cc.u32 t133, t124, 0x0;\n\t"
        "subc.cc.u32 t134, t125, 0x1;\n\t"
        "subc.cc.u32 t135, t126, 0xffffffff;\n\t"
        "subc.cc.u32 t136, t127, 0x0;\n\t"
        "subc.u32 t137, 0x0, 0x0;\n\t"
        "xor.b32 t138, t128, t119;\n\t"
        "and.b32 t139, t138, t137;\n\t"
        "xor.b32 t140, t139, t128;\n\t"
        "xor.b32 t141, t129, t120;\n\t"
        "and.b32 t142, t141, t137;\n\t"
        "xor.b32 t143, t142, t129;\n\t"
        "xor.b32 t144, t130, t121;\n\t"
        "and.b32 t145, t144, t137;\n\t"
        "xor.b32 t146, t145, t130;\n\t"
        "xor.b32 t147, t131, t122;\n\t"
        "and.b32 t148, t147, t137;\n\t"
        "xor.b32 t149, t148, t131;\n\t"
        "xor.b32 t150, t132, t123;\n\t"
        "and.b32 t151, t150, t137;\n\t"
        "xor.b32 t152, t151, t132;\n\t"
        "xor.b32 t153, t133, t124;\n\t"
        "and.b32 t154, t153, t137;\n\t"
        "xor.b32 t155, t154, t133;\n\t"
        "xor.b32 t156, t134, t125;\n\t"
        "and.b32 t157, t156, t137;\n\t"
        "xor.b32 t158, t157, t134;\n\t"
        "xor.b32 t159, t135, t126;\n\t"
        "and.b32 t160, t159, t137;\n\t"
        "xor.b32 t161, t160, t135;\n\t"
        "mov.u32 %0, t140;\n\t"
        "mov.u32 %1, t143;\n\t"
        "mov.u32 %2, t146;\n\t"
        "mov.u32 %3, t149;\n\t"
        "mov.u32 %4, t152;\n\t"
        "mov.u32 %5, t155;\n\t"
        "mov.u32 %6, t158;\n\t"
        "mov.u32 %7, t161;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127, t128, t129, t130, t131, t132, t133, t134, t135, t136, t137, t138, t139, t140, t141, t142, t143, t144, t145, t146, t147, t148, t149, t150, t151, t152, t153, t154, t155, t156, t157, t158, t159, t160, t161;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t2 = (uint32_t)((uint32_t)(a_0_i * a_2_i));
    t3 = (uint32_t)(((uint64_t)a_0_i * a_2_i) >> 32);
    t4 = (uint32_t)((uint32_t)(a_0_i * a_4_i));
    t5 = (uint32_t)(((uint64_t)a_0_i * a_4_i) >> 32);
    t6 = (uint32_t)((uint32_t)(a_0_i * a_6_i));
    t7 = (uint32_t)(((uint64_t)a_0_i * a_6_i) >> 32);
    t17 = (uint32_t)((uint32_t)(a_0_i * a_1_i));
    t18 = (uint32_t)(((uint64_t)a_0_i * a_1_i) >> 32);
    t19 = (uint32_t)((uint32_t)(a_0_i * a_3_i));
    t20 = (uint32_t)(((uint64_t)a_0_i * a_3_i) >> 32);
    t21 = (uint32_t)((uint32_t)(a_0_i * a_5_i));
    t22 = (uint32_t)(((uint64_t)a_0_i * a_5_i) >> 32);
    t23 = (uint32_t)((uint32_t)(a_0_i * a_7_i));
    t24 = (uint32_t)(((uint64_t)a_0_i * a_7_i) >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_3_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_3_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_5_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_5_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_7_i) + 0x0u + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_7_i) >> 32) + 0x0u + cf_; t9 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * a_2_i) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_2_i) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_4_i) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_4_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_6_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_6_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t25 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_2_i * a_4_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_4_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_6_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_6_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_2_i * a_3_i) + t21; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_3_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_5_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_5_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_7_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_7_i) >> 32) + 0x0u + cf_; t26 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_3_i * a_5_i) + t8; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_5_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_7_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_7_i) >> 32) + 0x0u + cf_; t11 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_3_i * a_4_i) + t23; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_4_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_6_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_6_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_4_i * a_6_i) + t10; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_6_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_4_i * a_5_i) + t25; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_5_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_7_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_7_i) >> 32) + 0x0u + cf_; t28 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_5_i * a_7_i) + t12; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_7_i) >> 32) + 0x0u + cf_; t13 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_5_i * a_6_i) + t27; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_6_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_6_i * a_7_i) + t29; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_7_i) >> 32) + 0x0u + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t18 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t19 + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t20 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t21 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t22 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t23 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t8 + t24 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t25 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t26 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t27 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t28 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t29 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t30 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t45 = (uint32_t)w_;
    t46 = (uint32_t)(t17 << 1);
    t47 = (uint32_t)(((((uint64_t)t32 << 32) | t17) << 1) >> 32);
    t48 = (uint32_t)(((((uint64_t)t33 << 32) | t32) << 1) >> 32);
    t49 = (uint32_t)(((((uint64_t)t34 << 32) | t33) << 1) >> 32);
    t50 = (uint32_t)(((((uint64_t)t35 << 32) | t34) << 1) >> 32);
    t51 = (uint32_t)(((((uint64_t)t36 << 32) | t35) << 1) >> 32);
    t52 = (uint32_t)(((((uint64_t)t37 << 32) | t36) << 1) >> 32);
    t53 = (uint32_t)(((((uint64_t)t38 << 32) | t37) << 1) >> 32);
    t54 = (uint32_t)(((((uint64_t)t39 << 32) | t38) << 1) >> 32);
    t55 = (uint32_t)(((((uint64_t)t40 << 32) | t39) << 1) >> 32);
    t56 = (uint32_t)(((((uint64_t)t41 << 32) | t40) << 1) >> 32);
    t57 = (uint32_t)(((((uint64_t)t42 << 32) | t41) << 1) >> 32);
    t58 = (uint32_t)(((((uint64_t)t43 << 32) | t42) << 1) >> 32);
    t59 = (uint32_t)(((((uint64_t)t44 << 32) | t43) << 1) >> 32);
    t60 = (uint32_t)(((((uint64_t)t45 << 32) | t44) << 1) >> 32);
    w_ = (uint64_t)(uint32_t)(a_0_i * a_0_i) + 0x0u; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * a_0_i) >> 32) + t46 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_1_i) + t47 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_1_i) >> 32) + t48 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_2_i) + t49 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_2_i) >> 32) + t50 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_3_i) + t51 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_3_i) >> 32) + t52 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_4_i) + t53 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_4_i) >> 32) + t54 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_5_i) + t55 + cf_; t71 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_5_i) >> 32) + t56 + cf_; t72 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_6_i) + t57 + cf_; t73 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_6_i) >> 32) + t58 + cf_; t74 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_7_i) + t59 + cf_; t75 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_7_i) >> 32) + t60 + cf_; t76 = (uint32_t)w_;
    w_ = (uint64_t)t64 + t61; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t65 + t62 + cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t66 + t63 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t67 + t64 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t68 + t65 + cf_; t81 = (uint32_t)w_;
    t82 = (uint32_t)(t61 << 1);
    t83 = (uint32_t)(((((uint64_t)t62 << 32) | t61) << 1) >> 32);
    w_ = (uint64_t)t80 + t82; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t81 + t83 + cf_; t85 = (uint32_t)w_;
    w_ = (uint64_t)t85 - t61; t86 = (uint32_t)w_;
    w_ = (uint64_t)t77 + t61; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t78 + t62 + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t79 + t63 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t84 + t77 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t86 + t78 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t79 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t84 + cf_; t93 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t86 + cf_; t94 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t95 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t96 = (uint32_t)w_;
    w_ = (uint64_t)t89 + t61; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t90 + t62 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t91 + t63 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t92 + t77 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t93 + t78 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t94 + t79 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t95 + t84 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t96 + t86 + cf_; t104 = (uint32_t)w_;
    w_ = (uint64_t)t88 - t61; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t97 - t62 - cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t98 - t63 - cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t99 - t77 - cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t100 - t78 - cf_; t109 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t101 - t79 - cf_; t110 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t102 - t84 - cf_; t111 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t103 - t86 - cf_; t112 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t104 - 0x0u - cf_; t113 = (uint32_t)w_;
    w_ = (uint64_t)t64 + t61; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t65 + t62 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t66 + t63 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t67 + t87 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t68 + t105 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t69 + t106 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t70 + t107 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t71 + t108 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t72 + t109 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t73 + t110 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t74 + t111 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t75 + t112 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t76 + t113 + cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t127 = (uint32_t)w_;
    w_ = (uint64_t)t119 - 0xffffffffu; t128 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t120 - 0xffffffffu - cf_; t129 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t121 - 0xffffffffu - cf_; t130 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t122 - 0x0u - cf_; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t123 - 0x0u - cf_; t132 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t124 - 0x0u - cf_; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t125 - 0x1u - cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t126 - 0xffffffffu - cf_; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t127 - 0x0u - cf_; t136 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t137 = (uint32_t)w_;
    t138 = (uint32_t)(t128 ^ t119);
    t139 = (uint32_t)(t138 & t137);
    t140 = (uint32_t)(t139 ^ t128);
    t141 = (uint32_t)(t129 ^ t120);
    t142 = (uint32_t)(t141 & t137);
    t143 = (uint32_t)(t142 ^ t129);
    t144 = (uint32_t)(t130 ^ t121);
    t145 = (uint32_t)(t144 & t137);
    t146 = (uint32_t)(t145 ^ t130);
    t147 = (uint32_t)(t131 ^ t122);
    t148 = (uint32_t)(t147 & t137);
    t149 = (uint32_t)(t148 ^ t131);
    t150 = (uint32_t)(t132 ^ t123);
    t151 = (uint32_t)(t150 & t137);
    t152 = (uint32_t)(t151 ^ t132);
    t153 = (uint32_t)(t133 ^ t124);
    t154 = (uint32_t)(t153 & t137);
    t155 = (uint32_t)(t154 ^ t133);
    t156 = (uint32_t)(t134 ^ t125);
    t157 = (uint32_t)(t156 & t137);
    t158 = (uint32_t)(t157 ^ t134);
    t159 = (uint32_t)(t135 ^ t126);
    t160 = (uint32_t)(t159 & t137);
    t161 = (uint32_t)(t160 ^ t135);
    r[0] = t140;
    r[1] = t143;
    r[2] = t146;
    r[3] = t149;
    r[4] = t152;
    r[5] = t155;
    r[6] = t158;
    r[7] = t161;
#endif
  }

  // c = a*b for a small integer b (pseudo.py:705-728 / monty.py:876-978)
  static MAB_DEV void mli(uint32_t (&r)[8], const uint32_t (&a)[8], uint32_t b) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<118>;\n\t"
        "mul.lo.u32 t0, %8, %16;\n\t"
        "mul.hi.u32 t8, %8, %16;\n\t"
        "mul.lo.u32 t1, %9, %16;\n\t"
        "mul.hi.u32 t9, %9, %16;\n\t"
        "mul.lo.u32 t2, %10, %16;\n\t"
        "mul.hi.u32 t10, %10, %16;\n\t"
        "mul.lo.u32 t3, %11, %16;\n\t"
        "mul.hi.u32 t11, %11, %16;\n\t"
        "mul.lo.u32 t4, %12, %16;\n\t"
        "mul.hi.u32 t12, %12, %16;\n\t"
        "mul.lo.u32 t5, %13, %16;\n\t"
        "mul.hi.u32 t13, %13, %16;\n\t"
        "mul.lo.u32 t6, %14, %16;\n\t"
        "mul.hi.u32 t14, %14, %16;\n\t"
        "mul.lo.u32 t7, %15, %16;\n\t"
        "mul.hi.u32 t15, %15, %16;\n\t"
        "add.cc.u32 t16, t1, t8;\n\t"
        "addc.cc.u32 t17, t2, t9;\n\t"
        "addc.cc.u32 t18, t3, t10;\n\t"
        "addc.cc.u32 t19, t4, t11;\n\t"
        "addc.cc.u32 t20, t5, t12;\n\t"
        "addc.cc.u32 t21, t6, t13;\n\t"
        "addc.cc.u32 t22, t7, t14;\n\t"
        "addc.u32 t23, t15, 0x0;\n\t"
        "add.cc.u32 t24, t0, t23;\n\t"
        "addc.cc.u32 t25, t16, 0x0;\n\t"
        "addc.cc.u32 t26, t17, 0x0;\n\t"
        "addc.cc.u32 t27, t18, 0x0;\n\t"
        "addc.cc.u32 t28, t19, 0x0;\n\t"
        "addc.cc.u32 t29, t20, 0x0;\n\t"
        "addc.cc.u32 t30, t21, 0x0;\n\t"
        "addc.cc.u32 t31, t22, 0x0;\n\t"
        "addc.u32 t32, 0x0, 0x0;\n\t"
        "add.cc.u32 t33, t31, t23;\n\t"
        "addc.u32 t34, t32, 0x0;\n\t"
        "sub.cc.u32 t35, t27, t23;\n\t"
        "subc.cc.u32 t36, t28, 0x0;\n\t"
        "subc.cc.u32 t37, t29, 0x0;\n\t"
        "subc.cc.u32 t38, t30, 0x0;\n\t"
        "subc.cc.u32 t39, t33, 0x0;\n\t"
        "subc.u32 t40, t34, 0x0;\n\t"
        "sub.cc.u32 t41, t38, t23;\n\t"
        "subc.cc.u32 t42, t39, 0x0;\n\t"
        "subc.u32 t43, t40, 0x0;\n\t"
        "add.cc.u32 t44, t24, t43;\n\t"
        "addc.cc.u32 t45, t25, 0x0;\n\t"
        "addc.cc.u32 t46, t26, 0x0;\n\t"
        "addc.cc.u32 t47, t35, 0x0;\n\t"
        "addc.cc.u32 t48, t36, 0x0;\n\t"
        "addc.cc.u32 t49, t37, 0x0;\n\t"
        "addc.cc.u32 t50, t41, 0x0;\n\t"
        "addc.cc.u32 t51, t42, 0x0;\n\t"
        "addc.u32 t52, 0x0, 0x0;\n\t"
        "add.cc.u32 t53, t51, t43;\n\t"
        "addc.u32 t54, t52, 0x0;\n\t"
        "sub.cc.u32 t55, t47, t43;\n\t"
        "subc.cc.u32 t56, t48, 0x0;\n\t"
        "subc.cc.u32 t57, t49, 0x0;\n\t"
        "subc.cc.u32 t58, t50, 0x0;\n\t"
        "subc.cc.u32 t59, t53, 0x0;\n\t"
        "subc.u32 t60, t54, 0x0;\n\t"
        "sub.cc.u32 t61, t58, t43;\n\t"
        "subc.cc.u32 t62, t59, 0x0;\n\t"
        "subc.u32 t63, t60, 0x0;\n\t"
        "add.cc.u32 t64, t44, t63;\n\t"
        "addc.cc.u32 t65, t45, 0x0;\n\t"
        "addc.cc.u32 t66, t46, 0x0;\n\t"
        "addc.cc.u32 t67, t55, 0x0;\n\t"
        "addc.cc.u32 t68, t56, 0x0;\n\t"
        "addc.cc.u32 t69, t57, 0x0;\n\t"
        "addc.cc.u32 t70, t61, 0x0;\n\t"
        "addc.cc.u32 t71, t62, 0x0;\n\t"
        "addc.u32 t72, 0x0, 0x0;\n\t"
        "add.cc.u32 t73, t71, t63;\n\t"
        "addc.u32 t74, t72, 0x0;\n\t"
        "sub.cc.u32 t75, t67, t63;\n\t"
        "subc.cc.u32 t76, t68, 0x0;\n\t"
        "subc.cc.u32 t77, t69, 0x0;\n\t"
        "subc.cc.u32 t78, t70, 0x0;\n\t"
        "subc.cc.u32 t79, t73, 0x0;\n\t"
        "subc.u32 t80, t74, 0x0;\n\t"
        "sub.cc.u32 t81, t78, t63;\n\t"
        "subc.cc.u32 t82, t79, 0x0;\n\t"
        "subc.u32 t83, t80, 0x0;\n\t"
        "sub.cc.u32 t84, t64, 0xffffffff;\n\t"
        "subc.cc.u32 t85, t65, 0xffffffff;\n\t"
        "subc.cc.u32 t86, t66, 0xffffffff;\n\t"
        "subc.cc.u32 t87, t75, 0x0;\n\t"
        "subc.cc.u32 t88, t76, 0x0;\n\t"
        "subc.cc.u32 t89, t77, 0x0;\n\t"
        "subc.cc.u32 t90, t81, 0x1;\n\t"
        "subc.cc.u32 t91, t82, 0xffffffff;\n\t"
        "subc.cc.u32 t92, t83, 0x0;\n\t"
        "subc.u32 t93, 0x0, 0x0;\n\t"
        "xor.b32 t94, t84, t64;\n\t"
        "and.b32 t95, t94, t93;\n\t"
        "xor.b32 t96, t95, t84;\n\t"
        "xor.b32 t97, t85, t65;\n\t"
        "and.b32 t98, t97, t93;\n\t"
        "xor.b32 t99, t98, t85;\n\t"
        "xor.b32 t100, t86, t66;\n\t"
        "and.b32 t101, t100, t93;\n\t"
        "xor.b32 t102, t101, t86;\n\t"
        "xor.b32 t103, t87, t75;\n\t"
        "and.b32 t104, t103, t93;\n\t"
        "xor.b32 t105, t104, t87;\n\t"
        "xor.b32 t106, t88, t76;\n\t"
        "and.b32 t107, t106, t93;\n\t"
        "xor.b32 t108, t107, t88;\n\t"
        "xor.b32 t109, t89, t77;\n\t"
        "and.b32 t110, t109, t93;\n\t"
        "xor.b32 t111, t110, t89;\n\t"
        "xor.b32 t112, t90, t81;\n\t"
        "and.b32 t113, t112, t93;\n\t"
        "xor.b32 t114, t113, t90;\n\t"
        "xor.b32 t115, t91, t82;\n\t"
        "and.b32 t116, t115, t93;\n\t"
        "xor.b32 t117, t116, t91;\n\t"
        "mov.u32 %0, t96;\n\t"
        "mov.u32 %1, t99;\n\t"
        "mov.u32 %2, t102;\n\t"
        "mov.u32 %3, t105;\n\t"
        "mov.u32 %4, t108;\n\t"
        "mov.u32 %5, t111;\n\t"
        "mov.u32 %6, t114;\n\t"
        "mov.u32 %7, t117;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t b_i = b;
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t0 = (uint32_t)((uint32_t)(a_0_i * b_i));
    t8 = (uint32_t)(((uint64_t)a_0_i * b_i) >> 32);
    t1 = (uint32_t)((uint32_t)(a_1_i * b_i));
    t9 = (uint32_t)(((uint64_t)a_1_i * b_i) >> 32);
    t2 = (uint32_t)((uint32_t)(a_2_i * b_i));
    t10 = (uint32_t)(((uint64_t)a_2_i * b_i) >> 32);
    t3 = (uint32_t)((uint32_t)(a_3_i * b_i));
    t11 = (uint32_t)(((uint64_t)a_3_i * b_i) >> 32);
    t4 = (uint32_t)((uint32_t)(a_4_i * b_i));
    t12 = (uint32_t)(((uint64_t)a_4_i * b_i) >> 32);
    t5 = (uint32_t)((uint32_t)(a_5_i * b_i));
    t13 = (uint32_t)(((uint64_t)a_5_i * b_i) >> 32);
    t6 = (uint32_t)((uint32_t)(a_6_i * b_i));
    t14 = (uint32_t)(((uint64_t)a_6_i * b_i) >> 32);
    t7 = (uint32_t)((uint32_t)(a_7_i * b_i));
    t15 = (uint32_t)(((uint64_t)a_7_i * b_i) >> 32);
    w_ = (uint64_t)t1 + t8; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t9 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t10 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t11 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t12 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t13 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t14 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + 0x0u + cf_; t23 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t23; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t16 + 0x0u + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t17 + 0x0u + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t18 + 0x0u + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t19 + 0x0u + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t20 + 0x0u + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t21 + 0x0u + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t22 + 0x0u + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)t31 + t23; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t34 = (uint32_t)w_;
    w_ = (uint64_t)t27 - t23; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t28 - 0x0u - cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t29 - 0x0u - cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t30 - 0x0u - cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t33 - 0x0u - cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t34 - 0x0u - cf_; t40 = (uint32_t)w_;
    w_ = (uint64_t)t38 - t23; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t39 - 0x0u - cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t40 - 0x0u - cf_; t43 = (uint32_t)w_;
    w_ = (uint64_t)t24 + t43; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t25 + 0x0u + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t26 + 0x0u + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t35 + 0x0u + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t36 + 0x0u + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t37 + 0x0u + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t41 + 0x0u + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t42 + 0x0u + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t52 = (uint32_t)w_;
    w_ = (uint64_t)t51 + t43; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t52 + 0x0u + cf_; t54 = (uint32_t)w_;
    w_ = (uint64_t)t47 - t43; t55 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t48 - 0x0u - cf_; t56 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t49 - 0x0u - cf_; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t50 - 0x0u - cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t53 - 0x0u - cf_; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t54 - 0x0u - cf_; t60 = (uint32_t)w_;
    w_ = (uint64_t)t58 - t43; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t59 - 0x0u - cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t60 - 0x0u - cf_; t63 = (uint32_t)w_;
    w_ = (uint64_t)t44 + t63; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t45 + 0x0u + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t46 + 0x0u + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t55 + 0x0u + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t56 + 0x0u + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t57 + 0x0u + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t61 + 0x0u + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t62 + 0x0u + cf_; t71 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t72 = (uint32_t)w_;
    w_ = (uint64_t)t71 + t63; t73 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t72 + 0x0u + cf_; t74 = (uint32_t)w_;
    w_ = (uint64_t)t67 - t63; t75 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t68 - 0x0u - cf_; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t69 - 0x0u - cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t70 - 0x0u - cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t73 - 0x0u - cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t74 - 0x0u - cf_; t80 = (uint32_t)w_;
    w_ = (uint64_t)t78 - t63; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t79 - 0x0u - cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t80 - 0x0u - cf_; t83 = (uint32_t)w_;
    w_ = (uint64_t)t64 - 0xffffffffu; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t65 - 0xffffffffu - cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t66 - 0xffffffffu - cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t75 - 0x0u - cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t76 - 0x0u - cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t77 - 0x0u - cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t81 - 0x1u - cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t82 - 0xffffffffu - cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t83 - 0x0u - cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t93 = (uint32_t)w_;
    t94 = (uint32_t)(t84 ^ t64);
    t95 = (uint32_t)(t94 & t93);
    t96 = (uint32_t)(t95 ^ t84);
    t97 = (uint32_t)(t85 ^ t65);
    t98 = (uint32_t)(t97 & t93);
    t99 = (uint32_t)(t98 ^ t85);
    t100 = (uint32_t)(t86 ^ t66);
    t101 = (uint32_t)(t100 & t93);
    t102 = (uint32_t)(t101 ^ t86);
    t103 = (uint32_t)(t87 ^ t75);
    t104 = (uint32_t)(t103 & t93);
    t105 = (uint32_t)(t104 ^ t87);
    t106 = (uint32_t)(t88 ^ t76);
    t107 = (uint32_t)(t106 & t93);
    t108 = (uint32_t)(t107 ^ t88);
    t109 = (uint32_t)(t89 ^ t77);
    t110 = (uint32_t)(t109 & t93);
    t111 = (uint32_t)(t110 ^ t89);
    t112 = (uint32_t)(t90 ^ t81);
    t113 = (uint32_t)(t112 & t93);
    t114 = (uint32_t)(t113 ^ t90);
    t115 = (uint32_t)(t91 ^ t82);
    t116 = (uint32_t)(t115 & t93);
    t117 = (uint32_t)(t116 ^ t91);
    r[0] = t96;
    r[1] = t99;
    r[2] = t102;
    r[3] = t105;
    r[4] = t108;
    r[5] = t111;
    r[6] = t114;
    r[7] = t117;
#endif
  }

  // r = a*b + c, small integer b: modmli + modadd fused (rfc7748.c:209,212)
  static MAB_DEV void mla(uint32_t (&r)[8], const uint32_t (&a)[8], uint32_t b, const uint32_t (&c)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<119>;\n\t"
        "mad.lo.cc.u32 t0, %8, %24, %16;\n\t"
        "madc.hi.cc.u32 t1, %8, %24, %17;\n\t"
        "madc.lo.cc.u32 t2, %10, %24, %18;\n\t"
        "madc.hi.cc.u32 t3, %10, %24, %19;\n\t"
        "madc.lo.cc.u32 t4, %12, %24, %20;\n\t"
        "madc.hi.cc.u32 t5, %12, %24, %21;\n\t"
        "madc.lo.cc.u32 t6, %14, %24, %22;\n\t"
        "madc.hi.cc.u32 t7, %14, %24, %23;\n\t"
        "addc.u32 t8, 0x0, 0x0;\n\t"
        "mul.lo.u32 t9, %9, %24;\n\t"
        "mul.hi.u32 t10, %9, %24;\n\t"
        "mul.lo.u32 t11, %11, %24;\n\t"
        "mul.hi.u32 t12, %11, %24;\n\t"
        "mul.lo.u32 t13, %13, %24;\n\t"
        "mul.hi.u32 t14, %13, %24;\n\t"
        "mul.lo.u32 t15, %15, %24;\n\t"
        "mul.hi.u32 t16, %15, %24;\n\t"
        "add.cc.u32 t17, t1, t9;\n\t"
        "addc.cc.u32 t18, t2, t10;\n\t"
        "addc.cc.u32 t19, t3, t11;\n\t"
        "addc.cc.u32 t20, t4, t12;\n\t"
        "addc.cc.u32 t21, t5, t13;\n\t"
        "addc.cc.u32 t22, t6, t14;\n\t"
        "addc.cc.u32 t23, t7, t15;\n\t"
        "addc.u32 t24, t8, t16;\n\t"
        "add.cc.u32 t25, t0, t24;\n\t"
        "addc.cc.u32 t26, t17, 0x0;\n\t"
        "addc.cc.u32 t27, t18, 0x0;\n\t"
        "addc.cc.u32 t28, t19, 0x0;\n\t"
        "addc.cc.u32 t29, t20, 0x0;\n\t"
        "addc.cc.u32 t30, t21, 0x0;\n\t"
        "addc.cc.u32 t31, t22, 0x0;\n\t"
        "addc.cc.u32 t32, t23, 0x0;\n\t"
        "addc.u32 t33, 0x0, 0x0;\n\t"
        "add.cc.u32 t34, t32, t24;\n\t"
        "addc.u32 t35, t33, 0x0;\n\t"
        "sub.cc.u32 t36, t28, t24;\n\t"
        "subc.cc.u32 t37, t29, 0x0;\n\t"
        "subc.cc.u32 t38, t30, 0x0;\n\t"
        "subc.cc.u32 t39, t31, 0x0;\n\t"
        "subc.cc.u32 t40, t34, 0x0;\n\t"
        "subc.u32 t41, t35, 0x0;\n\t"
        "sub.cc.u32 t42, t39, t24;\n\t"
        "subc.cc.u32 t43, t40, 0x0;\n\t"
        "subc.u32 t44, t41, 0x0;\n\t"
        "add.cc.u32 t45, t25, t44;\n\t"
        "addc.cc.u32 t46, t26, 0x0;\n\t"
        "addc.cc.u32 t47, t27, 0x0;\n\t"
        "addc.cc.u32 t48, t36, 0x0;\n\t"
        "addc.cc.u32 t49, t37, 0x0;\n\t"
        "addc.cc.u32 t50, t38, 0x0;\n\t"
        "addc.cc.u32 t51, t42, 0x0;\n\t"
        "addc.cc.u32 t52, t43, 0x0;\n\t"
        "addc.u32 t53, 0x0, 0x0;\n\t"
        "add.cc.u32 t54, t52, t44;\n\t"
        "addc.u32 t55, t53, 0x0;\n\t"
        "sub.cc.u32 t56, t48, t44;\n\t"
        "subc.cc.u32 t57, t49, 0x0;\n\t"
        "subc.cc.u32 t58, t50, 0x0;\n\t"
        "subc.cc.u32 t59, t51, 0x0;\n\t"
        "subc.cc.u32 t60, t54, 0x0;\n\t"
        "subc.u32 t61, t55, 0x0;\n\t"
        "sub.cc.u32 t62, t59, t44;\n\t"
        "subc.cc.u32 t63, t60, 0x0;\n\t"
        "subc.u32 t64, t61, 0x0;\n\t"
        "add.cc.u32 t65, t45, t64;\n\t"
        "addc.cc.u32 t66, t46, 0x0;\n\t"
        "addc.cc.u32 t67, t47, 0x0;\n\t"
        "addc.cc.u32 t68, t56, 0x0;\n\t"
        "addc.cc.u32 t69, t57, 0x0;\n\t"
        "addc.cc.u32 t70, t58, 0x0;\n\t"
        "addc.cc.u32 t71, t62, 0x0;\n\t"
        "addc.cc.u32 t72, t63, 0x0;\n\t"
        "addc.u32 t73, 0x0, 0x0;\n\t"
        "add.cc.u32 t74, t72, t64;\n\t"
        "addc.u32 t75, t73, 0x0;\n\t"
        "sub.cc.u32 t76, t68, t64;\n\t"
        "subc.cc.u32 t77, t69, 0x0;\n\t"
        "subc.cc.u32 t78, t70, 0x0;\n\t"
        "subc.cc.u32 t79, t71, 0x0;\n\t"
        "subc.cc.u32 t80, t74, 0x0;\n\t"
        "subc.u32 t81, t75, 0x0;\n\t"
        "sub.cc.u32 t82, t79, t64;\n\t"
        "subc.cc.u32 t83, t80, 0x0;\n\t"
        "subc.u32 t84, t81, 0x0;\n\t"
        "sub.cc.u32 t85, t65, 0xffffffff;\n\t"
        "subc.cc.u32 t86, t66, 0xffffffff;\n\t"
        "subc.cc.u32 t87, t67, 0xffffffff;\n\t"
        "subc.cc.u32 t88, t76, 0x0;\n\t"
        "subc.cc.u32 t89, t77, 0x0;\n\t"
        "subc.cc.u32 t90, t78, 0x0;\n\t"
        "subc.cc.u32 t91, t82, 0x1;\n\t"
        "subc.cc.u32 t92, t83, 0xffffffff;\n\t"
        "subc.cc.u32 t93, t84, 0x0;\n\t"
        "subc.u32 t94, 0x0, 0x0;\n\t"
        "xor.b32 t95, t85, t65;\n\t"
        "and.b32 t96, t95, t94;\n\t"
        "xor.b32 t97, t96, t85;\n\t"
        "xor.b32 t98, t86, t66;\n\t"
        "and.b32 t99, t98, t94;\n\t"
        "xor.b32 t100, t99, t86;\n\t"
        "xor.b32 t101, t87, t67;\n\t"
        "and.b32 t102, t101, t94;\n\t"
        "xor.b32 t103, t102, t87;\n\t"
        "xor.b32 t104, t88, t76;\n\t"
        "and.b32 t105, t104, t94;\n\t"
        "xor.b32 t106, t105, t88;\n\t"
        "xor.b32 t107, t89, t77;\n\t"
        "and.b32 t108, t107, t94;\n\t"
        "xor.b32 t109, t108, t89;\n\t"
        "xor.b32 t110, t90, t78;\n\t"
        "and.b32 t111, t110, t94;\n\t"
        "xor.b32 t112, t111, t90;\n\t"
        "xor.b32 t113, t91, t82;\n\t"
        "and.b32 t114, t113, t94;\n\t"
        "xor.b32 t115, t114, t91;\n\t"
        "xor.b32 t116, t92, t83;\n\t"
        "and.b32 t117, t116, t94;\n\t"
        "xor.b32 t118, t117, t92;\n\t"
        "mov.u32 %0, t97;\n\t"
        "mov.u32 %1, t100;\n\t"
        "mov.u32 %2, t103;\n\t"
        "mov.u32 %3, t106;\n\t"
        "mov.u32 %4, t109;\n\t"
        "mov.u32 %5, t112;\n\t"
        "mov.u32 %6, t115;\n\t"
        "mov.u32 %7, t118;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]), "r"(b));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t c_0_i = c[0];
    const uint32_t c_1_i = c[1];
    const uint32_t c_2_i = c[2];
    const uint32_t c_3_i = c[3];
    const uint32_t c_4_i = c[4];
    const uint32_t c_5_i = c[5];
    const uint32_t c_6_i = c[6];
    const uint32_t c_7_i = c[7];
    const uint32_t b_i = b;
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_i) + c_0_i; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_i) >> 32) + c_1_i + cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_i) + c_2_i + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_i) >> 32) + c_3_i + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_i) + c_4_i + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_i) >> 32) + c_5_i + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_i) + c_6_i + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_i) >> 32) + c_7_i + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t8 = (uint32_t)w_;
    t9 = (uint32_t)((uint32_t)(a_1_i * b_i));
    t10 = (uint32_t)(((uint64_t)a_1_i * b_i) >> 32);
    t11 = (uint32_t)((uint32_t)(a_3_i * b_i));
    t12 = (uint32_t)(((uint64_t)a_3_i * b_i) >> 32);
    t13 = (uint32_t)((uint32_t)(a_5_i * b_i));
    t14 = (uint32_t)(((uint64_t)a_5_i * b_i) >> 32);
    t15 = (uint32_t)((uint32_t)(a_7_i * b_i));
    t16 = (uint32_t)(((uint64_t)a_7_i * b_i) >> 32);
    w_ = (uint64_t)t1 + t9; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t10 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t11 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t12 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t13 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t14 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t15 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t8 + t16 + cf_; t24 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t24; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t17 + 0x0u + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t18 + 0x0u + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t19 + 0x0u + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t20 + 0x0u + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t21 + 0x0u + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t22 + 0x0u + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t23 + 0x0u + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t33 = (uint32_t)w_;
    w_ = (uint64_t)t32 + t24; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t33 + 0x0u + cf_; t35 = (uint32_t)w_;
    w_ = (uint64_t)t28 - t24; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t29 - 0x0u - cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t30 - 0x0u - cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t31 - 0x0u - cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t34 - 0x0u - cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t35 - 0x0u - cf_; t41 = (uint32_t)w_;
    w_ = (uint64_t)t39 - t24; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t40 - 0x0u - cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t41 - 0x0u - cf_; t44 = (uint32_t)w_;
    w_ = (uint64_t)t25 + t44; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t26 + 0x0u + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t27 + 0x0u + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t36 + 0x0u + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t37 + 0x0u + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t38 + 0x0u + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t42 + 0x0u + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t43 + 0x0u + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t53 = (uint32_t)w_;
    w_ = (uint64_t)t52 + t44; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t53 + 0x0u + cf_; t55 = (uint32_t)w_;
    w_ = (uint64_t)t48 - t44; t56 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t49 - 0x0u - cf_; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t50 - 0x0u - cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t51 - 0x0u - cf_; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t54 - 0x0u - cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t55 - 0x0u - cf_; t61 = (uint32_t)w_;
    w_ = (uint64_t)t59 - t44; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t60 - 0x0u - cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t61 - 0x0u - cf_; t64 = (uint32_t)w_;
    w_ = (uint64_t)t45 + t64; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t46 + 0x0u + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t47 + 0x0u + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t56 + 0x0u + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t57 + 0x0u + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t58 + 0x0u + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t62 + 0x0u + cf_; t71 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + 0x0u + cf_; t72 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t73 = (uint32_t)w_;
    w_ = (uint64_t)t72 + t64; t74 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t73 + 0x0u + cf_; t75 = (uint32_t)w_;
    w_ = (uint64_t)t68 - t64; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t69 - 0x0u - cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t70 - 0x0u - cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t71 - 0x0u - cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t74 - 0x0u - cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t75 - 0x0u - cf_; t81 = (uint32_t)w_;
    w_ = (uint64_t)t79 - t64; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t80 - 0x0u - cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t81 - 0x0u - cf_; t84 = (uint32_t)w_;
    w_ = (uint64_t)t65 - 0xffffffffu; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t66 - 0xffffffffu - cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t67 - 0xffffffffu - cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t76 - 0x0u - cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t77 - 0x0u - cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t78 - 0x0u - cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t82 - 0x1u - cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t83 - 0xffffffffu - cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t84 - 0x0u - cf_; t93 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t94 = (uint32_t)w_;
    t95 = (uint32_t)(t85 ^ t65);
    t96 = (uint32_t)(t95 & t94);
    t97 = (uint32_t)(t96 ^ t85);
    t98 = (uint32_t)(t86 ^ t66);
    t99 = (uint32_t)(t98 & t94);
    t100 = (uint32_t)(t99 ^ t86);
    t101 = (uint32_t)(t87 ^ t67);
    t102 = (uint32_t)(t101 & t94);
    t103 = (uint32_t)(t102 ^ t87);
    t104 = (uint32_t)(t88 ^ t76);
    t105 = (uint32_t)(t104 & t94);
    t106 = (uint32_t)(t105 ^ t88);
    t107 = (uint32_t)(t89 ^ t77);
    t108 = (uint32_t)(t107 & t94);
    t109 = (uint32_t)(t108 ^ t89);
    t110 = (uint32_t)(t90 ^ t78);
    t111 = (uint32_t)(t110 & t94);
    t112 = (uint32_t)(t111 ^ t90);
    t113 = (uint32_t)(t91 ^ t82);
    t114 = (uint32_t)(t113 & t94);
    t115 = (uint32_t)(t114 ^ t91);
    t116 = (uint32_t)(t92 ^ t83);
    t117 = (uint32_t)(t116 & t94);
    t118 = (uint32_t)(t117 ^ t92);
    r[0] = t97;
    r[1] = t100;
    r[2] = t103;
    r[3] = t106;
    r[4] = t109;
    r[5] = t112;
    r[6] = t115;
    r[7] = t118;
#endif
  }

  // n = a+b (pseudo.py:286-304)
  static MAB_DEV void add(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<43>;\n\t"
        "add.cc.u32 t0, %8, %16;\n\t"
        "addc.cc.u32 t1, %9, %17;\n\t"
        "addc.cc.u32 t2, %10, %18;\n\t"
        "addc.cc.u32 t3, %11, %19;\n\t"
        "addc.cc.u32 t4, %12, %20;\n\t"
        "addc.cc.u32 t5, %13, %21;\n\t"
        "addc.cc.u32 t6, %14, %22;\n\t"
        "addc.cc.u32 t7, %15, %23;\n\t"
        "madc.lo.u32 t8, 0, 0, 0;\n\t"
        "sub.cc.u32 t9, t0, 0xffffffff;\n\t"
        "subc.cc.u32 t10, t1, 0xffffffff;\n\t"
        "subc.cc.u32 t11, t2, 0xffffffff;\n\t"
        "subc.cc.u32 t12, t3, 0x0;\n\t"
        "subc.cc.u32 t13, t4, 0x0;\n\t"
        "subc.cc.u32 t14, t5, 0x0;\n\t"
        "subc.cc.u32 t15, t6, 0x1;\n\t"
        "subc.cc.u32 t16, t7, 0xffffffff;\n\t"
        "subc.cc.u32 t17, t8, 0x0;\n\t"
        "subc.u32 t18, 0x0, 0x0;\n\t"
        "xor.b32 t19, t9, t0;\n\t"
        "and.b32 t20, t19, t18;\n\t"
        "xor.b32 t21, t20, t9;\n\t"
        "xor.b32 t22, t10, t1;\n\t"
        "and.b32 t23, t22, t18;\n\t"
        "xor.b32 t24, t23, t10;\n\t"
        "xor.b32 t25, t11, t2;\n\t"
        "and.b32 t26, t25, t18;\n\t"
        "xor.b32 t27, t26, t11;\n\t"
        "xor.b32 t28, t12, t3;\n\t"
        "and.b32 t29, t28, t18;\n\t"
        "xor.b32 t30, t29, t12;\n\t"
        "xor.b32 t31, t13, t4;\n\t"
        "and.b32 t32, t31, t18;\n\t"
        "xor.b32 t33, t32, t13;\n\t"
        "xor.b32 t34, t14, t5;\n\t"
        "and.b32 t35, t34, t18;\n\t"
        "xor.b32 t36, t35, t14;\n\t"
        "xor.b32 t37, t15, t6;\n\t"
        "and.b32 t38, t37, t18;\n\t"
        "xor.b32 t39, t38, t15;\n\t"
        "xor.b32 t40, t16, t7;\n\t"
        "and.b32 t41, t40, t18;\n\t"
        "xor.b32 t42, t41, t16;\n\t"
        "mov.u32 %0, t21;\n\t"
        "mov.u32 %1, t24;\n\t"
        "mov.u32 %2, t27;\n\t"
        "mov.u32 %3, t30;\n\t"
        "mov.u32 %4, t33;\n\t"
        "mov.u32 %5, t36;\n\t"
        "mov.u32 %6, t39;\n\t"
        "mov.u32 %7, t42;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t b_0_i = b[0];
    const uint32_t b_1_i = b[1];
    const uint32_t b_2_i = b[2];
    const uint32_t b_3_i = b[3];
    const uint32_t b_4_i = b[4];
    const uint32_t b_5_i = b[5];
    const uint32_t b_6_i = b[6];
    const uint32_t b_7_i = b[7];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    w_ = (uint64_t)a_0_i + b_0_i; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_1_i + b_1_i + cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_2_i + b_2_i + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_3_i + b_3_i + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_4_i + b_4_i + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_5_i + b_5_i + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_6_i + b_6_i + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_7_i + b_7_i + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t8 = (uint32_t)w_;
    w_ = (uint64_t)t0 - 0xffffffffu; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t1 - 0xffffffffu - cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t2 - 0xffffffffu - cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t3 - 0x0u - cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t4 - 0x0u - cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t5 - 0x0u - cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t6 - 0x1u - cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t7 - 0xffffffffu - cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t8 - 0x0u - cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t18 = (uint32_t)w_;
    t19 = (uint32_t)(t9 ^ t0);
    t20 = (uint32_t)(t19 & t18);
    t21 = (uint32_t)(t20 ^ t9);
    t22 = (uint32_t)(t10 ^ t1);
    t23 = (uint32_t)(t22 & t18);
    t24 = (uint32_t)(t23 ^ t10);
    t25 = (uint32_t)(t11 ^ t2);
    t26 = (uint32_t)(t25 & t18);
    t27 = (uint32_t)(t26 ^ t11);
    t28 = (uint32_t)(t12 ^ t3);
    t29 = (uint32_t)(t28 & t18);
    t30 = (uint32_t)(t29 ^ t12);
    t31 = (uint32_t)(t13 ^ t4);
    t32 = (uint32_t)(t31 & t18);
    t33 = (uint32_t)(t32 ^ t13);
    t34 = (uint32_t)(t14 ^ t5);
    t35 = (uint32_t)(t34 & t18);
    t36 = (uint32_t)(t35 ^ t14);
    t37 = (uint32_t)(t15 ^ t6);
    t38 = (uint32_t)(t37 & t18);
    t39 = (uint32_t)(t38 ^ t15);
    t40 = (uint32_t)(t16 ^ t7);
    t41 = (uint32_t)(t40 & t18);
    t42 = (uint32_t)(t41 ^ t16);
    r[0] = t21;
    r[1] = t24;
    r[2] = t27;
    r[3] = t30;
    r[4] = t33;
    r[5] = t36;
    r[6] = t39;
    r[7] = t42;
#endif
  }

  // n = a-b (pseudo.py:307-326)
  static MAB_DEV void sub(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<18>;\n\t"
        "sub.cc.u32 t0, %8, %16;\n\t"
        "subc.cc.u32 t1, %9, %17;\n\t"
        "subc.cc.u32 t2, %10, %18;\n\t"
        "subc.cc.u32 t3, %11, %19;\n\t"
        "subc.cc.u32 t4, %12, %20;\n\t"
        "subc.cc.u32 t5, %13, %21;\n\t"
        "subc.cc.u32 t6, %14, %22;\n\t"
        "subc.cc.u32 t7, %15, %23;\n\t"
        "subc.u32 t8, 0x0, 0x0;\n\t"
        "and.b32 t9, t8, 0x1;\n\t"
        "add.cc.u32 t10, t0, t8;\n\t"
        "addc.cc.u32 t11, t1, t8;\n\t"
        "addc.cc.u32 t12, t2, t8;\n\t"
        "addc.cc.u32 t13, t3, 0x0;\n\t"
        "addc.cc.u32 t14, t4, 0x0;\n\t"
        "addc.cc.u32 t15, t5, 0x0;\n\t"
        "addc.cc.u32 t16, t6, t9;\n\t"
        "addc.u32 t17, t7, t8;\n\t"
        "mov.u32 %0, t10;\n\t"
        "mov.u32 %1, t11;\n\t"
        "mov.u32 %2, t12;\n\t"
        "mov.u32 %3, t13;\n\t"
        "mov.u32 %4, t14;\n\t"
        "mov.u32 %5, t15;\n\t"
        "mov.u32 %6, t16;\n\t"
        "mov.u32 %7, t17;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t b_0_i = b[0];
    const uint32_t b_1_i = b[1];
    const uint32_t b_2_i = b[2];
    const uint32_t b_3_i = b[3];
    const uint32_t b_4_i = b[4];
    const uint32_t b_5_i = b[5];
    const uint32_t b_6_i = b[6];
    const uint32_t b_7_i = b[7];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    w_ = (uint64_t)a_0_i - b_0_i; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_1_i - b_1_i - cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_2_i - b_2_i - cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_3_i - b_3_i - cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_4_i - b_4_i - cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_5_i - b_5_i - cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_6_i - b_6_i - cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_7_i - b_7_i - cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t8 = (uint32_t)w_;
    t9 = (uint32_t)(t8 & 0x1u);
    w_ = (uint64_t)t0 + t8; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t1 + t8 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t8 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + 0x0u + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + 0x0u + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + 0x0u + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t9 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t8 + cf_; t17 = (uint32_t)w_;
    r[0] = t10;
    r[1] = t11;
    r[2] = t12;
    r[3] = t13;
    r[4] = t14;
    r[5] = t15;
    r[6] = t16;
    r[7] = t17;
#endif
  }

  // no spare bit above Nbits in this plan: the product-operand forms are the general ones
  static constexpr bool TIGHT = false;
  static MAB_DEV void add_tt(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { add(r, a, b); }
  static MAB_DEV void sub_tt(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { sub(r, a, b); }

  // Weakly reduced products for chains (modpro, modnsqr): operands and results are any representative
  // below 2^256; the last step looks at the carry word only.  A chain ends with canon(), which restores
  // the [0, p) invariant that every other function of this field keeps.
  static constexpr bool WEAK = true;
  static MAB_DEV void mul_w(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<108>;\n\t"
        "mul.lo.u32 t0, %8, %16;\n\t"
        "mul.hi.u32 t1, %8, %16;\n\t"
        "mul.lo.u32 t2, %10, %16;\n\t"
        "mul.hi.u32 t3, %10, %16;\n\t"
        "mul.lo.u32 t4, %12, %16;\n\t"
        "mul.hi.u32 t5, %12, %16;\n\t"
        "mul.lo.u32 t6, %14, %16;\n\t"
        "mul.hi.u32 t7, %14, %16;\n\t"
        "mul.lo.u32 t17, %9, %16;\n\t"
        "mul.hi.u32 t18, %9, %16;\n\t"
        "mul.lo.u32 t19, %11, %16;\n\t"
        "mul.hi.u32 t20, %11, %16;\n\t"
        "mul.lo.u32 t21, %13, %16;\n\t"
        "mul.hi.u32 t22, %13, %16;\n\t"
        "mul.lo.u32 t23, %15, %16;\n\t"
        "mul.hi.u32 t24, %15, %16;\n\t"
        "mad.lo.cc.u32 t2, %9, %17, t2;\n\t"
        "madc.hi.cc.u32 t3, %9, %17, t3;\n\t"
        "madc.lo.cc.u32 t4, %11, %17, t4;\n\t"
        "madc.hi.cc.u32 t5, %11, %17, t5;\n\t"
        "madc.lo.cc.u32 t6, %13, %17, t6;\n\t"
        "madc.hi.cc.u32 t7, %13, %17, t7;\n\t"
        "madc.lo.cc.u32 t8, %15, %17, 0x0;\n\t"
        "madc.hi.u32 t9, %15, %17, 0x0;\n\t"
        "mad.lo.cc.u32 t17, %8, %17, t17;\n\t"
        "madc.hi.cc.u32 t18, %8, %17, t18;\n\t"
        "madc.lo.cc.u32 t19, %10, %17, t19;\n\t"
        "madc.hi.cc.u32 t20, %10, %17, t20;\n\t"
        "madc.lo.cc.u32 t21, %12, %17, t21;\n\t"
        "madc.hi.cc.u32 t22, %12, %17, t22;\n\t"
        "madc.lo.cc.u32 t23, %14, %17, t23;\n\t"
        "madc.hi.cc.u32 t24, %14, %17, t24;\n\t"
        "addc.u32 t25, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t2, %8, %18, t2;\n\t"
        "madc.hi.cc.u32 t3, %8, %18, t3;\n\t"
        "madc.lo.cc.u32 t4, %10, %18, t4;\n\t"
        "madc.hi.cc.u32 t5, %10, %18, t5;\n\t"
        "madc.lo.cc.u32 t6, %12, %18, t6;\n\t"
        "madc.hi.cc.u32 t7, %12, %18, t7;\n\t"
        "madc.lo.cc.u32 t8, %14, %18, t8;\n\t"
        "madc.hi.cc.u32 t9, %14, %18, t9;\n\t"
        "addc.u32 t10, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t19, %9, %18, t19;\n\t"
        "madc.hi.cc.u32 t20, %9, %18, t20;\n\t"
        "madc.lo.cc.u32 t21, %11, %18, t21;\n\t"
        "madc.hi.cc.u32 t22, %11, %18, t22;\n\t"
        "madc.lo.cc.u32 t23, %13, %18, t23;\n\t"
        "madc.hi.cc.u32 t24, %13, %18, t24;\n\t"
        "madc.lo.cc.u32 t25, %15, %18, t25;\n\t"
        "madc.hi.u32 t26, %15, %18, 0x0;\n\t"
        "mad.lo.cc.u32 t4, %9, %19, t4;\n\t"
        "madc.hi.cc.u32 t5, %9, %19, t5;\n\t"
        "madc.lo.cc.u32 t6, %11, %19, t6;\n\t"
        "madc.hi.cc.u32 t7, %11, %19, t7;\n\t"
        "madc.lo.cc.u32 t8, %13, %19, t8;\n\t"
        "madc.hi.cc.u32 t9, %13, %19, t9;\n\t"
        "madc.lo.cc.u32 t10, %15, %19, t10;\n\t"
        "madc.hi.u32 t11, %15, %19, 0x0;\n\t"
        "mad.lo.cc.u32 t19, %8, %19, t19;\n\t"
        "madc.hi.cc.u32 t20, %8, %19, t20;\n\t"
        "madc.lo.cc.u32 t21, %10, %19, t21;\n\t"
        "madc.hi.cc.u32 t22, %10, %19, t22;\n\t"
        "madc.lo.cc.u32 t23, %12, %19, t23;\n\t"
        "madc.hi.cc.u32 t24, %12, %19, t24;\n\t"
        "madc.lo.cc.u32 t25, %14, %19, t25;\n\t"
        "madc.hi.cc.u32 t26, %14, %19, t26;\n\t"
        "addc.u32 t27, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t4, %8, %20, t4;\n\t"
        "madc.hi.cc.u32 t5, %8, %20, t5;\n\t"
        "madc.lo.cc.u32 t6, %10, %20, t6;\n\t"
        "madc.hi.cc.u32 t7, %10, %20, t7;\n\t"
        "madc.lo.cc.u32 t8, %12, %20, t8;\n\t"
        "madc.hi.cc.u32 t9, %12, %20, t9;\n\t"
        "madc.lo.cc.u32 t10, %14, %20, t10;\n\t"
        "madc.hi.cc.u32 t11, %14, %20, t11;\n\t"
        "addc.u32 t12, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t21, %9, %20, t21;\n\t"
        "madc.hi.cc.u32 t22, %9, %20, t22;\n\t"
        "madc.lo.cc.u32 t23, %11, %20, t23;\n\t"
        "madc.hi.cc.u32 t24, %11, %20, t24;\n\t"
        "madc.lo.cc.u32 t25, %13, %20, t25;\n\t"
        "madc.hi.cc.u32 t26, %13, %20, t26;\n\t"
        "madc.lo.cc.u32 t27, %15, %20, t27;\n\t"
        "madc.hi.u32 t28, %15, %20, 0x0;\n\t"
        "mad.lo.cc.u32 t6, %9, %21, t6;\n\t"
        "madc.hi.cc.u32 t7, %9, %21, t7;\n\t"
        "madc.lo.cc.u32 t8, %11, %21, t8;\n\t"
        "madc.hi.cc.u32 t9, %11, %21, t9;\n\t"
        "madc.lo.cc.u32 t10, %13, %21, t10;\n\t"
        "madc.hi.cc.u32 t11, %13, %21, t11;\n\t"
        "madc.lo.cc.u32 t12, %15, %21, t12;\n\t"
        "madc.hi.u32 t13, %15, %21, 0x0;\n\t"
        "mad.lo.cc.u32 t21, %8, %21, t21;\n\t"
        "madc.hi.cc.u32 t22, %8, %21, t22;\n\t"
        "madc.lo.cc.u32 t23, %10, %21, t23;\n\t"
        "madc.hi.cc.u32 t24, %10, %21, t24;\n\t"
        "madc.lo.cc.u32 t25, %12, %21, t25;\n\t"
        "madc.hi.cc.u32 t26, %12, %21, t26;\n\t"
        "madc.lo.cc.u32 t27, %14, %21, t27;\n\t"
        "madc.hi.cc.u32 t28, %14, %21, t28;\n\t"
        "addc.u32 t29, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t6, %8, %22, t6;\n\t"
        "madc.hi.cc.u32 t7, %8, %22, t7;\n\t"
        "madc.lo.cc.u32 t8, %10, %22, t8;\n\t"
        "madc.hi.cc.u32 t9, %10, %22, t9;\n\t"
        "madc.lo.cc.u32 t10, %12, %22, t10;\n\t"
        "madc.hi.cc.u32 t11, %12, %22, t11;\n\t"
        "madc.lo.cc.u32 t12, %14, %22, t12;\n\t"
        "madc.hi.cc.u32 t13, %14, %22, t13;\n\t"
        "addc.u32 t14, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t23, %9, %22, t23;\n\t"
        "madc.hi.cc.u32 t24, %9, %22, t24;\n\t"
        "madc.lo.cc.u32 t25, %11, %22, t25;\n\t"
        "madc.hi.cc.u32 t26, %11, %22, t26;\n\t"
        "madc.lo.cc.u32 t27, %13, %22, t27;\n\t"
        "madc.hi.cc.u32 t28, %13, %22, t28;\n\t"
        "madc.lo.cc.u32 t29, %15, %22, t29;\n\t"
        "madc.hi.u32 t30, %15, %22, 0x0;\n\t"
        "mad.lo.cc.u32 t8, %9, %23, t8;\n\t"
        "madc.hi.cc.u32 t9, %9, %23, t9;\n\t"
        "madc.lo.cc.u32 t10, %11, %23, t10;\n\t"
        "madc.hi.cc.u32 t11, %11, %23, t11;\n\t"
        "madc.lo.cc.u32 t12, %13, %23, t12;\n\t"
        "madc.hi.cc.u32 t13, %13, %23, t13;\n\t"
        "madc.lo.cc.u32 t14, %15, %23, t14;\n\t"
        "madc.hi.u32 t15, %15, %23, 0x0;\n\t"
        "mad.lo.cc.u32 t23, %8, %23, t23;\n\t"
        "madc.hi.cc.u32 t24, %8, %23, t24;\n\t"
        "madc.lo.cc.u32 t25, %10, %23, t25;\n\t"
        "madc.hi.cc.u32 t26, %10, %23, t26;\n\t"
        "madc.lo.cc.u32 t27, %12, %23, t27;\n\t"
        "madc.hi.cc.u32 t28, %12, %23, t28;\n\t"
        "madc.lo.cc.u32 t29, %14, %23, t29;\n\t"
        "madc.hi.cc.u32 t30, %14, %23, t30;\n\t"
        "addc.cc.u32 t31, 0x0, 0x0;\n\t"
        "addc.cc.u32 t32, t1, t17;\n\t"
        "addc.cc.u32 t33, t2, t18;\n\t"
        "addc.cc.u32 t34, t3, t19;\n\t"
        "addc.cc.u32 t35, t4, t20;\n\t"
        "addc.cc.u32 t36, t5, t21;\n\t"
        "addc.cc.u32 t37, t6, t22;\n\t"
        "addc.cc.u32 t38, t7, t23;\n\t"
        "addc.cc.u32 t39, t8, t24;\n\t"
        "addc.cc.u32 t40, t9, t25;\n\t"
        "addc.cc.u32 t41, t10, t26;\n\t"
        "addc.cc.u32 t42, t11, t27;\n\t"
        "addc.cc.u32 t43, t12, t28;\n\t"
        "addc.cc.u32 t44, t13, t29;\n\t"
        "addc.cc.u32 t45, t14, t30;\n\t"
        "addc.u32 t46, t15, t31;\n\t"
        "add.cc.u32 t47, t34, t0;\n\t"
        "addc.cc.u32 t48, t35, t32;\n\t"
        "addc.cc.u32 t49, t36, t33;\n\t"
        "addc.cc.u32 t50, t37, t34;\n\t"
        "addc.u32 t51, t38, t35;\n\t"
        "shl.b32 t52, t0, 1;\n\t"
        "shf.l.wrap.b32 t53, t0, t32, 1;\n\t"
        "add.cc.u32 t54, t50, t52;\n\t"
        "addc.u32 t55, t51, t53;\n\t"
        "sub.u32 t56, t55, t0;\n\t"
        "add.cc.u32 t57, t47, t0;\n\t"
        "addc.cc.u32 t58, t48, t32;\n\t"
        "addc.cc.u32 t59, t49, t33;\n\t"
        "addc.cc.u32 t60, t54, t47;\n\t"
        "addc.cc.u32 t61, t56, t48;\n\t"
        "addc.cc.u32 t62, 0x0, t49;\n\t"
        "addc.cc.u32 t63, 0x0, t54;\n\t"
        "addc.cc.u32 t64, 0x0, t56;\n\t"
        "addc.cc.u32 t65, 0x0, 0x0;\n\t"
        "addc.u32 t66, 0x0, 0x0;\n\t"
        "add.cc.u32 t67, t59, t0;\n\t"
        "addc.cc.u32 t68, t60, t32;\n\t"
        "addc.cc.u32 t69, t61, t33;\n\t"
        "addc.cc.u32 t70, t62, t47;\n\t"
        "addc.cc.u32 t71, t63, t48;\n\t"
        "addc.cc.u32 t72, t64, t49;\n\t"
        "addc.cc.u32 t73, t65, t54;\n\t"
        "addc.u32 t74, t66, t56;\n\t"
        "sub.cc.u32 t75, t58, t0;\n\t"
        "subc.cc.u32 t76, t67, t32;\n\t"
        "subc.cc.u32 t77, t68, t33;\n\t"
        "subc.cc.u32 t78, t69, t47;\n\t"
        "subc.cc.u32 t79, t70, t48;\n\t"
        "subc.cc.u32 t80, t71, t49;\n\t"
        "subc.cc.u32 t81, t72, t54;\n\t"
        "subc.cc.u32 t82, t73, t56;\n\t"
        "subc.u32 t83, t74, 0x0;\n\t"
        "add.cc.u32 t84, t34, t0;\n\t"
        "addc.cc.u32 t85, t35, t32;\n\t"
        "addc.cc.u32 t86, t36, t33;\n\t"
        "addc.cc.u32 t87, t37, t57;\n\t"
        "addc.cc.u32 t88, t38, t75;\n\t"
        "addc.cc.u32 t89, t39, t76;\n\t"
        "addc.cc.u32 t90, t40, t77;\n\t"
        "addc.cc.u32 t91, t41, t78;\n\t"
        "addc.cc.u32 t92, t42, t79;\n\t"
        "addc.cc.u32 t93, t43, t80;\n\t"
        "addc.cc.u32 t94, t44, t81;\n\t"
        "addc.cc.u32 t95, t45, t82;\n\t"
        "addc.cc.u32 t96, t46, t83;\n\t"
        "addc.u32 t97, 0x0, 0x0;\n\t"
        "sub.u32 t98, 0x0, t97;\n\t"
        "and.b32 t99, t98, 0xfffffffe;\n\t"
        "add.cc.u32 t100, t89, t97;\n\t"
        "addc.cc.u32 t101, t90, 0x0;\n\t"
        "addc.cc.u32 t102, t91, 0x0;\n\t"
        "addc.cc.u32 t103, t92, t98;\n\t"
        "addc.cc.u32 t104, t93, t98;\n\t"
        "addc.cc.u32 t105, t94, t98;\n\t"
        "addc.cc.u32 t106, t95, t99;\n\t"
        "addc.u32 t107, t96, 0x0;\n\t"
        "mov.u32 %0, t100;\n\t"
        "mov.u32 %1, t101;\n\t"
        "mov.u32 %2, t102;\n\t"
        "mov.u32 %3, t103;\n\t"
        "mov.u32 %4, t104;\n\t"
        "mov.u32 %5, t105;\n\t"
        "mov.u32 %6, t106;\n\t"
        "mov.u32 %7, t107;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t b_0_i = b[0];
    const uint32_t b_1_i = b[1];
    const uint32_t b_2_i = b[2];
    const uint32_t b_3_i = b[3];
    const uint32_t b_4_i = b[4];
    const uint32_t b_5_i = b[5];
    const uint32_t b_6_i = b[6];
    const uint32_t b_7_i = b[7];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t0 = (uint32_t)((uint32_t)(a_0_i * b_0_i));
    t1 = (uint32_t)(((uint64_t)a_0_i * b_0_i) >> 32);
    t2 = (uint32_t)((uint32_t)(a_2_i * b_0_i));
    t3 = (uint32_t)(((uint64_t)a_2_i * b_0_i) >> 32);
    t4 = (uint32_t)((uint32_t)(a_4_i * b_0_i));
    t5 = (uint32_t)(((uint64_t)a_4_i * b_0_i) >> 32);
    t6 = (uint32_t)((uint32_t)(a_6_i * b_0_i));
    t7 = (uint32_t)(((uint64_t)a_6_i * b_0_i) >> 32);
    t17 = (uint32_t)((uint32_t)(a_1_i * b_0_i));
    t18 = (uint32_t)(((uint64_t)a_1_i * b_0_i) >> 32);
    t19 = (uint32_t)((uint32_t)(a_3_i * b_0_i));
    t20 = (uint32_t)(((uint64_t)a_3_i * b_0_i) >> 32);
    t21 = (uint32_t)((uint32_t)(a_5_i * b_0_i));
    t22 = (uint32_t)(((uint64_t)a_5_i * b_0_i) >> 32);
    t23 = (uint32_t)((uint32_t)(a_7_i * b_0_i));
    t24 = (uint32_t)(((uint64_t)a_7_i * b_0_i) >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * b_1_i) + t2; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_1_i) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_1_i) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_1_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_1_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_1_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_1_i) + 0x0u + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_1_i) >> 32) + 0x0u + cf_; t9 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_1_i) + t17; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_1_i) >> 32) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_1_i) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_1_i) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_1_i) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_1_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_1_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_1_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t25 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_2_i) + t2; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_2_i) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_2_i) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_2_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_2_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_2_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_2_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_2_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_2_i) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_2_i) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_2_i) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_2_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_2_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_2_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_2_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_2_i) >> 32) + 0x0u + cf_; t26 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_3_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_3_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_3_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_3_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_3_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_3_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_3_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_3_i) >> 32) + 0x0u + cf_; t11 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_3_i) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_3_i) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_3_i) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_3_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_3_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_3_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_3_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_3_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_4_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_4_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_4_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_4_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_4_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_4_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_4_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_4_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_4_i) + t21; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_4_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_4_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_4_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_4_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_4_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_4_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_4_i) >> 32) + 0x0u + cf_; t28 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_5_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_5_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_5_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_5_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_5_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_5_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_5_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_5_i) >> 32) + 0x0u + cf_; t13 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_5_i) + t21; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_5_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_5_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_5_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_5_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_5_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_5_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_5_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_6_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_6_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_6_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_6_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_6_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_6_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_6_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_6_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t14 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_6_i) + t23; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_6_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_6_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_6_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_6_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_6_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_6_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_6_i) >> 32) + 0x0u + cf_; t30 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_7_i) + t8; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_7_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_7_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_7_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_7_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_7_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_7_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_7_i) >> 32) + 0x0u + cf_; t15 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_7_i) + t23; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_7_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_7_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_7_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_7_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_7_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_7_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_7_i) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t1 + t17 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t18 + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t19 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t20 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t21 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t22 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t23 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t8 + t24 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t25 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t26 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t27 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t28 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t29 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + t30 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + t31 + cf_; t46 = (uint32_t)w_;
    w_ = (uint64_t)t34 + t0; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t35 + t32 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t36 + t33 + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t37 + t34 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t38 + t35 + cf_; t51 = (uint32_t)w_;
    t52 = (uint32_t)(t0 << 1);
    t53 = (uint32_t)(((((uint64_t)t32 << 32) | t0) << 1) >> 32);
    w_ = (uint64_t)t50 + t52; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t51 + t53 + cf_; t55 = (uint32_t)w_;
    w_ = (uint64_t)t55 - t0; t56 = (uint32_t)w_;
    w_ = (uint64_t)t47 + t0; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t48 + t32 + cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t49 + t33 + cf_; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t54 + t47 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t56 + t48 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t49 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t54 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t56 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t66 = (uint32_t)w_;
    w_ = (uint64_t)t59 + t0; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t60 + t32 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t61 + t33 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t62 + t47 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + t48 + cf_; t71 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t64 + t49 + cf_; t72 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t65 + t54 + cf_; t73 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t66 + t56 + cf_; t74 = (uint32_t)w_;
    w_ = (uint64_t)t58 - t0; t75 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t67 - t32 - cf_; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t68 - t33 - cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t69 - t47 - cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t70 - t48 - cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t71 - t49 - cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t72 - t54 - cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t73 - t56 - cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t74 - 0x0u - cf_; t83 = (uint32_t)w_;
    w_ = (uint64_t)t34 + t0; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t35 + t32 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t36 + t33 + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t37 + t57 + cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t38 + t75 + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t39 + t76 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + t77 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t41 + t78 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t42 + t79 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t43 + t80 + cf_; t93 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t44 + t81 + cf_; t94 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t45 + t82 + cf_; t95 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t46 + t83 + cf_; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t97 = (uint32_t)w_;
    w_ = (uint64_t)0x0u - t97; t98 = (uint32_t)w_;
    t99 = (uint32_t)(t98 & 0xfffffffeu);
    w_ = (uint64_t)t89 + t97; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t90 + 0x0u + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t91 + 0x0u + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t92 + t98 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t93 + t98 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t94 + t98 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t95 + t99 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t96 + 0x0u + cf_; t107 = (uint32_t)w_;
    r[0] = t100;
    r[1] = t101;
    r[2] = t102;
    r[3] = t103;
    r[4] = t104;
    r[5] = t105;
    r[6] = t106;
    r[7] = t107;
#endif
  }

  static MAB_DEV void sqr_w(uint32_t (&r)[8], const uint32_t (&a)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<138>;\n\t"
        "mul.lo.u32 t2, %8, %10;\n\t"
        "mul.hi.u32 t3, %8, %10;\n\t"
        "mul.lo.u32 t4, %8, %12;\n\t"
        "mul.hi.u32 t5, %8, %12;\n\t"
        "mul.lo.u32 t6, %8, %14;\n\t"
        "mul.hi.u32 t7, %8, %14;\n\t"
        "mul.lo.u32 t17, %8, %9;\n\t"
        "mul.hi.u32 t18, %8, %9;\n\t"
        "mul.lo.u32 t19, %8, %11;\n\t"
        "mul.hi.u32 t20, %8, %11;\n\t"
        "mul.lo.u32 t21, %8, %13;\n\t"
        "mul.hi.u32 t22, %8, %13;\n\t"
        "mul.lo.u32 t23, %8, %15;\n\t"
        "mul.hi.u32 t24, %8, %15;\n\t"
        "mad.lo.cc.u32 t4, %9, %11, t4;\n\t"
        "madc.hi.cc.u32 t5, %9, %11, t5;\n\t"
        "madc.lo.cc.u32 t6, %9, %13, t6;\n\t"
        "madc.hi.cc.u32 t7, %9, %13, t7;\n\t"
        "madc.lo.cc.u32 t8, %9, %15, 0x0;\n\t"
        "madc.hi.u32 t9, %9, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t19, %9, %10, t19;\n\t"
        "madc.hi.cc.u32 t20, %9, %10, t20;\n\t"
        "madc.lo.cc.u32 t21, %9, %12, t21;\n\t"
        "madc.hi.cc.u32 t22, %9, %12, t22;\n\t"
        "madc.lo.cc.u32 t23, %9, %14, t23;\n\t"
        "madc.hi.cc.u32 t24, %9, %14, t24;\n\t"
        "madc.lo.u32 t25, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t6, %10, %12, t6;\n\t"
        "madc.hi.cc.u32 t7, %10, %12, t7;\n\t"
        "madc.lo.cc.u32 t8, %10, %14, t8;\n\t"
        "madc.hi.cc.u32 t9, %10, %14, t9;\n\t"
        "madc.lo.u32 t10, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t21, %10, %11, t21;\n\t"
        "madc.hi.cc.u32 t22, %10, %11, t22;\n\t"
        "madc.lo.cc.u32 t23, %10, %13, t23;\n\t"
        "madc.hi.cc.u32 t24, %10, %13, t24;\n\t"
        "madc.lo.cc.u32 t25, %10, %15, t25;\n\t"
        "madc.hi.u32 t26, %10, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t8, %11, %13, t8;\n\t"
        "madc.hi.cc.u32 t9, %11, %13, t9;\n\t"
        "madc.lo.cc.u32 t10, %11, %15, t10;\n\t"
        "madc.hi.u32 t11, %11, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t23, %11, %12, t23;\n\t"
        "madc.hi.cc.u32 t24, %11, %12, t24;\n\t"
        "madc.lo.cc.u32 t25, %11, %14, t25;\n\t"
        "madc.hi.cc.u32 t26, %11, %14, t26;\n\t"
        "madc.lo.u32 t27, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t10, %12, %14, t10;\n\t"
        "madc.hi.cc.u32 t11, %12, %14, t11;\n\t"
        "madc.lo.u32 t12, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t25, %12, %13, t25;\n\t"
        "madc.hi.cc.u32 t26, %12, %13, t26;\n\t"
        "madc.lo.cc.u32 t27, %12, %15, t27;\n\t"
        "madc.hi.u32 t28, %12, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t12, %13, %15, t12;\n\t"
        "madc.hi.u32 t13, %13, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t27, %13, %14, t27;\n\t"
        "madc.hi.cc.u32 t28, %13, %14, t28;\n\t"
        "madc.lo.u32 t29, 0, 0, 0;\n\t"
        "mad.lo.cc.u32 t29, %14, %15, t29;\n\t"
        "madc.hi.cc.u32 t30, %14, %15, 0x0;\n\t"
        "addc.cc.u32 t32, t2, t18;\n\t"
        "addc.cc.u32 t33, t3, t19;\n\t"
        "addc.cc.u32 t34, t4, t20;\n\t"
        "addc.cc.u32 t35, t5, t21;\n\t"
        "addc.cc.u32 t36, t6, t22;\n\t"
        "addc.cc.u32 t37, t7, t23;\n\t"
        "addc.cc.u32 t38, t8, t24;\n\t"
        "addc.cc.u32 t39, t9, t25;\n\t"
        "addc.cc.u32 t40, t10, t26;\n\t"
        "addc.cc.u32 t41, t11, t27;\n\t"
        "addc.cc.u32 t42, t12, t28;\n\t"
        "addc.cc.u32 t43, t13, t29;\n\t"
        "addc.cc.u32 t44, 0x0, t30;\n\t"
        "madc.lo.u32 t45, 0, 0, 0;\n\t"
        "shl.b32 t46, t17, 1;\n\t"
        "shf.l.wrap.b32 t47, t17, t32, 1;\n\t"
        "shf.l.wrap.b32 t48, t32, t33, 1;\n\t"
        "shf.l.wrap.b32 t49, t33, t34, 1;\n\t"
        "shf.l.wrap.b32 t50, t34, t35, 1;\n\t"
        "shf.l.wrap.b32 t51, t35, t36, 1;\n\t"
        "shf.l.wrap.b32 t52, t36, t37, 1;\n\t"
        "shf.l.wrap.b32 t53, t37, t38, 1;\n\t"
        "shf.l.wrap.b32 t54, t38, t39, 1;\n\t"
        "shf.l.wrap.b32 t55, t39, t40, 1;\n\t"
        "shf.l.wrap.b32 t56, t40, t41, 1;\n\t"
        "shf.l.wrap.b32 t57, t41, t42, 1;\n\t"
        "shf.l.wrap.b32 t58, t42, t43, 1;\n\t"
        "shf.l.wrap.b32 t59, t43, t44, 1;\n\t"
        "shf.l.wrap.b32 t60, t44, t45, 1;\n\t"
        "mad.lo.cc.u32 t61, %8, %8, 0x0;\n\t"
        "madc.hi.cc.u32 t62, %8, %8, t46;\n\t"
        "madc.lo.cc.u32 t63, %9, %9, t47;\n\t"
        "madc.hi.cc.u32 t64, %9, %9, t48;\n\t"
        "madc.lo.cc.u32 t65, %10, %10, t49;\n\t"
        "madc.hi.cc.u32 t66, %10, %10, t50;\n\t"
        "madc.lo.cc.u32 t67, %11, %11, t51;\n\t"
        "madc.hi.cc.u32 t68, %11, %11, t52;\n\t"
        "madc.lo.cc.u32 t69, %12, %12, t53;\n\t"
        "madc.hi.cc.u32 t70, %12, %12, t54;\n\t"
        "madc.lo.cc.u32 t71, %13, %13, t55;\n\t"
        "madc.hi.cc.u32 t72, %13, %13, t56;\n\t"
        "madc.lo.cc.u32 t73, %14, %14, t57;\n\t"
        "madc.hi.cc.u32 t74, %14, %14, t58;\n\t"
        "madc.lo.cc.u32 t75, %15, %15, t59;\n\t"
        "madc.hi.u32 t76, %15, %15, t60;\n\t"
        "add.cc.u32 t77, t64, t61;\n\t"
        "addc.cc.u32 t78, t65, t62;\n\t"
        "addc.cc.u32 t79, t66, t63;\n\t"
        "addc.cc.u32 t80, t67, t64;\n\t"
        "addc.u32 t81, t68, t65;\n\t"
        "shl.b32 t82, t61, 1;\n\t"
        "shf.l.wrap.b32 t83, t61, t62, 1;\n\t"
        "add.cc.u32 t84, t80, t82;\n\t"
        "addc.u32 t85, t81, t83;\n\t"
        "sub.u32 t86, t85, t61;\n\t"
        "add.cc.u32 t87, t77, t61;\n\t"
        "addc.cc.u32 t88, t78, t62;\n\t"
        "addc.cc.u32 t89, t79, t63;\n\t"
        "addc.cc.u32 t90, t84, t77;\n\t"
        "addc.cc.u32 t91, t86, t78;\n\t"
        "addc.cc.u32 t92, 0x0, t79;\n\t"
        "addc.cc.u32 t93, 0x0, t84;\n\t"
        "addc.cc.u32 t94, 0x0, t86;\n\t"
        "addc.cc.u32 t95, 0x0, 0x0;\n\t"
        "madc.lo.u32 t96, 0, 0, 0;\n\t"
        "add.cc.u32 t97, t89, t61;\n\t"
        "addc.cc.u32 t98, t90, t62;\n\t"
        "addc.cc.u32 t99, t91, t63;\n\t"
        "addc.cc.u32 t100, t92, t77;\n\t"
        "addc.cc.u32 t101, t93, t78;\n\t"
        "addc.cc.u32 t102, t94, t79;\n\t"
        "addc.cc.u32 t103, t95, t84;\n\t"
        "addc.u32 t104, t96, t86;\n\t"
        "sub.cc.u32 t105, t88, t61;\n\t"
        "subc.cc.u32 t106, t97, t62;\n\t"
        "subc.cc.u32 t107, t98, t63;\n\t"
        "subc.cc.u32 t108, t99, t77;\n\t"
        "subc.cc.u32 t109, t100, t78;\n\t"
        "subc.cc.u32 t110, t101, t79;\n\t"
        "subc.cc.u32 t111, t102, t84;\n\t"
        "subc.cc.u32 t112, t103, t86;\n\t"
        "subc.u32 t113, t104, 0x0;\n\t"
        "add.cc.u32 t114, t64, t61;\n\t"
        "addc.cc.u32 t115, t65, t62;\n\t"
        "addc.cc.u32 t116, t66, t63;\n\t"
        "addc.cc.u32 t117, t67, t87;\n\t"
        "addc.cc.u32 t118, t68, t105;\n\t"
        "addc.cc.u32 t119, t69, t106;\n\t"
        "addc.cc.u32 t120, t70, t107;\n\t"
        "addc.cc.u32 t121, t71, t108;\n\t"
        "addc.cc.u32 t122, t72, t109;\n\t"
        "addc.cc.u32 t123, t73, t110;\n\t"
        "addc.cc.u32 t124, t74, t111;\n\t"
        "addc.cc.u32 t125, t75, t112;\n\t"
        "addc.cc.u32 t126, t76, t113;\n\t"
        "madc.lo.u32 t127, 0, 0, 0;\n\t"
        "sub.u32 t128, 0x0, t127;\n\t"
        "and.b32 t129, t128, 0xfffffffe;\n\t"
        "add.cc.u32 t130, t119, t127;\n\t"
        "addc.cc.u32 t131, t120, 0x0;\n\t"
        "addc.cc.u32 t132, t121, 0x0;\n\t"
        "addc.cc.u32 t133, t122, t128;\n\t"
        "addc.cc.u32 t134, t123, t128;\n\t"
        "addc.cc.u32 t135, t124, t128;\n\t"
        "addc.cc.u32 t136, t125, t129;\n\t"
        "addc.u32 t137, t126, 0x0;\n\t"
        "mov.u32 %0, t130;\n\t"
        "mov.u32 %1, t131;\n\t"
        "mov.u32 %2, t132;\n\t"
        "mov.u32 %3, t133;\n\t"
        "mov.u32 %4, t134;\n\t"
        "mov.u32 %5, t135;\n\t"
        "mov.u32 %6, t136;\n\t"
        "mov.u32 %7, t137;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127, t128, t129, t130, t131, t132, t133, t134, t135, t136, t137;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t2 = (uint32_t)((uint32_t)(a_0_i * a_2_i));
    t3 = (uint32_t)(((uint64_t)a_0_i * a_2_i) >> 32);
    t4 = (uint32_t)((uint32_t)(a_0_i * a_4_i));
    t5 = (uint32_t)(((uint64_t)a_0_i * a_4_i) >> 32);
    t6 = (uint32_t)((uint32_t)(a_0_i * a_6_i));
    t7 = (uint32_t)(((uint64_t)a_0_i * a_6_i) >> 32);
    t17 = (uint32_t)((uint32_t)(a_0_i * a_1_i));
    t18 = (uint32_t)(((uint64_t)a_0_i * a_1_i) >> 32);
    t19 = (uint32_t)((uint32_t)(a_0_i * a_3_i));
    t20 = (uint32_t)(((uint64_t)a_0_i * a_3_i) >> 32);
    t21 = (uint32_t)((uint32_t)(a_0_i * a_5_i));
    t22 = (uint32_t)(((uint64_t)a_0_i * a_5_i) >> 32);
    t23 = (uint32_t)((uint32_t)(a_0_i * a_7_i));
    t24 = (uint32_t)(((uint64_t)a_0_i * a_7_i) >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_3_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_3_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_5_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_5_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_7_i) + 0x0u + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_7_i) >> 32) + 0x0u + cf_; t9 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * a_2_i) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_2_i) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_4_i) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_4_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_6_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_6_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t25 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_2_i * a_4_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_4_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_6_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_6_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_2_i * a_3_i) + t21; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_3_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_5_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_5_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_7_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_7_i) >> 32) + 0x0u + cf_; t26 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_3_i * a_5_i) + t8; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_5_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_7_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_7_i) >> 32) + 0x0u + cf_; t11 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_3_i * a_4_i) + t23; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_4_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_6_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_6_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_4_i * a_6_i) + t10; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_6_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_4_i * a_5_i) + t25; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_5_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_7_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_7_i) >> 32) + 0x0u + cf_; t28 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_5_i * a_7_i) + t12; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_7_i) >> 32) + 0x0u + cf_; t13 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_5_i * a_6_i) + t27; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_6_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_6_i * a_7_i) + t29; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_7_i) >> 32) + 0x0u + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t18 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t19 + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t20 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t21 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t22 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t23 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t8 + t24 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t25 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t26 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t27 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t28 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t29 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t30 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t45 = (uint32_t)w_;
    t46 = (uint32_t)(t17 << 1);
    t47 = (uint32_t)(((((uint64_t)t32 << 32) | t17) << 1) >> 32);
    t48 = (uint32_t)(((((uint64_t)t33 << 32) | t32) << 1) >> 32);
    t49 = (uint32_t)(((((uint64_t)t34 << 32) | t33) << 1) >> 32);
    t50 = (uint32_t)(((((uint64_t)t35 << 32) | t34) << 1) >> 32);
    t51 = (uint32_t)(((((uint64_t)t36 << 32) | t35) << 1) >> 32);
    t52 = (uint32_t)(((((uint64_t)t37 << 32) | t36) << 1) >> 32);
    t53 = (uint32_t)(((((uint64_t)t38 << 32) | t37) << 1) >> 32);
    t54 = (uint32_t)(((((uint64_t)t39 << 32) | t38) << 1) >> 32);
    t55 = (uint32_t)(((((uint64_t)t40 << 32) | t39) << 1) >> 32);
    t56 = (uint32_t)(((((uint64_t)t41 << 32) | t40) << 1) >> 32);
    t57 = (uint32_t)(((((uint64_t)t42 << 32) | t41) << 1) >> 32);
    t58 = (uint32_t)(((((uint64_t)t43 << 32) | t42) << 1) >> 32);
    t59 = (uint32_t)(((((uint64_t)t44 << 32) | t43) << 1) >> 32);
    t60 = (uint32_t)(((((uint64_t)t45 << 32) | t44) << 1) >> 32);
    w_ = (uint64_t)(uint32_t)(a_0_i * a_0_i) + 0x0u; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * a_0_i) >> 32) + t46 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_1_i) + t47 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_1_i) >> 32) + t48 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_2_i) + t49 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_2_i) >> 32) + t50 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_3_i) + t51 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_3_i) >> 32) + t52 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_4_i) + t53 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_4_i) >> 32) + t54 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_5_i) + t55 + cf_; t71 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_5_i) >> 32) + t56 + cf_; t72 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_6_i) + t57 + cf_; t73 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_6_i) >> 32) + t58 + cf_; t74 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_7_i) + t59 + cf_; t75 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_7_i) >> 32) + t60 + cf_; t76 = (uint32_t)w_;
    w_ = (uint64_t)t64 + t61; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t65 + t62 + cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t66 + t63 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t67 + t64 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t68 + t65 + cf_; t81 = (uint32_t)w_;
    t82 = (uint32_t)(t61 << 1);
    t83 = (uint32_t)(((((uint64_t)t62 << 32) | t61) << 1) >> 32);
    w_ = (uint64_t)t80 + t82; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t81 + t83 + cf_; t85 = (uint32_t)w_;
    w_ = (uint64_t)t85 - t61; t86 = (uint32_t)w_;
    w_ = (uint64_t)t77 + t61; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t78 + t62 + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t79 + t63 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t84 + t77 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t86 + t78 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t79 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t84 + cf_; t93 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t86 + cf_; t94 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t95 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t96 = (uint32_t)w_;
    w_ = (uint64_t)t89 + t61; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t90 + t62 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t91 + t63 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t92 + t77 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t93 + t78 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t94 + t79 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t95 + t84 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t96 + t86 + cf_; t104 = (uint32_t)w_;
    w_ = (uint64_t)t88 - t61; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t97 - t62 - cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t98 - t63 - cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t99 - t77 - cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t100 - t78 - cf_; t109 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t101 - t79 - cf_; t110 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t102 - t84 - cf_; t111 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t103 - t86 - cf_; t112 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t104 - 0x0u - cf_; t113 = (uint32_t)w_;
    w_ = (uint64_t)t64 + t61; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t65 + t62 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t66 + t63 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t67 + t87 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t68 + t105 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t69 + t106 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t70 + t107 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t71 + t108 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t72 + t109 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t73 + t110 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t74 + t111 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t75 + t112 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t76 + t113 + cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t127 = (uint32_t)w_;
    w_ = (uint64_t)0x0u - t127; t128 = (uint32_t)w_;
    t129 = (uint32_t)(t128 & 0xfffffffeu);
    w_ = (uint64_t)t119 + t127; t130 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t120 + 0x0u + cf_; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t121 + 0x0u + cf_; t132 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t122 + t128 + cf_; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t123 + t128 + cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t124 + t128 + cf_; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t125 + t129 + cf_; t136 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t126 + 0x0u + cf_; t137 = (uint32_t)w_;
    r[0] = t130;
    r[1] = t131;
    r[2] = t132;
    r[3] = t133;
    r[4] = t134;
    r[5] = t135;
    r[6] = t136;
    r[7] = t137;
#endif
  }

  // n = -b (pseudo.py:329-348)
  static MAB_DEV void neg(uint32_t (&r)[8], const uint32_t (&b)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<18>;\n\t"
        "sub.cc.u32 t0, 0x0, %8;\n\t"
        "subc.cc.u32 t1, 0x0, %9;\n\t"
        "subc.cc.u32 t2, 0x0, %10;\n\t"
        "subc.cc.u32 t3, 0x0, %11;\n\t"
        "subc.cc.u32 t4, 0x0, %12;\n\t"
        "subc.cc.u32 t5, 0x0, %13;\n\t"
        "subc.cc.u32 t6, 0x0, %14;\n\t"
        "subc.cc.u32 t7, 0x0, %15;\n\t"
        "subc.u32 t8, 0x0, 0x0;\n\t"
        "and.b32 t9, t8, 0x1;\n\t"
        "add.cc.u32 t10, t0, t8;\n\t"
        "addc.cc.u32 t11, t1, t8;\n\t"
        "addc.cc.u32 t12, t2, t8;\n\t"
        "addc.cc.u32 t13, t3, 0x0;\n\t"
        "addc.cc.u32 t14, t4, 0x0;\n\t"
        "addc.cc.u32 t15, t5, 0x0;\n\t"
        "addc.cc.u32 t16, t6, t9;\n\t"
        "addc.u32 t17, t7, t8;\n\t"
        "mov.u32 %0, t10;\n\t"
        "mov.u32 %1, t11;\n\t"
        "mov.u32 %2, t12;\n\t"
        "mov.u32 %3, t13;\n\t"
        "mov.u32 %4, t14;\n\t"
        "mov.u32 %5, t15;\n\t"
        "mov.u32 %6, t16;\n\t"
        "mov.u32 %7, t17;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    const uint32_t b_0_i = b[0];
    const uint32_t b_1_i = b[1];
    const uint32_t b_2_i = b[2];
    const uint32_t b_3_i = b[3];
    const uint32_t b_4_i = b[4];
    const uint32_t b_5_i = b[5];
    const uint32_t b_6_i = b[6];
    const uint32_t b_7_i = b[7];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    w_ = (uint64_t)0x0u - b_0_i; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_1_i - cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_2_i - cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_3_i - cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_4_i - cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_5_i - cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_6_i - cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_7_i - cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t8 = (uint32_t)w_;
    t9 = (uint32_t)(t8 & 0x1u);
    w_ = (uint64_t)t0 + t8; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t1 + t8 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t8 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + 0x0u + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + 0x0u + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + 0x0u + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t9 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t8 + cf_; t17 = (uint32_t)w_;
    r[0] = t10;
    r[1] = t11;
    r[2] = t12;
    r[3] = t13;
    r[4] = t14;
    r[5] = t15;
    r[6] = t16;
    r[7] = t17;
#endif
  }

  // canonical residue of a stored value; returns 1 iff it was already < p
  // (flatten/modfsb, pseudo.py:255-283)
  static MAB_DEV uint32_t canon(uint32_t (&r)[8], const uint32_t (&a)[8]) {
    uint32_t lt;
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<34>;\n\t"
        "sub.cc.u32 t1, %9, 0xffffffff;\n\t"
        "subc.cc.u32 t2, %10, 0xffffffff;\n\t"
        "subc.cc.u32 t3, %11, 0xffffffff;\n\t"
        "subc.cc.u32 t4, %12, 0x0;\n\t"
        "subc.cc.u32 t5, %13, 0x0;\n\t"
        "subc.cc.u32 t6, %14, 0x0;\n\t"
        "subc.cc.u32 t7, %15, 0x1;\n\t"
        "subc.cc.u32 t8, %16, 0xffffffff;\n\t"
        "subc.u32 t9, 0x0, 0x0;\n\t"
        "xor.b32 t10, t1, %9;\n\t"
        "and.b32 t11, t10, t9;\n\t"
        "xor.b32 t12, t11, t1;\n\t"
        "xor.b32 t13, t2, %10;\n\t"
        "and.b32 t14, t13, t9;\n\t"
        "xor.b32 t15, t14, t2;\n\t"
        "xor.b32 t16, t3, %11;\n\t"
        "and.b32 t17, t16, t9;\n\t"
        "xor.b32 t18, t17, t3;\n\t"
        "xor.b32 t19, t4, %12;\n\t"
        "and.b32 t20, t19, t9;\n\t"
        "xor.b32 t21, t20, t4;\n\t"
        "xor.b32 t22, t5, %13;\n\t"
        "and.b32 t23, t22, t9;\n\t"
        "xor.b32 t24, t23, t5;\n\t"
        "xor.b32 t25, t6, %14;\n\t"
        "and.b32 t26, t25, t9;\n\t"
        "xor.b32 t27, t26, t6;\n\t"
        "xor.b32 t28, t7, %15;\n\t"
        "and.b32 t29, t28, t9;\n\t"
        "xor.b32 t30, t29, t7;\n\t"
        "xor.b32 t31, t8, %16;\n\t"
        "and.b32 t32, t31, t9;\n\t"
        "xor.b32 t33, t32, t8;\n\t"
        "and.b32 t0, t9, 0x1;\n\t"
        "mov.u32 %0, t12;\n\t"
        "mov.u32 %1, t15;\n\t"
        "mov.u32 %2, t18;\n\t"
        "mov.u32 %3, t21;\n\t"
        "mov.u32 %4, t24;\n\t"
        "mov.u32 %5, t27;\n\t"
        "mov.u32 %6, t30;\n\t"
        "mov.u32 %7, t33;\n\t"
        "mov.u32 %8, t0;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(lt)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    w_ = (uint64_t)a_0_i - 0xffffffffu; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_1_i - 0xffffffffu - cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_2_i - 0xffffffffu - cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_3_i - 0x0u - cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_4_i - 0x0u - cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_5_i - 0x0u - cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_6_i - 0x1u - cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_7_i - 0xffffffffu - cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t9 = (uint32_t)w_;
    t10 = (uint32_t)(t1 ^ a_0_i);
    t11 = (uint32_t)(t10 & t9);
    t12 = (uint32_t)(t11 ^ t1);
    t13 = (uint32_t)(t2 ^ a_1_i);
    t14 = (uint32_t)(t13 & t9);
    t15 = (uint32_t)(t14 ^ t2);
    t16 = (uint32_t)(t3 ^ a_2_i);
    t17 = (uint32_t)(t16 & t9);
    t18 = (uint32_t)(t17 ^ t3);
    t19 = (uint32_t)(t4 ^ a_3_i);
    t20 = (uint32_t)(t19 & t9);
    t21 = (uint32_t)(t20 ^ t4);
    t22 = (uint32_t)(t5 ^ a_4_i);
    t23 = (uint32_t)(t22 & t9);
    t24 = (uint32_t)(t23 ^ t5);
    t25 = (uint32_t)(t6 ^ a_5_i);
    t26 = (uint32_t)(t25 & t9);
    t27 = (uint32_t)(t26 ^ t6);
    t28 = (uint32_t)(t7 ^ a_6_i);
    t29 = (uint32_t)(t28 & t9);
    t30 = (uint32_t)(t29 ^ t7);
    t31 = (uint32_t)(t8 ^ a_7_i);
    t32 = (uint32_t)(t31 & t9);
    t33 = (uint32_t)(t32 ^ t8);
    t0 = (uint32_t)(t9 & 0x1u);
    r[0] = t12;
    r[1] = t15;
    r[2] = t18;
    r[3] = t21;
    r[4] = t24;
    r[5] = t27;
    r[6] = t30;
    r[7] = t33;
    lt = t0;
#endif
    return lt;
  }

  static MAB_DEV void set_p(uint32_t (&r)[8]) { r[0] = 0xffffffffu; r[1] = 0xffffffffu; r[2] = 0xffffffffu; r[3] = 0x00000000u; r[4] = 0x00000000u; r[5] = 0x00000000u; r[6] = 0x00000001u; r[7] = 0xffffffffu; }
  static MAB_DEV void set_one(uint32_t (&r)[8]) { r[0] = 0x00000001u; r[1] = 0x00000000u; r[2] = 0x00000000u; r[3] = 0xffffffffu; r[4] = 0xffffffffu; r[5] = 0xffffffffu; r[6] = 0xfffffffeu; r[7] = 0x00000000u; }
  static MAB_DEV void set_roi(uint32_t (&r)[8]) { r[0] = 0xfffffffeu; r[1] = 0xffffffffu; r[2] = 0xffffffffu; r[3] = 0x00000001u; r[4] = 0x00000000u; r[5] = 0x00000000u; r[6] = 0x00000002u; r[7] = 0xfffffffeu; }
  static MAB_DEV void set_r2(uint32_t (&r)[8]) { r[0] = 0x00000003u; r[1] = 0x00000000u; r[2] = 0xffffffffu; r[3] = 0xfffffffbu; r[4] = 0xfffffffeu; r[5] = 0xffffffffu; r[6] = 0xfffffffdu; r[7] = 0x00000004u; }
  // short-Weierstrass curve y^2 = x^3 - 3x + b (curve.py:157-166): b in stored form
  static constexpr bool HAS_WEIERSTRASS = true;
  static MAB_DEV void set_b(uint32_t (&r)[8]) { r[0] = 0x29c4bddfu; r[1] = 0xd89cdf62u; r[2] = 0x78843090u; r[3] = 0xacf005cdu; r[4] = 0xf7212ed6u; r[5] = 0xe5a220abu; r[6] = 0x04874834u; r[7] = 0xdc30061du; }

  // nres: multiply by R^2 mod p (monty.py:1386-1399); redc: multiply by 1 (monty.py:1402-1416)
  static MAB_DEV void nres(uint32_t (&r)[8], const uint32_t (&a)[8]) { uint32_t c[L]; set_r2(c); mul(r, a, c); }
  static MAB_DEV void redc(uint32_t (&r)[8], const uint32_t (&a)[8]) { uint32_t c[L]; c[0] = 1;
    for (int i = 1; i < L; i++) c[i] = 0;
    mul(r, a, c); }

  // z = w^PE, straight-line addition chain (pseudo.py:758-785; our own chain finder)
  static MAB_DEV void pro(uint32_t (&z)[8], const uint32_t (&w)[8]) {
    uint32_t x[L];
    for (int i = 0; i < L; i++) x[i] = w[i];
    uint32_t t0[L];
    uint32_t t1[L];
    uint32_t t2[L];
    uint32_t t3[L];
    uint32_t t4[L];
    sqr_w(t0, x);
    mul_w(t0, t0, x);
    sqr_w(t1, t0);
    sqr_w(t1, t1);
    mul_w(t1, t1, t0);
    sqr_w(t2, t1);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr_w(t2, t2);
    mul_w(t2, t2, t1);
    sqr_w(t3, t2);
    MAB_NOUNROLL
    for (int i = 1; i < 8; i++) sqr_w(t3, t3);
    mul_w(t3, t3, t2);
    sqr_w(t4, t3);
    MAB_NOUNROLL
    for (int i = 1; i < 16; i++) sqr_w(t4, t4);
    mul_w(t4, t4, t3);
    sqr_w(z, t4);
    MAB_NOUNROLL
    for (int i = 1; i < 32; i++) sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 128; i++) sqr_w(z, z);
    mul_w(z, z, t4);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 32; i++) sqr_w(z, z);
    mul_w(z, z, t4);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 16; i++) sqr_w(z, z);
    mul_w(z, z, t3);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 8; i++) sqr_w(z, z);
    mul_w(z, z, t2);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr_w(z, z);
    mul_w(z, z, t1);
    sqr_w(z, z);
    sqr_w(z, z);
    mul_w(z, z, t0);
    if (WEAK) (void)canon(z, z);
  }
};
