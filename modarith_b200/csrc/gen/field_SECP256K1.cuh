// Automatically generated field arithmetic for sm_100a -- do not edit.
// Command line : python -m modarith_b200.gen.monty_sm100 SECP256K1
// modulus SECP256K1 = 0xfffffffffffffffffffffffffffffffffffffffffffffffffffffffefffffc2f
// plan PseudoMersenne33: 8 saturated 32-bit limbs; stored values < 2^256; R = 2^0
//   mul   :  73 IMAD.WIDE   2 IMAD  ~ 46 ALU-pipe ops
//   sqr   :  45 IMAD.WIDE   2 IMAD  ~ 65 ALU-pipe ops
//   mli   :   9 IMAD.WIDE   1 IMAD  ~ 21 ALU-pipe ops
//   mla   :   9 IMAD.WIDE   1 IMAD  ~ 22 ALU-pipe ops
//   add   :   0 IMAD.WIDE   2 IMAD  ~ 20 ALU-pipe ops
//   sub   :   0 IMAD.WIDE   0 IMAD  ~ 24 ALU-pipe ops
//   canon :   0 IMAD.WIDE   0 IMAD  ~ 34 ALU-pipe ops
//   modpro: 253 squarings + 18 multiplies (exponent (p-1-2^k)/2^(k+1), k=1)
#pragma once
#include "mab_common.cuh"

struct F_SECP256K1 {
  static constexpr int L = 8;
  static constexpr int NBITS = 256;
  static constexpr int NBYTES = 32;
  static constexpr int PM1D2 = 1;
  static constexpr bool MONTGOMERY = false;
  static constexpr int PRO_SQR = 253, PRO_MUL = 18;
  static constexpr int LADDER_MINBLOCKS = 3;   // resident 128-thread CTAs per SM for k_rfc7748
  static constexpr bool LADDER_STASH = false;   // scalar and x1 in shared memory (see rfc7748_sm100.cuh)
  static constexpr bool HAS_CURVE = false;
  static constexpr uint32_t A24 = 0;
  static constexpr int COF = 0;
  static constexpr uint32_t GENERATOR = 0;
  static const char* name() { return "SECP256K1"; }

  // c = a*b (pseudo.py:616-659 / monty.py:663-872)
  static MAB_DEV void mul(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<91>;\n\t"
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
        "addc.u32 t31, 0x0, 0x0;\n\t"
        "add.cc.u32 t32, t8, t24;\n\t"
        "addc.cc.u32 t33, t9, t25;\n\t"
        "addc.cc.u32 t34, t10, t26;\n\t"
        "addc.cc.u32 t35, t11, t27;\n\t"
        "addc.cc.u32 t36, t12, t28;\n\t"
        "addc.cc.u32 t37, t13, t29;\n\t"
        "addc.cc.u32 t38, t14, t30;\n\t"
        "addc.u32 t39, t15, t31;\n\t"
        "mad.lo.cc.u32 t40, t32, 0x3d1, t0;\n\t"
        "madc.hi.cc.u32 t41, t32, 0x3d1, t1;\n\t"
        "madc.lo.cc.u32 t42, t34, 0x3d1, t2;\n\t"
        "madc.hi.cc.u32 t43, t34, 0x3d1, t3;\n\t"
        "madc.lo.cc.u32 t44, t36, 0x3d1, t4;\n\t"
        "madc.hi.cc.u32 t45, t36, 0x3d1, t5;\n\t"
        "madc.lo.cc.u32 t46, t38, 0x3d1, t6;\n\t"
        "madc.hi.cc.u32 t47, t38, 0x3d1, t7;\n\t"
        "addc.u32 t48, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t49, t33, 0x3d1, t17;\n\t"
        "madc.hi.cc.u32 t50, t33, 0x3d1, t18;\n\t"
        "madc.lo.cc.u32 t51, t35, 0x3d1, t19;\n\t"
        "madc.hi.cc.u32 t52, t35, 0x3d1, t20;\n\t"
        "madc.lo.cc.u32 t53, t37, 0x3d1, t21;\n\t"
        "madc.hi.cc.u32 t54, t37, 0x3d1, t22;\n\t"
        "madc.lo.cc.u32 t55, t39, 0x3d1, t23;\n\t"
        "madc.hi.u32 t56, t39, 0x3d1, 0x0;\n\t"
        "add.cc.u32 t57, t41, t49;\n\t"
        "addc.cc.u32 t58, t42, t50;\n\t"
        "addc.cc.u32 t59, t43, t51;\n\t"
        "addc.cc.u32 t60, t44, t52;\n\t"
        "addc.cc.u32 t61, t45, t53;\n\t"
        "addc.cc.u32 t62, t46, t54;\n\t"
        "addc.cc.u32 t63, t47, t55;\n\t"
        "addc.u32 t64, t48, t56;\n\t"
        "add.cc.u32 t65, t57, t32;\n\t"
        "addc.cc.u32 t66, t58, t33;\n\t"
        "addc.cc.u32 t67, t59, t34;\n\t"
        "addc.cc.u32 t68, t60, t35;\n\t"
        "addc.cc.u32 t69, t61, t36;\n\t"
        "addc.cc.u32 t70, t62, t37;\n\t"
        "addc.cc.u32 t71, t63, t38;\n\t"
        "addc.cc.u32 t72, t64, t39;\n\t"
        "addc.u32 t73, 0x0, 0x0;\n\t"
        "mul.lo.u32 t74, t72, 0x3d1;\n\t"
        "mul.hi.u32 t75, t72, 0x3d1;\n\t"
        "mad.lo.u32 t76, t73, 0x3d1, t75;\n\t"
        "add.cc.u32 t77, t76, t72;\n\t"
        "addc.u32 t78, t73, 0x0;\n\t"
        "add.cc.u32 t79, t40, t74;\n\t"
        "addc.cc.u32 t80, t65, t77;\n\t"
        "addc.cc.u32 t81, t66, t78;\n\t"
        "addc.cc.u32 t82, t67, 0x0;\n\t"
        "addc.cc.u32 t83, t68, 0x0;\n\t"
        "addc.cc.u32 t84, t69, 0x0;\n\t"
        "addc.cc.u32 t85, t70, 0x0;\n\t"
        "addc.cc.u32 t86, t71, 0x0;\n\t"
        "addc.u32 t87, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t88, t87, 0x3d1, t79;\n\t"
        "addc.cc.u32 t89, t80, t87;\n\t"
        "addc.u32 t90, t81, 0x0;\n\t"
        "mov.u32 %0, t88;\n\t"
        "mov.u32 %1, t89;\n\t"
        "mov.u32 %2, t90;\n\t"
        "mov.u32 %3, t82;\n\t"
        "mov.u32 %4, t83;\n\t"
        "mov.u32 %5, t84;\n\t"
        "mov.u32 %6, t85;\n\t"
        "mov.u32 %7, t86;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90;
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
    w_ = (uint64_t)0x0u + 0x0u + cf_; t31 = (uint32_t)w_;
    w_ = (uint64_t)t8 + t24; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t25 + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t26 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t27 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t28 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t29 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + t30 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + t31 + cf_; t39 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t32 * 0x3d1u) + t0; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t32 * 0x3d1u) >> 32) + t1 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t34 * 0x3d1u) + t2 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t34 * 0x3d1u) >> 32) + t3 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t36 * 0x3d1u) + t4 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t36 * 0x3d1u) >> 32) + t5 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t38 * 0x3d1u) + t6 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t38 * 0x3d1u) >> 32) + t7 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t48 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t33 * 0x3d1u) + t17; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t33 * 0x3d1u) >> 32) + t18 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t35 * 0x3d1u) + t19 + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t35 * 0x3d1u) >> 32) + t20 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t37 * 0x3d1u) + t21 + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t37 * 0x3d1u) >> 32) + t22 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t39 * 0x3d1u) + t23 + cf_; t55 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t39 * 0x3d1u) >> 32) + 0x0u + cf_; t56 = (uint32_t)w_;
    w_ = (uint64_t)t41 + t49; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t42 + t50 + cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t43 + t51 + cf_; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t44 + t52 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t45 + t53 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t46 + t54 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t47 + t55 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t48 + t56 + cf_; t64 = (uint32_t)w_;
    w_ = (uint64_t)t57 + t32; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t58 + t33 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t59 + t34 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t60 + t35 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t61 + t36 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t62 + t37 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + t38 + cf_; t71 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t64 + t39 + cf_; t72 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t73 = (uint32_t)w_;
    t74 = (uint32_t)((uint32_t)(t72 * 0x3d1u));
    t75 = (uint32_t)(((uint64_t)t72 * 0x3d1u) >> 32);
    w_ = (uint64_t)(uint32_t)(t73 * 0x3d1u) + t75; t76 = (uint32_t)w_;
    w_ = (uint64_t)t76 + t72; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t73 + 0x0u + cf_; t78 = (uint32_t)w_;
    w_ = (uint64_t)t40 + t74; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t65 + t77 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t66 + t78 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t67 + 0x0u + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t68 + 0x0u + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t69 + 0x0u + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t70 + 0x0u + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t71 + 0x0u + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t87 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t87 * 0x3d1u) + t79; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t80 + t87 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t81 + 0x0u + cf_; t90 = (uint32_t)w_;
    r[0] = t88;
    r[1] = t89;
    r[2] = t90;
    r[3] = t82;
    r[4] = t83;
    r[5] = t84;
    r[6] = t85;
    r[7] = t86;
#endif
  }

  // c = a*a (pseudo.py:663-702 / monty.py:982-1165)
  static MAB_DEV void sqr(uint32_t (&r)[8], const uint32_t (&a)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<128>;\n\t"
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
        "addc.u32 t25, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t6, %10, %12, t6;\n\t"
        "madc.hi.cc.u32 t7, %10, %12, t7;\n\t"
        "madc.lo.cc.u32 t8, %10, %14, t8;\n\t"
        "madc.hi.cc.u32 t9, %10, %14, t9;\n\t"
        "addc.u32 t10, 0x0, 0x0;\n\t"
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
        "addc.u32 t27, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t10, %12, %14, t10;\n\t"
        "madc.hi.cc.u32 t11, %12, %14, t11;\n\t"
        "addc.u32 t12, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t25, %12, %13, t25;\n\t"
        "madc.hi.cc.u32 t26, %12, %13, t26;\n\t"
        "madc.lo.cc.u32 t27, %12, %15, t27;\n\t"
        "madc.hi.u32 t28, %12, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t12, %13, %15, t12;\n\t"
        "madc.hi.u32 t13, %13, %15, 0x0;\n\t"
        "mad.lo.cc.u32 t27, %13, %14, t27;\n\t"
        "madc.hi.cc.u32 t28, %13, %14, t28;\n\t"
        "addc.u32 t29, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t29, %14, %15, t29;\n\t"
        "madc.hi.u32 t30, %14, %15, 0x0;\n\t"
        "add.cc.u32 t32, t2, t18;\n\t"
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
        "addc.u32 t45, 0x0, 0x0;\n\t"
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
        "mad.lo.cc.u32 t77, t69, 0x3d1, t61;\n\t"
        "madc.hi.cc.u32 t78, t69, 0x3d1, t62;\n\t"
        "madc.lo.cc.u32 t79, t71, 0x3d1, t63;\n\t"
        "madc.hi.cc.u32 t80, t71, 0x3d1, t64;\n\t"
        "madc.lo.cc.u32 t81, t73, 0x3d1, t65;\n\t"
        "madc.hi.cc.u32 t82, t73, 0x3d1, t66;\n\t"
        "madc.lo.cc.u32 t83, t75, 0x3d1, t67;\n\t"
        "madc.hi.cc.u32 t84, t75, 0x3d1, t68;\n\t"
        "addc.u32 t85, 0x0, 0x0;\n\t"
        "mul.lo.u32 t86, t70, 0x3d1;\n\t"
        "mul.hi.u32 t87, t70, 0x3d1;\n\t"
        "mul.lo.u32 t88, t72, 0x3d1;\n\t"
        "mul.hi.u32 t89, t72, 0x3d1;\n\t"
        "mul.lo.u32 t90, t74, 0x3d1;\n\t"
        "mul.hi.u32 t91, t74, 0x3d1;\n\t"
        "mul.lo.u32 t92, t76, 0x3d1;\n\t"
        "mul.hi.u32 t93, t76, 0x3d1;\n\t"
        "add.cc.u32 t94, t78, t86;\n\t"
        "addc.cc.u32 t95, t79, t87;\n\t"
        "addc.cc.u32 t96, t80, t88;\n\t"
        "addc.cc.u32 t97, t81, t89;\n\t"
        "addc.cc.u32 t98, t82, t90;\n\t"
        "addc.cc.u32 t99, t83, t91;\n\t"
        "addc.cc.u32 t100, t84, t92;\n\t"
        "addc.u32 t101, t85, t93;\n\t"
        "add.cc.u32 t102, t94, t69;\n\t"
        "addc.cc.u32 t103, t95, t70;\n\t"
        "addc.cc.u32 t104, t96, t71;\n\t"
        "addc.cc.u32 t105, t97, t72;\n\t"
        "addc.cc.u32 t106, t98, t73;\n\t"
        "addc.cc.u32 t107, t99, t74;\n\t"
        "addc.cc.u32 t108, t100, t75;\n\t"
        "addc.cc.u32 t109, t101, t76;\n\t"
        "addc.u32 t110, 0x0, 0x0;\n\t"
        "mul.lo.u32 t111, t109, 0x3d1;\n\t"
        "mul.hi.u32 t112, t109, 0x3d1;\n\t"
        "mad.lo.u32 t113, t110, 0x3d1, t112;\n\t"
        "add.cc.u32 t114, t113, t109;\n\t"
        "addc.u32 t115, t110, 0x0;\n\t"
        "add.cc.u32 t116, t77, t111;\n\t"
        "addc.cc.u32 t117, t102, t114;\n\t"
        "addc.cc.u32 t118, t103, t115;\n\t"
        "addc.cc.u32 t119, t104, 0x0;\n\t"
        "addc.cc.u32 t120, t105, 0x0;\n\t"
        "addc.cc.u32 t121, t106, 0x0;\n\t"
        "addc.cc.u32 t122, t107, 0x0;\n\t"
        "addc.cc.u32 t123, t108, 0x0;\n\t"
        "addc.u32 t124, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t125, t124, 0x3d1, t116;\n\t"
        "addc.cc.u32 t126, t117, t124;\n\t"
        "addc.u32 t127, t118, 0x0;\n\t"
        "mov.u32 %0, t125;\n\t"
        "mov.u32 %1, t126;\n\t"
        "mov.u32 %2, t127;\n\t"
        "mov.u32 %3, t119;\n\t"
        "mov.u32 %4, t120;\n\t"
        "mov.u32 %5, t121;\n\t"
        "mov.u32 %6, t122;\n\t"
        "mov.u32 %7, t123;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127;
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
    w_ = (((uint64_t)a_6_i * a_7_i) >> 32) + 0x0u + cf_; t30 = (uint32_t)w_;
    w_ = (uint64_t)t2 + t18; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
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
    w_ = (uint64_t)(uint32_t)(t69 * 0x3d1u) + t61; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t69 * 0x3d1u) >> 32) + t62 + cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t71 * 0x3d1u) + t63 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t71 * 0x3d1u) >> 32) + t64 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t73 * 0x3d1u) + t65 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t73 * 0x3d1u) >> 32) + t66 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t75 * 0x3d1u) + t67 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t75 * 0x3d1u) >> 32) + t68 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t85 = (uint32_t)w_;
    t86 = (uint32_t)((uint32_t)(t70 * 0x3d1u));
    t87 = (uint32_t)(((uint64_t)t70 * 0x3d1u) >> 32);
    t88 = (uint32_t)((uint32_t)(t72 * 0x3d1u));
    t89 = (uint32_t)(((uint64_t)t72 * 0x3d1u) >> 32);
    t90 = (uint32_t)((uint32_t)(t74 * 0x3d1u));
    t91 = (uint32_t)(((uint64_t)t74 * 0x3d1u) >> 32);
    t92 = (uint32_t)((uint32_t)(t76 * 0x3d1u));
    t93 = (uint32_t)(((uint64_t)t76 * 0x3d1u) >> 32);
    w_ = (uint64_t)t78 + t86; t94 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t79 + t87 + cf_; t95 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t80 + t88 + cf_; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t81 + t89 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t82 + t90 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t83 + t91 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t84 + t92 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t85 + t93 + cf_; t101 = (uint32_t)w_;
    w_ = (uint64_t)t94 + t69; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t95 + t70 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t96 + t71 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t97 + t72 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t98 + t73 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t99 + t74 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t100 + t75 + cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t101 + t76 + cf_; t109 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t110 = (uint32_t)w_;
    t111 = (uint32_t)((uint32_t)(t109 * 0x3d1u));
    t112 = (uint32_t)(((uint64_t)t109 * 0x3d1u) >> 32);
    w_ = (uint64_t)(uint32_t)(t110 * 0x3d1u) + t112; t113 = (uint32_t)w_;
    w_ = (uint64_t)t113 + t109; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t110 + 0x0u + cf_; t115 = (uint32_t)w_;
    w_ = (uint64_t)t77 + t111; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t102 + t114 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t103 + t115 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t104 + 0x0u + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t105 + 0x0u + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t106 + 0x0u + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t107 + 0x0u + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t108 + 0x0u + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t124 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t124 * 0x3d1u) + t116; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t117 + t124 + cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t118 + 0x0u + cf_; t127 = (uint32_t)w_;
    r[0] = t125;
    r[1] = t126;
    r[2] = t127;
    r[3] = t119;
    r[4] = t120;
    r[5] = t121;
    r[6] = t122;
    r[7] = t123;
#endif
  }

  // c = a*b for a small integer b (pseudo.py:705-728 / monty.py:876-978)
  static MAB_DEV void mli(uint32_t (&r)[8], const uint32_t (&a)[8], uint32_t b) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<40>;\n\t"
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
        "mul.lo.u32 t24, t23, 0x3d1;\n\t"
        "mul.hi.u32 t25, t23, 0x3d1;\n\t"
        "add.cc.u32 t26, t25, t23;\n\t"
        "addc.u32 t27, 0x0, 0x0;\n\t"
        "add.cc.u32 t28, t0, t24;\n\t"
        "addc.cc.u32 t29, t16, t26;\n\t"
        "addc.cc.u32 t30, t17, t27;\n\t"
        "addc.cc.u32 t31, t18, 0x0;\n\t"
        "addc.cc.u32 t32, t19, 0x0;\n\t"
        "addc.cc.u32 t33, t20, 0x0;\n\t"
        "addc.cc.u32 t34, t21, 0x0;\n\t"
        "addc.cc.u32 t35, t22, 0x0;\n\t"
        "addc.u32 t36, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t37, t36, 0x3d1, t28;\n\t"
        "addc.cc.u32 t38, t29, t36;\n\t"
        "addc.u32 t39, t30, 0x0;\n\t"
        "mov.u32 %0, t37;\n\t"
        "mov.u32 %1, t38;\n\t"
        "mov.u32 %2, t39;\n\t"
        "mov.u32 %3, t31;\n\t"
        "mov.u32 %4, t32;\n\t"
        "mov.u32 %5, t33;\n\t"
        "mov.u32 %6, t34;\n\t"
        "mov.u32 %7, t35;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39;
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
    t24 = (uint32_t)((uint32_t)(t23 * 0x3d1u));
    t25 = (uint32_t)(((uint64_t)t23 * 0x3d1u) >> 32);
    w_ = (uint64_t)t25 + t23; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t24; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t16 + t26 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t17 + t27 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t18 + 0x0u + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t19 + 0x0u + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t20 + 0x0u + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t21 + 0x0u + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t22 + 0x0u + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t36 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t36 * 0x3d1u) + t28; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t29 + t36 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t30 + 0x0u + cf_; t39 = (uint32_t)w_;
    r[0] = t37;
    r[1] = t38;
    r[2] = t39;
    r[3] = t31;
    r[4] = t32;
    r[5] = t33;
    r[6] = t34;
    r[7] = t35;
#endif
  }

  // r = a*b + c, small integer b: modmli + modadd fused (rfc7748.c:209,212)
  static MAB_DEV void mla(uint32_t (&r)[8], const uint32_t (&a)[8], uint32_t b, const uint32_t (&c)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<41>;\n\t"
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
        "mul.lo.u32 t25, t24, 0x3d1;\n\t"
        "mul.hi.u32 t26, t24, 0x3d1;\n\t"
        "add.cc.u32 t27, t26, t24;\n\t"
        "addc.u32 t28, 0x0, 0x0;\n\t"
        "add.cc.u32 t29, t0, t25;\n\t"
        "addc.cc.u32 t30, t17, t27;\n\t"
        "addc.cc.u32 t31, t18, t28;\n\t"
        "addc.cc.u32 t32, t19, 0x0;\n\t"
        "addc.cc.u32 t33, t20, 0x0;\n\t"
        "addc.cc.u32 t34, t21, 0x0;\n\t"
        "addc.cc.u32 t35, t22, 0x0;\n\t"
        "addc.cc.u32 t36, t23, 0x0;\n\t"
        "addc.u32 t37, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t38, t37, 0x3d1, t29;\n\t"
        "addc.cc.u32 t39, t30, t37;\n\t"
        "addc.u32 t40, t31, 0x0;\n\t"
        "mov.u32 %0, t38;\n\t"
        "mov.u32 %1, t39;\n\t"
        "mov.u32 %2, t40;\n\t"
        "mov.u32 %3, t32;\n\t"
        "mov.u32 %4, t33;\n\t"
        "mov.u32 %5, t34;\n\t"
        "mov.u32 %6, t35;\n\t"
        "mov.u32 %7, t36;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40;
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
    t25 = (uint32_t)((uint32_t)(t24 * 0x3d1u));
    t26 = (uint32_t)(((uint64_t)t24 * 0x3d1u) >> 32);
    w_ = (uint64_t)t26 + t24; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t28 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t25; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t17 + t27 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t18 + t28 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t19 + 0x0u + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t20 + 0x0u + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t21 + 0x0u + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t22 + 0x0u + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t23 + 0x0u + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t37 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t37 * 0x3d1u) + t29; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t30 + t37 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t31 + 0x0u + cf_; t40 = (uint32_t)w_;
    r[0] = t38;
    r[1] = t39;
    r[2] = t40;
    r[3] = t32;
    r[4] = t33;
    r[5] = t34;
    r[6] = t35;
    r[7] = t36;
#endif
  }

  // n = a+b (pseudo.py:286-304)
  static MAB_DEV void add(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<22>;\n\t"
        "add.cc.u32 t0, %8, %16;\n\t"
        "addc.cc.u32 t1, %9, %17;\n\t"
        "addc.cc.u32 t2, %10, %18;\n\t"
        "addc.cc.u32 t3, %11, %19;\n\t"
        "addc.cc.u32 t4, %12, %20;\n\t"
        "addc.cc.u32 t5, %13, %21;\n\t"
        "addc.cc.u32 t6, %14, %22;\n\t"
        "addc.cc.u32 t7, %15, %23;\n\t"
        "addc.u32 t8, 0x0, 0x0;\n\t"
        "mad.lo.u32 t9, t8, 0x3d1, 0x0;\n\t"
        "add.cc.u32 t10, t0, t9;\n\t"
        "addc.cc.u32 t11, t1, t8;\n\t"
        "addc.cc.u32 t12, t2, 0x0;\n\t"
        "addc.cc.u32 t13, t3, 0x0;\n\t"
        "addc.cc.u32 t14, t4, 0x0;\n\t"
        "addc.cc.u32 t15, t5, 0x0;\n\t"
        "addc.cc.u32 t16, t6, 0x0;\n\t"
        "addc.cc.u32 t17, t7, 0x0;\n\t"
        "addc.u32 t18, 0x0, 0x0;\n\t"
        "mad.lo.u32 t19, t18, 0x3d1, 0x0;\n\t"
        "add.cc.u32 t20, t10, t19;\n\t"
        "addc.u32 t21, t11, t18;\n\t"
        "mov.u32 %0, t20;\n\t"
        "mov.u32 %1, t21;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21;
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
    w_ = (uint64_t)(uint32_t)(t8 * 0x3d1u) + 0x0u; t9 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t9; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t1 + t8 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + 0x0u + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + 0x0u + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + 0x0u + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + 0x0u + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + 0x0u + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + 0x0u + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t18 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t18 * 0x3d1u) + 0x0u; t19 = (uint32_t)w_;
    w_ = (uint64_t)t10 + t19; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t18 + cf_; t21 = (uint32_t)w_;
    r[0] = t20;
    r[1] = t21;
    r[2] = t12;
    r[3] = t13;
    r[4] = t14;
    r[5] = t15;
    r[6] = t16;
    r[7] = t17;
#endif
  }

  // n = a-b (pseudo.py:307-326)
  static MAB_DEV void sub(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<24>;\n\t"
        "sub.cc.u32 t0, %8, %16;\n\t"
        "subc.cc.u32 t1, %9, %17;\n\t"
        "subc.cc.u32 t2, %10, %18;\n\t"
        "subc.cc.u32 t3, %11, %19;\n\t"
        "subc.cc.u32 t4, %12, %20;\n\t"
        "subc.cc.u32 t5, %13, %21;\n\t"
        "subc.cc.u32 t6, %14, %22;\n\t"
        "subc.cc.u32 t7, %15, %23;\n\t"
        "subc.u32 t8, 0x0, 0x0;\n\t"
        "and.b32 t9, t8, 0x3d1;\n\t"
        "and.b32 t10, t8, 0x1;\n\t"
        "sub.cc.u32 t11, t0, t9;\n\t"
        "subc.cc.u32 t12, t1, t10;\n\t"
        "subc.cc.u32 t13, t2, 0x0;\n\t"
        "subc.cc.u32 t14, t3, 0x0;\n\t"
        "subc.cc.u32 t15, t4, 0x0;\n\t"
        "subc.cc.u32 t16, t5, 0x0;\n\t"
        "subc.cc.u32 t17, t6, 0x0;\n\t"
        "subc.cc.u32 t18, t7, 0x0;\n\t"
        "subc.u32 t19, 0x0, 0x0;\n\t"
        "and.b32 t20, t19, 0x3d1;\n\t"
        "and.b32 t21, t19, 0x1;\n\t"
        "sub.cc.u32 t22, t11, t20;\n\t"
        "subc.u32 t23, t12, t21;\n\t"
        "mov.u32 %0, t22;\n\t"
        "mov.u32 %1, t23;\n\t"
        "mov.u32 %2, t13;\n\t"
        "mov.u32 %3, t14;\n\t"
        "mov.u32 %4, t15;\n\t"
        "mov.u32 %5, t16;\n\t"
        "mov.u32 %6, t17;\n\t"
        "mov.u32 %7, t18;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23;
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
    t9 = (uint32_t)(t8 & 0x3d1u);
    t10 = (uint32_t)(t8 & 0x1u);
    w_ = (uint64_t)t0 - t9; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t1 - t10 - cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t2 - 0x0u - cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t3 - 0x0u - cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t4 - 0x0u - cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t5 - 0x0u - cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t6 - 0x0u - cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t7 - 0x0u - cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t19 = (uint32_t)w_;
    t20 = (uint32_t)(t19 & 0x3d1u);
    t21 = (uint32_t)(t19 & 0x1u);
    w_ = (uint64_t)t11 - t20; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t12 - t21 - cf_; t23 = (uint32_t)w_;
    r[0] = t22;
    r[1] = t23;
    r[2] = t13;
    r[3] = t14;
    r[4] = t15;
    r[5] = t16;
    r[6] = t17;
    r[7] = t18;
#endif
  }

  // no spare bit above Nbits in this plan: the product-operand forms are the general ones
  static constexpr bool TIGHT = false;
  static MAB_DEV void add_tt(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { add(r, a, b); }
  static MAB_DEV void sub_tt(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { sub(r, a, b); }

  // no separate weakly-reduced products in this plan: chains use the ordinary ones
  static constexpr bool WEAK = false;
  static MAB_DEV void mul_w(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { mul(r, a, b); }
  static MAB_DEV void sqr_w(uint32_t (&r)[8], const uint32_t (&a)[8]) { sqr(r, a); }

  // n = -b (pseudo.py:329-348)
  static MAB_DEV void neg(uint32_t (&r)[8], const uint32_t (&b)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<24>;\n\t"
        "sub.cc.u32 t0, 0x0, %8;\n\t"
        "subc.cc.u32 t1, 0x0, %9;\n\t"
        "subc.cc.u32 t2, 0x0, %10;\n\t"
        "subc.cc.u32 t3, 0x0, %11;\n\t"
        "subc.cc.u32 t4, 0x0, %12;\n\t"
        "subc.cc.u32 t5, 0x0, %13;\n\t"
        "subc.cc.u32 t6, 0x0, %14;\n\t"
        "subc.cc.u32 t7, 0x0, %15;\n\t"
        "subc.u32 t8, 0x0, 0x0;\n\t"
        "and.b32 t9, t8, 0x3d1;\n\t"
        "and.b32 t10, t8, 0x1;\n\t"
        "sub.cc.u32 t11, t0, t9;\n\t"
        "subc.cc.u32 t12, t1, t10;\n\t"
        "subc.cc.u32 t13, t2, 0x0;\n\t"
        "subc.cc.u32 t14, t3, 0x0;\n\t"
        "subc.cc.u32 t15, t4, 0x0;\n\t"
        "subc.cc.u32 t16, t5, 0x0;\n\t"
        "subc.cc.u32 t17, t6, 0x0;\n\t"
        "subc.cc.u32 t18, t7, 0x0;\n\t"
        "subc.u32 t19, 0x0, 0x0;\n\t"
        "and.b32 t20, t19, 0x3d1;\n\t"
        "and.b32 t21, t19, 0x1;\n\t"
        "sub.cc.u32 t22, t11, t20;\n\t"
        "subc.u32 t23, t12, t21;\n\t"
        "mov.u32 %0, t22;\n\t"
        "mov.u32 %1, t23;\n\t"
        "mov.u32 %2, t13;\n\t"
        "mov.u32 %3, t14;\n\t"
        "mov.u32 %4, t15;\n\t"
        "mov.u32 %5, t16;\n\t"
        "mov.u32 %6, t17;\n\t"
        "mov.u32 %7, t18;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23;
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
    t9 = (uint32_t)(t8 & 0x3d1u);
    t10 = (uint32_t)(t8 & 0x1u);
    w_ = (uint64_t)t0 - t9; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t1 - t10 - cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t2 - 0x0u - cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t3 - 0x0u - cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t4 - 0x0u - cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t5 - 0x0u - cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t6 - 0x0u - cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t7 - 0x0u - cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t19 = (uint32_t)w_;
    t20 = (uint32_t)(t19 & 0x3d1u);
    t21 = (uint32_t)(t19 & 0x1u);
    w_ = (uint64_t)t11 - t20; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t12 - t21 - cf_; t23 = (uint32_t)w_;
    r[0] = t22;
    r[1] = t23;
    r[2] = t13;
    r[3] = t14;
    r[4] = t15;
    r[5] = t16;
    r[6] = t17;
    r[7] = t18;
#endif
  }

  // canonical residue of a stored value; returns 1 iff it was already < p
  // (flatten/modfsb, pseudo.py:255-283)
  static MAB_DEV uint32_t canon(uint32_t (&r)[8], const uint32_t (&a)[8]) {
    uint32_t lt;
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<34>;\n\t"
        "sub.cc.u32 t1, %9, 0xfffffc2f;\n\t"
        "subc.cc.u32 t2, %10, 0xfffffffe;\n\t"
        "subc.cc.u32 t3, %11, 0xffffffff;\n\t"
        "subc.cc.u32 t4, %12, 0xffffffff;\n\t"
        "subc.cc.u32 t5, %13, 0xffffffff;\n\t"
        "subc.cc.u32 t6, %14, 0xffffffff;\n\t"
        "subc.cc.u32 t7, %15, 0xffffffff;\n\t"
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
    w_ = (uint64_t)a_0_i - 0xfffffc2fu; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_1_i - 0xfffffffeu - cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_2_i - 0xffffffffu - cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_3_i - 0xffffffffu - cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_4_i - 0xffffffffu - cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_5_i - 0xffffffffu - cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_6_i - 0xffffffffu - cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
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

  static MAB_DEV void set_p(uint32_t (&r)[8]) { r[0] = 0xfffffc2fu; r[1] = 0xfffffffeu; r[2] = 0xffffffffu; r[3] = 0xffffffffu; r[4] = 0xffffffffu; r[5] = 0xffffffffu; r[6] = 0xffffffffu; r[7] = 0xffffffffu; }
  static MAB_DEV void set_one(uint32_t (&r)[8]) { r[0] = 0x00000001u; r[1] = 0x00000000u; r[2] = 0x00000000u; r[3] = 0x00000000u; r[4] = 0x00000000u; r[5] = 0x00000000u; r[6] = 0x00000000u; r[7] = 0x00000000u; }
  static MAB_DEV void set_roi(uint32_t (&r)[8]) { r[0] = 0xfffffc2eu; r[1] = 0xfffffffeu; r[2] = 0xffffffffu; r[3] = 0xffffffffu; r[4] = 0xffffffffu; r[5] = 0xffffffffu; r[6] = 0xffffffffu; r[7] = 0xffffffffu; }
  static MAB_DEV void set_r2(uint32_t (&r)[8]) { r[0] = 0x00000001u; r[1] = 0x00000000u; r[2] = 0x00000000u; r[3] = 0x00000000u; r[4] = 0x00000000u; r[5] = 0x00000000u; r[6] = 0x00000000u; r[7] = 0x00000000u; }
  static constexpr bool HAS_WEIERSTRASS = false;

  // nres: copy (pseudo.py:952-962); redc: copy + final subtract (pseudo.py:965-976)
  static MAB_DEV void nres(uint32_t (&r)[8], const uint32_t (&a)[8]) { for (int i = 0; i < L; i++) r[i] = a[i]; }
  static MAB_DEV void redc(uint32_t (&r)[8], const uint32_t (&a)[8]) { (void)canon(r, a); }

  // z = w^PE, straight-line addition chain (pseudo.py:758-785; our own chain finder)
  static MAB_DEV void pro(uint32_t (&z)[8], const uint32_t (&w)[8]) {
    uint32_t x[L];
    for (int i = 0; i < L; i++) x[i] = w[i];
    uint32_t t0[L];
    uint32_t t1[L];
    uint32_t t2[L];
    uint32_t t3[L];
    uint32_t t4[L];
    uint32_t t5[L];
    sqr_w(t0, x);
    mul_w(t0, t0, x);
    sqr_w(t1, t0);
    mul_w(t1, t1, x);
    sqr_w(t2, t1);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr_w(t2, t2);
    mul_w(t2, t2, t1);
    sqr_w(t3, t2);
    MAB_NOUNROLL
    for (int i = 1; i < 6; i++) sqr_w(t3, t3);
    mul_w(t3, t3, t2);
    sqr_w(t3, t3);
    mul_w(t3, t3, x);
    sqr_w(t4, t3);
    MAB_NOUNROLL
    for (int i = 1; i < 13; i++) sqr_w(t4, t4);
    mul_w(t4, t4, t3);
    sqr_w(t4, t4);
    mul_w(t4, t4, x);
    sqr_w(t5, t4);
    MAB_NOUNROLL
    for (int i = 1; i < 27; i++) sqr_w(t5, t5);
    mul_w(t5, t5, t4);
    sqr_w(t5, t5);
    mul_w(t5, t5, x);
    sqr_w(t4, t5);
    MAB_NOUNROLL
    for (int i = 1; i < 55; i++) sqr_w(t4, t4);
    mul_w(t4, t4, t5);
    sqr_w(t4, t4);
    mul_w(t4, t4, x);
    sqr_w(t5, t4);
    MAB_NOUNROLL
    for (int i = 1; i < 111; i++) sqr_w(t5, t5);
    mul_w(t5, t5, t4);
    sqr_w(t5, t5);
    mul_w(t5, t5, x);
    sqr_w(z, t5);
    MAB_NOUNROLL
    for (int i = 1; i < 14; i++) sqr_w(z, z);
    mul_w(z, z, t3);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 6; i++) sqr_w(z, z);
    mul_w(z, z, t2);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr_w(z, z);
    mul_w(z, z, t1);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    if (WEAK) (void)canon(z, z);
  }
};
