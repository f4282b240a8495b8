// Automatically generated field arithmetic for sm_100a -- do not edit.
// Command line : python -m modarith_b200.gen.monty_sm100 NIST256ORDER
// modulus NIST256ORDER = 0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551
// plan MontgomeryFull: 8 saturated 32-bit limbs; stored values < p; R = 2^256
//   mul   : 128 IMAD.WIDE   8 IMAD  ~ 73 ALU-pipe ops
//   sqr   : 128 IMAD.WIDE   8 IMAD  ~ 73 ALU-pipe ops
//   mli   : 200 IMAD.WIDE  16 IMAD  ~132 ALU-pipe ops
//   mla   : 200 IMAD.WIDE  16 IMAD  ~175 ALU-pipe ops
//   add   :   0 IMAD.WIDE   0 IMAD  ~ 43 ALU-pipe ops
//   sub   :   0 IMAD.WIDE   0 IMAD  ~ 21 ALU-pipe ops
//   canon :   0 IMAD.WIDE   0 IMAD  ~ 34 ALU-pipe ops
//   modpro: 250 squarings + 46 multiplies (exponent (p-1-2^k)/2^(k+1), k=4)
#pragma once
#include "mab_common.cuh"

struct F_NIST256ORDER {
  static constexpr int L = 8;
  static constexpr int NBITS = 256;
  static constexpr int NBYTES = 32;
  static constexpr int PM1D2 = 4;
  static constexpr bool MONTGOMERY = true;
  static constexpr int PRO_SQR = 250, PRO_MUL = 46;
  static constexpr int LADDER_MINBLOCKS = 3;   // resident 128-thread CTAs per SM for k_rfc7748
  static constexpr bool LADDER_STASH = false;   // scalar and x1 in shared memory (see rfc7748_sm100.cuh)
  static constexpr bool HAS_CURVE = false;
  static constexpr uint32_t A24 = 0;
  static constexpr int COF = 0;
  static constexpr uint32_t GENERATOR = 0;
  static const char* name() { return "NIST256ORDER"; }

  // c = a*b (pseudo.py:616-659 / monty.py:663-872)
  static MAB_DEV void mul(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<94>;\n\t"
        "mul.lo.u32 t0, %8, %16;\n\t"
        "mul.hi.u32 t1, %8, %16;\n\t"
        "mul.lo.u32 t2, %10, %16;\n\t"
        "mul.hi.u32 t3, %10, %16;\n\t"
        "mul.lo.u32 t4, %12, %16;\n\t"
        "mul.hi.u32 t5, %12, %16;\n\t"
        "mul.lo.u32 t6, %14, %16;\n\t"
        "mul.hi.u32 t7, %14, %16;\n\t"
        "mul.lo.u32 t19, %9, %16;\n\t"
        "mul.hi.u32 t20, %9, %16;\n\t"
        "mul.lo.u32 t21, %11, %16;\n\t"
        "mul.hi.u32 t22, %11, %16;\n\t"
        "mul.lo.u32 t23, %13, %16;\n\t"
        "mul.hi.u32 t24, %13, %16;\n\t"
        "mul.lo.u32 t25, %15, %16;\n\t"
        "mul.hi.u32 t26, %15, %16;\n\t"
        "mul.lo.u32 t36, t0, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t19, 0xf3b9cac2, t36, t19;\n\t"
        "madc.hi.cc.u32 t20, 0xf3b9cac2, t36, t20;\n\t"
        "madc.lo.cc.u32 t21, 0xbce6faad, t36, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xbce6faad, t36, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xffffffff, t36, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xffffffff, t36, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t36, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t36, t26;\n\t"
        "addc.u32 t27, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t0, 0xfc632551, t36, t0;\n\t"
        "madc.hi.cc.u32 t1, 0xfc632551, t36, t1;\n\t"
        "madc.lo.cc.u32 t2, 0xa7179e84, t36, t2;\n\t"
        "madc.hi.cc.u32 t3, 0xa7179e84, t36, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xffffffff, t36, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xffffffff, t36, t5;\n\t"
        "madc.lo.cc.u32 t6, 0x0, t36, t6;\n\t"
        "madc.hi.cc.u32 t7, 0x0, t36, t7;\n\t"
        "addc.u32 t8, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t19, %8, %17, t19;\n\t"
        "madc.hi.cc.u32 t20, %8, %17, t20;\n\t"
        "madc.lo.cc.u32 t21, %10, %17, t21;\n\t"
        "madc.hi.cc.u32 t22, %10, %17, t22;\n\t"
        "madc.lo.cc.u32 t23, %12, %17, t23;\n\t"
        "madc.hi.cc.u32 t24, %12, %17, t24;\n\t"
        "madc.lo.cc.u32 t25, %14, %17, t25;\n\t"
        "madc.hi.cc.u32 t26, %14, %17, t26;\n\t"
        "addc.u32 t27, t27, 0x0;\n\t"
        "mad.lo.cc.u32 t2, %9, %17, t2;\n\t"
        "madc.hi.cc.u32 t3, %9, %17, t3;\n\t"
        "madc.lo.cc.u32 t4, %11, %17, t4;\n\t"
        "madc.hi.cc.u32 t5, %11, %17, t5;\n\t"
        "madc.lo.cc.u32 t6, %13, %17, t6;\n\t"
        "madc.hi.cc.u32 t7, %13, %17, t7;\n\t"
        "madc.lo.cc.u32 t8, %15, %17, t8;\n\t"
        "madc.hi.u32 t9, %15, %17, 0x0;\n\t"
        "add.cc.u32 t37, t19, t1;\n\t"
        "mul.lo.u32 t38, t37, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t2, 0xf3b9cac2, t38, t2;\n\t"
        "madc.hi.cc.u32 t3, 0xf3b9cac2, t38, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xbce6faad, t38, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xbce6faad, t38, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xffffffff, t38, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xffffffff, t38, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t38, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xffffffff, t38, t9;\n\t"
        "addc.u32 t10, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t37, 0xfc632551, t38, t37;\n\t"
        "madc.hi.cc.u32 t20, 0xfc632551, t38, t20;\n\t"
        "madc.lo.cc.u32 t21, 0xa7179e84, t38, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xa7179e84, t38, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xffffffff, t38, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xffffffff, t38, t24;\n\t"
        "madc.lo.cc.u32 t25, 0x0, t38, t25;\n\t"
        "madc.hi.cc.u32 t26, 0x0, t38, t26;\n\t"
        "addc.u32 t27, t27, 0x0;\n\t"
        "mad.lo.cc.u32 t2, %8, %18, t2;\n\t"
        "madc.hi.cc.u32 t3, %8, %18, t3;\n\t"
        "madc.lo.cc.u32 t4, %10, %18, t4;\n\t"
        "madc.hi.cc.u32 t5, %10, %18, t5;\n\t"
        "madc.lo.cc.u32 t6, %12, %18, t6;\n\t"
        "madc.hi.cc.u32 t7, %12, %18, t7;\n\t"
        "madc.lo.cc.u32 t8, %14, %18, t8;\n\t"
        "madc.hi.cc.u32 t9, %14, %18, t9;\n\t"
        "addc.u32 t10, t10, 0x0;\n\t"
        "mad.lo.cc.u32 t21, %9, %18, t21;\n\t"
        "madc.hi.cc.u32 t22, %9, %18, t22;\n\t"
        "madc.lo.cc.u32 t23, %11, %18, t23;\n\t"
        "madc.hi.cc.u32 t24, %11, %18, t24;\n\t"
        "madc.lo.cc.u32 t25, %13, %18, t25;\n\t"
        "madc.hi.cc.u32 t26, %13, %18, t26;\n\t"
        "madc.lo.cc.u32 t27, %15, %18, t27;\n\t"
        "madc.hi.u32 t28, %15, %18, 0x0;\n\t"
        "add.cc.u32 t39, t2, t20;\n\t"
        "mul.lo.u32 t40, t39, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t21, 0xf3b9cac2, t40, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xf3b9cac2, t40, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xbce6faad, t40, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xbce6faad, t40, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t40, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t40, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t40, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xffffffff, t40, t28;\n\t"
        "addc.u32 t29, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t39, 0xfc632551, t40, t39;\n\t"
        "madc.hi.cc.u32 t3, 0xfc632551, t40, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xa7179e84, t40, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xa7179e84, t40, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xffffffff, t40, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xffffffff, t40, t7;\n\t"
        "madc.lo.cc.u32 t8, 0x0, t40, t8;\n\t"
        "madc.hi.cc.u32 t9, 0x0, t40, t9;\n\t"
        "addc.u32 t10, t10, 0x0;\n\t"
        "mad.lo.cc.u32 t21, %8, %19, t21;\n\t"
        "madc.hi.cc.u32 t22, %8, %19, t22;\n\t"
        "madc.lo.cc.u32 t23, %10, %19, t23;\n\t"
        "madc.hi.cc.u32 t24, %10, %19, t24;\n\t"
        "madc.lo.cc.u32 t25, %12, %19, t25;\n\t"
        "madc.hi.cc.u32 t26, %12, %19, t26;\n\t"
        "madc.lo.cc.u32 t27, %14, %19, t27;\n\t"
        "madc.hi.cc.u32 t28, %14, %19, t28;\n\t"
        "addc.u32 t29, t29, 0x0;\n\t"
        "mad.lo.cc.u32 t4, %9, %19, t4;\n\t"
        "madc.hi.cc.u32 t5, %9, %19, t5;\n\t"
        "madc.lo.cc.u32 t6, %11, %19, t6;\n\t"
        "madc.hi.cc.u32 t7, %11, %19, t7;\n\t"
        "madc.lo.cc.u32 t8, %13, %19, t8;\n\t"
        "madc.hi.cc.u32 t9, %13, %19, t9;\n\t"
        "madc.lo.cc.u32 t10, %15, %19, t10;\n\t"
        "madc.hi.u32 t11, %15, %19, 0x0;\n\t"
        "add.cc.u32 t41, t21, t3;\n\t"
        "mul.lo.u32 t42, t41, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t4, 0xf3b9cac2, t42, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xf3b9cac2, t42, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xbce6faad, t42, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xbce6faad, t42, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t42, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xffffffff, t42, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t42, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xffffffff, t42, t11;\n\t"
        "addc.u32 t12, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t41, 0xfc632551, t42, t41;\n\t"
        "madc.hi.cc.u32 t22, 0xfc632551, t42, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xa7179e84, t42, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xa7179e84, t42, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t42, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t42, t26;\n\t"
        "madc.lo.cc.u32 t27, 0x0, t42, t27;\n\t"
        "madc.hi.cc.u32 t28, 0x0, t42, t28;\n\t"
        "addc.u32 t29, t29, 0x0;\n\t"
        "mad.lo.cc.u32 t4, %8, %20, t4;\n\t"
        "madc.hi.cc.u32 t5, %8, %20, t5;\n\t"
        "madc.lo.cc.u32 t6, %10, %20, t6;\n\t"
        "madc.hi.cc.u32 t7, %10, %20, t7;\n\t"
        "madc.lo.cc.u32 t8, %12, %20, t8;\n\t"
        "madc.hi.cc.u32 t9, %12, %20, t9;\n\t"
        "madc.lo.cc.u32 t10, %14, %20, t10;\n\t"
        "madc.hi.cc.u32 t11, %14, %20, t11;\n\t"
        "addc.u32 t12, t12, 0x0;\n\t"
        "mad.lo.cc.u32 t23, %9, %20, t23;\n\t"
        "madc.hi.cc.u32 t24, %9, %20, t24;\n\t"
        "madc.lo.cc.u32 t25, %11, %20, t25;\n\t"
        "madc.hi.cc.u32 t26, %11, %20, t26;\n\t"
        "madc.lo.cc.u32 t27, %13, %20, t27;\n\t"
        "madc.hi.cc.u32 t28, %13, %20, t28;\n\t"
        "madc.lo.cc.u32 t29, %15, %20, t29;\n\t"
        "madc.hi.u32 t30, %15, %20, 0x0;\n\t"
        "add.cc.u32 t43, t4, t22;\n\t"
        "mul.lo.u32 t44, t43, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t23, 0xf3b9cac2, t44, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xf3b9cac2, t44, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xbce6faad, t44, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xbce6faad, t44, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t44, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xffffffff, t44, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t44, t29;\n\t"
        "madc.hi.cc.u32 t30, 0xffffffff, t44, t30;\n\t"
        "addc.u32 t31, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t43, 0xfc632551, t44, t43;\n\t"
        "madc.hi.cc.u32 t5, 0xfc632551, t44, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xa7179e84, t44, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xa7179e84, t44, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t44, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xffffffff, t44, t9;\n\t"
        "madc.lo.cc.u32 t10, 0x0, t44, t10;\n\t"
        "madc.hi.cc.u32 t11, 0x0, t44, t11;\n\t"
        "addc.u32 t12, t12, 0x0;\n\t"
        "mad.lo.cc.u32 t23, %8, %21, t23;\n\t"
        "madc.hi.cc.u32 t24, %8, %21, t24;\n\t"
        "madc.lo.cc.u32 t25, %10, %21, t25;\n\t"
        "madc.hi.cc.u32 t26, %10, %21, t26;\n\t"
        "madc.lo.cc.u32 t27, %12, %21, t27;\n\t"
        "madc.hi.cc.u32 t28, %12, %21, t28;\n\t"
        "madc.lo.cc.u32 t29, %14, %21, t29;\n\t"
        "madc.hi.cc.u32 t30, %14, %21, t30;\n\t"
        "addc.u32 t31, t31, 0x0;\n\t"
        "mad.lo.cc.u32 t6, %9, %21, t6;\n\t"
        "madc.hi.cc.u32 t7, %9, %21, t7;\n\t"
        "madc.lo.cc.u32 t8, %11, %21, t8;\n\t"
        "madc.hi.cc.u32 t9, %11, %21, t9;\n\t"
        "madc.lo.cc.u32 t10, %13, %21, t10;\n\t"
        "madc.hi.cc.u32 t11, %13, %21, t11;\n\t"
        "madc.lo.cc.u32 t12, %15, %21, t12;\n\t"
        "madc.hi.u32 t13, %15, %21, 0x0;\n\t"
        "add.cc.u32 t45, t23, t5;\n\t"
        "mul.lo.u32 t46, t45, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t6, 0xf3b9cac2, t46, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xf3b9cac2, t46, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xbce6faad, t46, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xbce6faad, t46, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t46, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xffffffff, t46, t11;\n\t"
        "madc.lo.cc.u32 t12, 0xffffffff, t46, t12;\n\t"
        "madc.hi.cc.u32 t13, 0xffffffff, t46, t13;\n\t"
        "addc.u32 t14, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t45, 0xfc632551, t46, t45;\n\t"
        "madc.hi.cc.u32 t24, 0xfc632551, t46, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xa7179e84, t46, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xa7179e84, t46, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t46, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xffffffff, t46, t28;\n\t"
        "madc.lo.cc.u32 t29, 0x0, t46, t29;\n\t"
        "madc.hi.cc.u32 t30, 0x0, t46, t30;\n\t"
        "addc.u32 t31, t31, 0x0;\n\t"
        "mad.lo.cc.u32 t6, %8, %22, t6;\n\t"
        "madc.hi.cc.u32 t7, %8, %22, t7;\n\t"
        "madc.lo.cc.u32 t8, %10, %22, t8;\n\t"
        "madc.hi.cc.u32 t9, %10, %22, t9;\n\t"
        "madc.lo.cc.u32 t10, %12, %22, t10;\n\t"
        "madc.hi.cc.u32 t11, %12, %22, t11;\n\t"
        "madc.lo.cc.u32 t12, %14, %22, t12;\n\t"
        "madc.hi.cc.u32 t13, %14, %22, t13;\n\t"
        "addc.u32 t14, t14, 0x0;\n\t"
        "mad.lo.cc.u32 t25, %9, %22, t25;\n\t"
        "madc.hi.cc.u32 t26, %9, %22, t26;\n\t"
        "madc.lo.cc.u32 t27, %11, %22, t27;\n\t"
        "madc.hi.cc.u32 t28, %11, %22, t28;\n\t"
        "madc.lo.cc.u32 t29, %13, %22, t29;\n\t"
        "madc.hi.cc.u32 t30, %13, %22, t30;\n\t"
        "madc.lo.cc.u32 t31, %15, %22, t31;\n\t"
        "madc.hi.u32 t32, %15, %22, 0x0;\n\t"
        "add.cc.u32 t47, t6, t24;\n\t"
        "mul.lo.u32 t48, t47, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t25, 0xf3b9cac2, t48, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xf3b9cac2, t48, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xbce6faad, t48, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xbce6faad, t48, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t48, t29;\n\t"
        "madc.hi.cc.u32 t30, 0xffffffff, t48, t30;\n\t"
        "madc.lo.cc.u32 t31, 0xffffffff, t48, t31;\n\t"
        "madc.hi.cc.u32 t32, 0xffffffff, t48, t32;\n\t"
        "addc.u32 t33, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t47, 0xfc632551, t48, t47;\n\t"
        "madc.hi.cc.u32 t7, 0xfc632551, t48, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xa7179e84, t48, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xa7179e84, t48, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t48, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xffffffff, t48, t11;\n\t"
        "madc.lo.cc.u32 t12, 0x0, t48, t12;\n\t"
        "madc.hi.cc.u32 t13, 0x0, t48, t13;\n\t"
        "addc.u32 t14, t14, 0x0;\n\t"
        "mad.lo.cc.u32 t25, %8, %23, t25;\n\t"
        "madc.hi.cc.u32 t26, %8, %23, t26;\n\t"
        "madc.lo.cc.u32 t27, %10, %23, t27;\n\t"
        "madc.hi.cc.u32 t28, %10, %23, t28;\n\t"
        "madc.lo.cc.u32 t29, %12, %23, t29;\n\t"
        "madc.hi.cc.u32 t30, %12, %23, t30;\n\t"
        "madc.lo.cc.u32 t31, %14, %23, t31;\n\t"
        "madc.hi.cc.u32 t32, %14, %23, t32;\n\t"
        "addc.u32 t33, t33, 0x0;\n\t"
        "mad.lo.cc.u32 t8, %9, %23, t8;\n\t"
        "madc.hi.cc.u32 t9, %9, %23, t9;\n\t"
        "madc.lo.cc.u32 t10, %11, %23, t10;\n\t"
        "madc.hi.cc.u32 t11, %11, %23, t11;\n\t"
        "madc.lo.cc.u32 t12, %13, %23, t12;\n\t"
        "madc.hi.cc.u32 t13, %13, %23, t13;\n\t"
        "madc.lo.cc.u32 t14, %15, %23, t14;\n\t"
        "madc.hi.u32 t15, %15, %23, 0x0;\n\t"
        "add.cc.u32 t49, t25, t7;\n\t"
        "mul.lo.u32 t50, t49, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t8, 0xf3b9cac2, t50, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xf3b9cac2, t50, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xbce6faad, t50, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xbce6faad, t50, t11;\n\t"
        "madc.lo.cc.u32 t12, 0xffffffff, t50, t12;\n\t"
        "madc.hi.cc.u32 t13, 0xffffffff, t50, t13;\n\t"
        "madc.lo.cc.u32 t14, 0xffffffff, t50, t14;\n\t"
        "madc.hi.cc.u32 t15, 0xffffffff, t50, t15;\n\t"
        "addc.u32 t16, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t49, 0xfc632551, t50, t49;\n\t"
        "madc.hi.cc.u32 t26, 0xfc632551, t50, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xa7179e84, t50, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xa7179e84, t50, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t50, t29;\n\t"
        "madc.hi.cc.u32 t30, 0xffffffff, t50, t30;\n\t"
        "madc.lo.cc.u32 t31, 0x0, t50, t31;\n\t"
        "madc.hi.cc.u32 t32, 0x0, t50, t32;\n\t"
        "addc.u32 t33, t33, 0x0;\n\t"
        "add.cc.u32 t51, t8, t26;\n\t"
        "addc.cc.u32 t52, t9, t27;\n\t"
        "addc.cc.u32 t53, t10, t28;\n\t"
        "addc.cc.u32 t54, t11, t29;\n\t"
        "addc.cc.u32 t55, t12, t30;\n\t"
        "addc.cc.u32 t56, t13, t31;\n\t"
        "addc.cc.u32 t57, t14, t32;\n\t"
        "addc.cc.u32 t58, t15, t33;\n\t"
        "addc.u32 t59, t16, 0x0;\n\t"
        "sub.cc.u32 t60, t51, 0xfc632551;\n\t"
        "subc.cc.u32 t61, t52, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t62, t53, 0xa7179e84;\n\t"
        "subc.cc.u32 t63, t54, 0xbce6faad;\n\t"
        "subc.cc.u32 t64, t55, 0xffffffff;\n\t"
        "subc.cc.u32 t65, t56, 0xffffffff;\n\t"
        "subc.cc.u32 t66, t57, 0x0;\n\t"
        "subc.cc.u32 t67, t58, 0xffffffff;\n\t"
        "subc.cc.u32 t68, t59, 0x0;\n\t"
        "subc.u32 t69, 0x0, 0x0;\n\t"
        "xor.b32 t70, t60, t51;\n\t"
        "and.b32 t71, t70, t69;\n\t"
        "xor.b32 t72, t71, t60;\n\t"
        "xor.b32 t73, t61, t52;\n\t"
        "and.b32 t74, t73, t69;\n\t"
        "xor.b32 t75, t74, t61;\n\t"
        "xor.b32 t76, t62, t53;\n\t"
        "and.b32 t77, t76, t69;\n\t"
        "xor.b32 t78, t77, t62;\n\t"
        "xor.b32 t79, t63, t54;\n\t"
        "and.b32 t80, t79, t69;\n\t"
        "xor.b32 t81, t80, t63;\n\t"
        "xor.b32 t82, t64, t55;\n\t"
        "and.b32 t83, t82, t69;\n\t"
        "xor.b32 t84, t83, t64;\n\t"
        "xor.b32 t85, t65, t56;\n\t"
        "and.b32 t86, t85, t69;\n\t"
        "xor.b32 t87, t86, t65;\n\t"
        "xor.b32 t88, t66, t57;\n\t"
        "and.b32 t89, t88, t69;\n\t"
        "xor.b32 t90, t89, t66;\n\t"
        "xor.b32 t91, t67, t58;\n\t"
        "and.b32 t92, t91, t69;\n\t"
        "xor.b32 t93, t92, t67;\n\t"
        "mov.u32 %0, t72;\n\t"
        "mov.u32 %1, t75;\n\t"
        "mov.u32 %2, t78;\n\t"
        "mov.u32 %3, t81;\n\t"
        "mov.u32 %4, t84;\n\t"
        "mov.u32 %5, t87;\n\t"
        "mov.u32 %6, t90;\n\t"
        "mov.u32 %7, t93;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t0 = (uint32_t)((uint32_t)(a_0_i * b_0_i));
    t1 = (uint32_t)(((uint64_t)a_0_i * b_0_i) >> 32);
    t2 = (uint32_t)((uint32_t)(a_2_i * b_0_i));
    t3 = (uint32_t)(((uint64_t)a_2_i * b_0_i) >> 32);
    t4 = (uint32_t)((uint32_t)(a_4_i * b_0_i));
    t5 = (uint32_t)(((uint64_t)a_4_i * b_0_i) >> 32);
    t6 = (uint32_t)((uint32_t)(a_6_i * b_0_i));
    t7 = (uint32_t)(((uint64_t)a_6_i * b_0_i) >> 32);
    t19 = (uint32_t)((uint32_t)(a_1_i * b_0_i));
    t20 = (uint32_t)(((uint64_t)a_1_i * b_0_i) >> 32);
    t21 = (uint32_t)((uint32_t)(a_3_i * b_0_i));
    t22 = (uint32_t)(((uint64_t)a_3_i * b_0_i) >> 32);
    t23 = (uint32_t)((uint32_t)(a_5_i * b_0_i));
    t24 = (uint32_t)(((uint64_t)a_5_i * b_0_i) >> 32);
    t25 = (uint32_t)((uint32_t)(a_7_i * b_0_i));
    t26 = (uint32_t)(((uint64_t)a_7_i * b_0_i) >> 32);
    t36 = (uint32_t)((uint32_t)(t0 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t36) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t36) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t36) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t36) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t36) + t0; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t36) >> 32) + t1 + cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t36) + t2 + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t36) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t36) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t36) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t8 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_1_i) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_1_i) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_1_i) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_1_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_1_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_1_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_1_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_1_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t27 + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_1_i) + t2; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_1_i) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_1_i) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_1_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_1_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_1_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_1_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_1_i) >> 32) + 0x0u + cf_; t9 = (uint32_t)w_;
    w_ = (uint64_t)t19 + t1; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t38 = (uint32_t)((uint32_t)(t37 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t38) + t2 + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t38) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t38) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t38) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t38) + t37; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t38) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t38) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t38) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t38) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t38) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t27 + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_2_i) + t2; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_2_i) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_2_i) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_2_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_2_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_2_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_2_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_2_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_2_i) + t21; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_2_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_2_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_2_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_2_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_2_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_2_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_2_i) >> 32) + 0x0u + cf_; t28 = (uint32_t)w_;
    w_ = (uint64_t)t2 + t20; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t40 = (uint32_t)((uint32_t)(t39 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t40) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t40) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t40) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t40) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t40) + t39; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t40) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t40) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t40) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t40) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t40) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_3_i) + t21; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_3_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_3_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_3_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_3_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_3_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_3_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_3_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t29 + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_3_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_3_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_3_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_3_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_3_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_3_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_3_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_3_i) >> 32) + 0x0u + cf_; t11 = (uint32_t)w_;
    w_ = (uint64_t)t21 + t3; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t42 = (uint32_t)((uint32_t)(t41 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t42) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t42) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t42) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t42) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t42) + t41; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t42) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t42) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t42) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t42) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t42) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t29 + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_4_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_4_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_4_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_4_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_4_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_4_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_4_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_4_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_4_i) + t23; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_4_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_4_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_4_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_4_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_4_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_4_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_4_i) >> 32) + 0x0u + cf_; t30 = (uint32_t)w_;
    w_ = (uint64_t)t4 + t22; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t44 = (uint32_t)((uint32_t)(t43 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t44) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t44) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t44) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t44) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t31 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t44) + t43; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t44) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t44) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t44) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t44) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t44) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_5_i) + t23; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_5_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_5_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_5_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_5_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_5_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_5_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_5_i) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t31 + 0x0u + cf_; t31 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_5_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_5_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_5_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_5_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_5_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_5_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_5_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_5_i) >> 32) + 0x0u + cf_; t13 = (uint32_t)w_;
    w_ = (uint64_t)t23 + t5; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t46 = (uint32_t)((uint32_t)(t45 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t46) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t46) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t46) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t46) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t14 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t46) + t45; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t46) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t46) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t46) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t46) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t46) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t31 + 0x0u + cf_; t31 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_6_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_6_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_6_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_6_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_6_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_6_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_6_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_6_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + 0x0u + cf_; t14 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_6_i) + t25; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_6_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_6_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_6_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_6_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_6_i) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_6_i) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_6_i) >> 32) + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)t6 + t24; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t48 = (uint32_t)((uint32_t)(t47 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t48) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t48) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t48) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t48) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t33 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t48) + t47; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t48) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t48) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t48) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t48) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t48) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + 0x0u + cf_; t14 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_7_i) + t25; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_7_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_7_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_7_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_7_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_7_i) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_7_i) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_7_i) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t33 + 0x0u + cf_; t33 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_7_i) + t8; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_7_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_7_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_7_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_7_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_7_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_7_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_7_i) >> 32) + 0x0u + cf_; t15 = (uint32_t)w_;
    w_ = (uint64_t)t25 + t7; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t50 = (uint32_t)((uint32_t)(t49 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t50) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t50) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t50) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t50) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t16 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t50) + t49; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t50) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t50) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t50) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t50) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t50) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t33 + 0x0u + cf_; t33 = (uint32_t)w_;
    w_ = (uint64_t)t8 + t26; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t27 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t28 + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t29 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t30 + cf_; t55 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t31 + cf_; t56 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + t32 + cf_; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + t33 + cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t16 + 0x0u + cf_; t59 = (uint32_t)w_;
    w_ = (uint64_t)t51 - 0xfc632551u; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t52 - 0xf3b9cac2u - cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t53 - 0xa7179e84u - cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t54 - 0xbce6faadu - cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t55 - 0xffffffffu - cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t56 - 0xffffffffu - cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t57 - 0x0u - cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t58 - 0xffffffffu - cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t59 - 0x0u - cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t69 = (uint32_t)w_;
    t70 = (uint32_t)(t60 ^ t51);
    t71 = (uint32_t)(t70 & t69);
    t72 = (uint32_t)(t71 ^ t60);
    t73 = (uint32_t)(t61 ^ t52);
    t74 = (uint32_t)(t73 & t69);
    t75 = (uint32_t)(t74 ^ t61);
    t76 = (uint32_t)(t62 ^ t53);
    t77 = (uint32_t)(t76 & t69);
    t78 = (uint32_t)(t77 ^ t62);
    t79 = (uint32_t)(t63 ^ t54);
    t80 = (uint32_t)(t79 & t69);
    t81 = (uint32_t)(t80 ^ t63);
    t82 = (uint32_t)(t64 ^ t55);
    t83 = (uint32_t)(t82 & t69);
    t84 = (uint32_t)(t83 ^ t64);
    t85 = (uint32_t)(t65 ^ t56);
    t86 = (uint32_t)(t85 & t69);
    t87 = (uint32_t)(t86 ^ t65);
    t88 = (uint32_t)(t66 ^ t57);
    t89 = (uint32_t)(t88 & t69);
    t90 = (uint32_t)(t89 ^ t66);
    t91 = (uint32_t)(t67 ^ t58);
    t92 = (uint32_t)(t91 & t69);
    t93 = (uint32_t)(t92 ^ t67);
    r[0] = t72;
    r[1] = t75;
    r[2] = t78;
    r[3] = t81;
    r[4] = t84;
    r[5] = t87;
    r[6] = t90;
    r[7] = t93;
#endif
  }

  // c = a*a (pseudo.py:663-702 / monty.py:982-1165)
  static MAB_DEV void sqr(uint32_t (&r)[8], const uint32_t (&a)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<94>;\n\t"
        "mul.lo.u32 t0, %8, %8;\n\t"
        "mul.hi.u32 t1, %8, %8;\n\t"
        "mul.lo.u32 t2, %10, %8;\n\t"
        "mul.hi.u32 t3, %10, %8;\n\t"
        "mul.lo.u32 t4, %12, %8;\n\t"
        "mul.hi.u32 t5, %12, %8;\n\t"
        "mul.lo.u32 t6, %14, %8;\n\t"
        "mul.hi.u32 t7, %14, %8;\n\t"
        "mul.lo.u32 t19, %9, %8;\n\t"
        "mul.hi.u32 t20, %9, %8;\n\t"
        "mul.lo.u32 t21, %11, %8;\n\t"
        "mul.hi.u32 t22, %11, %8;\n\t"
        "mul.lo.u32 t23, %13, %8;\n\t"
        "mul.hi.u32 t24, %13, %8;\n\t"
        "mul.lo.u32 t25, %15, %8;\n\t"
        "mul.hi.u32 t26, %15, %8;\n\t"
        "mul.lo.u32 t36, t0, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t19, 0xf3b9cac2, t36, t19;\n\t"
        "madc.hi.cc.u32 t20, 0xf3b9cac2, t36, t20;\n\t"
        "madc.lo.cc.u32 t21, 0xbce6faad, t36, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xbce6faad, t36, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xffffffff, t36, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xffffffff, t36, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t36, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t36, t26;\n\t"
        "addc.u32 t27, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t0, 0xfc632551, t36, t0;\n\t"
        "madc.hi.cc.u32 t1, 0xfc632551, t36, t1;\n\t"
        "madc.lo.cc.u32 t2, 0xa7179e84, t36, t2;\n\t"
        "madc.hi.cc.u32 t3, 0xa7179e84, t36, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xffffffff, t36, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xffffffff, t36, t5;\n\t"
        "madc.lo.cc.u32 t6, 0x0, t36, t6;\n\t"
        "madc.hi.cc.u32 t7, 0x0, t36, t7;\n\t"
        "addc.u32 t8, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t19, %8, %9, t19;\n\t"
        "madc.hi.cc.u32 t20, %8, %9, t20;\n\t"
        "madc.lo.cc.u32 t21, %10, %9, t21;\n\t"
        "madc.hi.cc.u32 t22, %10, %9, t22;\n\t"
        "madc.lo.cc.u32 t23, %12, %9, t23;\n\t"
        "madc.hi.cc.u32 t24, %12, %9, t24;\n\t"
        "madc.lo.cc.u32 t25, %14, %9, t25;\n\t"
        "madc.hi.cc.u32 t26, %14, %9, t26;\n\t"
        "addc.u32 t27, t27, 0x0;\n\t"
        "mad.lo.cc.u32 t2, %9, %9, t2;\n\t"
        "madc.hi.cc.u32 t3, %9, %9, t3;\n\t"
        "madc.lo.cc.u32 t4, %11, %9, t4;\n\t"
        "madc.hi.cc.u32 t5, %11, %9, t5;\n\t"
        "madc.lo.cc.u32 t6, %13, %9, t6;\n\t"
        "madc.hi.cc.u32 t7, %13, %9, t7;\n\t"
        "madc.lo.cc.u32 t8, %15, %9, t8;\n\t"
        "madc.hi.u32 t9, %15, %9, 0x0;\n\t"
        "add.cc.u32 t37, t19, t1;\n\t"
        "mul.lo.u32 t38, t37, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t2, 0xf3b9cac2, t38, t2;\n\t"
        "madc.hi.cc.u32 t3, 0xf3b9cac2, t38, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xbce6faad, t38, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xbce6faad, t38, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xffffffff, t38, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xffffffff, t38, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t38, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xffffffff, t38, t9;\n\t"
        "addc.u32 t10, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t37, 0xfc632551, t38, t37;\n\t"
        "madc.hi.cc.u32 t20, 0xfc632551, t38, t20;\n\t"
        "madc.lo.cc.u32 t21, 0xa7179e84, t38, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xa7179e84, t38, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xffffffff, t38, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xffffffff, t38, t24;\n\t"
        "madc.lo.cc.u32 t25, 0x0, t38, t25;\n\t"
        "madc.hi.cc.u32 t26, 0x0, t38, t26;\n\t"
        "addc.u32 t27, t27, 0x0;\n\t"
        "mad.lo.cc.u32 t2, %8, %10, t2;\n\t"
        "madc.hi.cc.u32 t3, %8, %10, t3;\n\t"
        "madc.lo.cc.u32 t4, %10, %10, t4;\n\t"
        "madc.hi.cc.u32 t5, %10, %10, t5;\n\t"
        "madc.lo.cc.u32 t6, %12, %10, t6;\n\t"
        "madc.hi.cc.u32 t7, %12, %10, t7;\n\t"
        "madc.lo.cc.u32 t8, %14, %10, t8;\n\t"
        "madc.hi.cc.u32 t9, %14, %10, t9;\n\t"
        "addc.u32 t10, t10, 0x0;\n\t"
        "mad.lo.cc.u32 t21, %9, %10, t21;\n\t"
        "madc.hi.cc.u32 t22, %9, %10, t22;\n\t"
        "madc.lo.cc.u32 t23, %11, %10, t23;\n\t"
        "madc.hi.cc.u32 t24, %11, %10, t24;\n\t"
        "madc.lo.cc.u32 t25, %13, %10, t25;\n\t"
        "madc.hi.cc.u32 t26, %13, %10, t26;\n\t"
        "madc.lo.cc.u32 t27, %15, %10, t27;\n\t"
        "madc.hi.u32 t28, %15, %10, 0x0;\n\t"
        "add.cc.u32 t39, t2, t20;\n\t"
        "mul.lo.u32 t40, t39, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t21, 0xf3b9cac2, t40, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xf3b9cac2, t40, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xbce6faad, t40, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xbce6faad, t40, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t40, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t40, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t40, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xffffffff, t40, t28;\n\t"
        "addc.u32 t29, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t39, 0xfc632551, t40, t39;\n\t"
        "madc.hi.cc.u32 t3, 0xfc632551, t40, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xa7179e84, t40, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xa7179e84, t40, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xffffffff, t40, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xffffffff, t40, t7;\n\t"
        "madc.lo.cc.u32 t8, 0x0, t40, t8;\n\t"
        "madc.hi.cc.u32 t9, 0x0, t40, t9;\n\t"
        "addc.u32 t10, t10, 0x0;\n\t"
        "mad.lo.cc.u32 t21, %8, %11, t21;\n\t"
        "madc.hi.cc.u32 t22, %8, %11, t22;\n\t"
        "madc.lo.cc.u32 t23, %10, %11, t23;\n\t"
        "madc.hi.cc.u32 t24, %10, %11, t24;\n\t"
        "madc.lo.cc.u32 t25, %12, %11, t25;\n\t"
        "madc.hi.cc.u32 t26, %12, %11, t26;\n\t"
        "madc.lo.cc.u32 t27, %14, %11, t27;\n\t"
        "madc.hi.cc.u32 t28, %14, %11, t28;\n\t"
        "addc.u32 t29, t29, 0x0;\n\t"
        "mad.lo.cc.u32 t4, %9, %11, t4;\n\t"
        "madc.hi.cc.u32 t5, %9, %11, t5;\n\t"
        "madc.lo.cc.u32 t6, %11, %11, t6;\n\t"
        "madc.hi.cc.u32 t7, %11, %11, t7;\n\t"
        "madc.lo.cc.u32 t8, %13, %11, t8;\n\t"
        "madc.hi.cc.u32 t9, %13, %11, t9;\n\t"
        "madc.lo.cc.u32 t10, %15, %11, t10;\n\t"
        "madc.hi.u32 t11, %15, %11, 0x0;\n\t"
        "add.cc.u32 t41, t21, t3;\n\t"
        "mul.lo.u32 t42, t41, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t4, 0xf3b9cac2, t42, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xf3b9cac2, t42, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xbce6faad, t42, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xbce6faad, t42, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t42, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xffffffff, t42, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t42, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xffffffff, t42, t11;\n\t"
        "addc.u32 t12, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t41, 0xfc632551, t42, t41;\n\t"
        "madc.hi.cc.u32 t22, 0xfc632551, t42, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xa7179e84, t42, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xa7179e84, t42, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t42, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t42, t26;\n\t"
        "madc.lo.cc.u32 t27, 0x0, t42, t27;\n\t"
        "madc.hi.cc.u32 t28, 0x0, t42, t28;\n\t"
        "addc.u32 t29, t29, 0x0;\n\t"
        "mad.lo.cc.u32 t4, %8, %12, t4;\n\t"
        "madc.hi.cc.u32 t5, %8, %12, t5;\n\t"
        "madc.lo.cc.u32 t6, %10, %12, t6;\n\t"
        "madc.hi.cc.u32 t7, %10, %12, t7;\n\t"
        "madc.lo.cc.u32 t8, %12, %12, t8;\n\t"
        "madc.hi.cc.u32 t9, %12, %12, t9;\n\t"
        "madc.lo.cc.u32 t10, %14, %12, t10;\n\t"
        "madc.hi.cc.u32 t11, %14, %12, t11;\n\t"
        "addc.u32 t12, t12, 0x0;\n\t"
        "mad.lo.cc.u32 t23, %9, %12, t23;\n\t"
        "madc.hi.cc.u32 t24, %9, %12, t24;\n\t"
        "madc.lo.cc.u32 t25, %11, %12, t25;\n\t"
        "madc.hi.cc.u32 t26, %11, %12, t26;\n\t"
        "madc.lo.cc.u32 t27, %13, %12, t27;\n\t"
        "madc.hi.cc.u32 t28, %13, %12, t28;\n\t"
        "madc.lo.cc.u32 t29, %15, %12, t29;\n\t"
        "madc.hi.u32 t30, %15, %12, 0x0;\n\t"
        "add.cc.u32 t43, t4, t22;\n\t"
        "mul.lo.u32 t44, t43, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t23, 0xf3b9cac2, t44, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xf3b9cac2, t44, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xbce6faad, t44, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xbce6faad, t44, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t44, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xffffffff, t44, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t44, t29;\n\t"
        "madc.hi.cc.u32 t30, 0xffffffff, t44, t30;\n\t"
        "addc.u32 t31, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t43, 0xfc632551, t44, t43;\n\t"
        "madc.hi.cc.u32 t5, 0xfc632551, t44, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xa7179e84, t44, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xa7179e84, t44, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t44, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xffffffff, t44, t9;\n\t"
        "madc.lo.cc.u32 t10, 0x0, t44, t10;\n\t"
        "madc.hi.cc.u32 t11, 0x0, t44, t11;\n\t"
        "addc.u32 t12, t12, 0x0;\n\t"
        "mad.lo.cc.u32 t23, %8, %13, t23;\n\t"
        "madc.hi.cc.u32 t24, %8, %13, t24;\n\t"
        "madc.lo.cc.u32 t25, %10, %13, t25;\n\t"
        "madc.hi.cc.u32 t26, %10, %13, t26;\n\t"
        "madc.lo.cc.u32 t27, %12, %13, t27;\n\t"
        "madc.hi.cc.u32 t28, %12, %13, t28;\n\t"
        "madc.lo.cc.u32 t29, %14, %13, t29;\n\t"
        "madc.hi.cc.u32 t30, %14, %13, t30;\n\t"
        "addc.u32 t31, t31, 0x0;\n\t"
        "mad.lo.cc.u32 t6, %9, %13, t6;\n\t"
        "madc.hi.cc.u32 t7, %9, %13, t7;\n\t"
        "madc.lo.cc.u32 t8, %11, %13, t8;\n\t"
        "madc.hi.cc.u32 t9, %11, %13, t9;\n\t"
        "madc.lo.cc.u32 t10, %13, %13, t10;\n\t"
        "madc.hi.cc.u32 t11, %13, %13, t11;\n\t"
        "madc.lo.cc.u32 t12, %15, %13, t12;\n\t"
        "madc.hi.u32 t13, %15, %13, 0x0;\n\t"
        "add.cc.u32 t45, t23, t5;\n\t"
        "mul.lo.u32 t46, t45, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t6, 0xf3b9cac2, t46, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xf3b9cac2, t46, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xbce6faad, t46, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xbce6faad, t46, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t46, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xffffffff, t46, t11;\n\t"
        "madc.lo.cc.u32 t12, 0xffffffff, t46, t12;\n\t"
        "madc.hi.cc.u32 t13, 0xffffffff, t46, t13;\n\t"
        "addc.u32 t14, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t45, 0xfc632551, t46, t45;\n\t"
        "madc.hi.cc.u32 t24, 0xfc632551, t46, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xa7179e84, t46, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xa7179e84, t46, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t46, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xffffffff, t46, t28;\n\t"
        "madc.lo.cc.u32 t29, 0x0, t46, t29;\n\t"
        "madc.hi.cc.u32 t30, 0x0, t46, t30;\n\t"
        "addc.u32 t31, t31, 0x0;\n\t"
        "mad.lo.cc.u32 t6, %8, %14, t6;\n\t"
        "madc.hi.cc.u32 t7, %8, %14, t7;\n\t"
        "madc.lo.cc.u32 t8, %10, %14, t8;\n\t"
        "madc.hi.cc.u32 t9, %10, %14, t9;\n\t"
        "madc.lo.cc.u32 t10, %12, %14, t10;\n\t"
        "madc.hi.cc.u32 t11, %12, %14, t11;\n\t"
        "madc.lo.cc.u32 t12, %14, %14, t12;\n\t"
        "madc.hi.cc.u32 t13, %14, %14, t13;\n\t"
        "addc.u32 t14, t14, 0x0;\n\t"
        "mad.lo.cc.u32 t25, %9, %14, t25;\n\t"
        "madc.hi.cc.u32 t26, %9, %14, t26;\n\t"
        "madc.lo.cc.u32 t27, %11, %14, t27;\n\t"
        "madc.hi.cc.u32 t28, %11, %14, t28;\n\t"
        "madc.lo.cc.u32 t29, %13, %14, t29;\n\t"
        "madc.hi.cc.u32 t30, %13, %14, t30;\n\t"
        "madc.lo.cc.u32 t31, %15, %14, t31;\n\t"
        "madc.hi.u32 t32, %15, %14, 0x0;\n\t"
        "add.cc.u32 t47, t6, t24;\n\t"
        "mul.lo.u32 t48, t47, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t25, 0xf3b9cac2, t48, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xf3b9cac2, t48, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xbce6faad, t48, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xbce6faad, t48, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t48, t29;\n\t"
        "madc.hi.cc.u32 t30, 0xffffffff, t48, t30;\n\t"
        "madc.lo.cc.u32 t31, 0xffffffff, t48, t31;\n\t"
        "madc.hi.cc.u32 t32, 0xffffffff, t48, t32;\n\t"
        "addc.u32 t33, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t47, 0xfc632551, t48, t47;\n\t"
        "madc.hi.cc.u32 t7, 0xfc632551, t48, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xa7179e84, t48, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xa7179e84, t48, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t48, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xffffffff, t48, t11;\n\t"
        "madc.lo.cc.u32 t12, 0x0, t48, t12;\n\t"
        "madc.hi.cc.u32 t13, 0x0, t48, t13;\n\t"
        "addc.u32 t14, t14, 0x0;\n\t"
        "mad.lo.cc.u32 t25, %8, %15, t25;\n\t"
        "madc.hi.cc.u32 t26, %8, %15, t26;\n\t"
        "madc.lo.cc.u32 t27, %10, %15, t27;\n\t"
        "madc.hi.cc.u32 t28, %10, %15, t28;\n\t"
        "madc.lo.cc.u32 t29, %12, %15, t29;\n\t"
        "madc.hi.cc.u32 t30, %12, %15, t30;\n\t"
        "madc.lo.cc.u32 t31, %14, %15, t31;\n\t"
        "madc.hi.cc.u32 t32, %14, %15, t32;\n\t"
        "addc.u32 t33, t33, 0x0;\n\t"
        "mad.lo.cc.u32 t8, %9, %15, t8;\n\t"
        "madc.hi.cc.u32 t9, %9, %15, t9;\n\t"
        "madc.lo.cc.u32 t10, %11, %15, t10;\n\t"
        "madc.hi.cc.u32 t11, %11, %15, t11;\n\t"
        "madc.lo.cc.u32 t12, %13, %15, t12;\n\t"
        "madc.hi.cc.u32 t13, %13, %15, t13;\n\t"
        "madc.lo.cc.u32 t14, %15, %15, t14;\n\t"
        "madc.hi.u32 t15, %15, %15, 0x0;\n\t"
        "add.cc.u32 t49, t25, t7;\n\t"
        "mul.lo.u32 t50, t49, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t8, 0xf3b9cac2, t50, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xf3b9cac2, t50, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xbce6faad, t50, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xbce6faad, t50, t11;\n\t"
        "madc.lo.cc.u32 t12, 0xffffffff, t50, t12;\n\t"
        "madc.hi.cc.u32 t13, 0xffffffff, t50, t13;\n\t"
        "madc.lo.cc.u32 t14, 0xffffffff, t50, t14;\n\t"
        "madc.hi.cc.u32 t15, 0xffffffff, t50, t15;\n\t"
        "addc.u32 t16, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t49, 0xfc632551, t50, t49;\n\t"
        "madc.hi.cc.u32 t26, 0xfc632551, t50, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xa7179e84, t50, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xa7179e84, t50, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t50, t29;\n\t"
        "madc.hi.cc.u32 t30, 0xffffffff, t50, t30;\n\t"
        "madc.lo.cc.u32 t31, 0x0, t50, t31;\n\t"
        "madc.hi.cc.u32 t32, 0x0, t50, t32;\n\t"
        "addc.u32 t33, t33, 0x0;\n\t"
        "add.cc.u32 t51, t8, t26;\n\t"
        "addc.cc.u32 t52, t9, t27;\n\t"
        "addc.cc.u32 t53, t10, t28;\n\t"
        "addc.cc.u32 t54, t11, t29;\n\t"
        "addc.cc.u32 t55, t12, t30;\n\t"
        "addc.cc.u32 t56, t13, t31;\n\t"
        "addc.cc.u32 t57, t14, t32;\n\t"
        "addc.cc.u32 t58, t15, t33;\n\t"
        "addc.u32 t59, t16, 0x0;\n\t"
        "sub.cc.u32 t60, t51, 0xfc632551;\n\t"
        "subc.cc.u32 t61, t52, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t62, t53, 0xa7179e84;\n\t"
        "subc.cc.u32 t63, t54, 0xbce6faad;\n\t"
        "subc.cc.u32 t64, t55, 0xffffffff;\n\t"
        "subc.cc.u32 t65, t56, 0xffffffff;\n\t"
        "subc.cc.u32 t66, t57, 0x0;\n\t"
        "subc.cc.u32 t67, t58, 0xffffffff;\n\t"
        "subc.cc.u32 t68, t59, 0x0;\n\t"
        "subc.u32 t69, 0x0, 0x0;\n\t"
        "xor.b32 t70, t60, t51;\n\t"
        "and.b32 t71, t70, t69;\n\t"
        "xor.b32 t72, t71, t60;\n\t"
        "xor.b32 t73, t61, t52;\n\t"
        "and.b32 t74, t73, t69;\n\t"
        "xor.b32 t75, t74, t61;\n\t"
        "xor.b32 t76, t62, t53;\n\t"
        "and.b32 t77, t76, t69;\n\t"
        "xor.b32 t78, t77, t62;\n\t"
        "xor.b32 t79, t63, t54;\n\t"
        "and.b32 t80, t79, t69;\n\t"
        "xor.b32 t81, t80, t63;\n\t"
        "xor.b32 t82, t64, t55;\n\t"
        "and.b32 t83, t82, t69;\n\t"
        "xor.b32 t84, t83, t64;\n\t"
        "xor.b32 t85, t65, t56;\n\t"
        "and.b32 t86, t85, t69;\n\t"
        "xor.b32 t87, t86, t65;\n\t"
        "xor.b32 t88, t66, t57;\n\t"
        "and.b32 t89, t88, t69;\n\t"
        "xor.b32 t90, t89, t66;\n\t"
        "xor.b32 t91, t67, t58;\n\t"
        "and.b32 t92, t91, t69;\n\t"
        "xor.b32 t93, t92, t67;\n\t"
        "mov.u32 %0, t72;\n\t"
        "mov.u32 %1, t75;\n\t"
        "mov.u32 %2, t78;\n\t"
        "mov.u32 %3, t81;\n\t"
        "mov.u32 %4, t84;\n\t"
        "mov.u32 %5, t87;\n\t"
        "mov.u32 %6, t90;\n\t"
        "mov.u32 %7, t93;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t0 = (uint32_t)((uint32_t)(a_0_i * a_0_i));
    t1 = (uint32_t)(((uint64_t)a_0_i * a_0_i) >> 32);
    t2 = (uint32_t)((uint32_t)(a_2_i * a_0_i));
    t3 = (uint32_t)(((uint64_t)a_2_i * a_0_i) >> 32);
    t4 = (uint32_t)((uint32_t)(a_4_i * a_0_i));
    t5 = (uint32_t)(((uint64_t)a_4_i * a_0_i) >> 32);
    t6 = (uint32_t)((uint32_t)(a_6_i * a_0_i));
    t7 = (uint32_t)(((uint64_t)a_6_i * a_0_i) >> 32);
    t19 = (uint32_t)((uint32_t)(a_1_i * a_0_i));
    t20 = (uint32_t)(((uint64_t)a_1_i * a_0_i) >> 32);
    t21 = (uint32_t)((uint32_t)(a_3_i * a_0_i));
    t22 = (uint32_t)(((uint64_t)a_3_i * a_0_i) >> 32);
    t23 = (uint32_t)((uint32_t)(a_5_i * a_0_i));
    t24 = (uint32_t)(((uint64_t)a_5_i * a_0_i) >> 32);
    t25 = (uint32_t)((uint32_t)(a_7_i * a_0_i));
    t26 = (uint32_t)(((uint64_t)a_7_i * a_0_i) >> 32);
    t36 = (uint32_t)((uint32_t)(t0 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t36) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t36) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t36) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t36) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t36) + t0; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t36) >> 32) + t1 + cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t36) + t2 + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t36) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t36) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t36) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t8 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * a_1_i) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * a_1_i) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_1_i) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_1_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_1_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_1_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_1_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_1_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t27 + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * a_1_i) + t2; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_1_i) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_1_i) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_1_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_1_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_1_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_1_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_1_i) >> 32) + 0x0u + cf_; t9 = (uint32_t)w_;
    w_ = (uint64_t)t19 + t1; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t38 = (uint32_t)((uint32_t)(t37 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t38) + t2 + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t38) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t38) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t38) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t38) + t37; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t38) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t38) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t38) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t38) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t38) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t27 + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * a_2_i) + t2; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * a_2_i) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_2_i) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_2_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_2_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_2_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_2_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_2_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * a_2_i) + t21; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_2_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_2_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_2_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_2_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_2_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_2_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_2_i) >> 32) + 0x0u + cf_; t28 = (uint32_t)w_;
    w_ = (uint64_t)t2 + t20; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t40 = (uint32_t)((uint32_t)(t39 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t40) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t40) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t40) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t40) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t40) + t39; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t40) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t40) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t40) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t40) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t40) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * a_3_i) + t21; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * a_3_i) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_3_i) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_3_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_3_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_3_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_3_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_3_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t29 + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * a_3_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_3_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_3_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_3_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_3_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_3_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_3_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_3_i) >> 32) + 0x0u + cf_; t11 = (uint32_t)w_;
    w_ = (uint64_t)t21 + t3; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t42 = (uint32_t)((uint32_t)(t41 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t42) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t42) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t42) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t42) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t42) + t41; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t42) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t42) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t42) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t42) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t42) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t29 + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * a_4_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * a_4_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_4_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_4_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_4_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_4_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_4_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_4_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * a_4_i) + t23; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_4_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_4_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_4_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_4_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_4_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_4_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_4_i) >> 32) + 0x0u + cf_; t30 = (uint32_t)w_;
    w_ = (uint64_t)t4 + t22; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t44 = (uint32_t)((uint32_t)(t43 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t44) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t44) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t44) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t44) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t31 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t44) + t43; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t44) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t44) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t44) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t44) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t44) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * a_5_i) + t23; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * a_5_i) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_5_i) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_5_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_5_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_5_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_5_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_5_i) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t31 + 0x0u + cf_; t31 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * a_5_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_5_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_5_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_5_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_5_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_5_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_5_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_5_i) >> 32) + 0x0u + cf_; t13 = (uint32_t)w_;
    w_ = (uint64_t)t23 + t5; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t46 = (uint32_t)((uint32_t)(t45 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t46) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t46) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t46) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t46) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t14 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t46) + t45; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t46) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t46) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t46) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t46) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t46) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t31 + 0x0u + cf_; t31 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * a_6_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * a_6_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_6_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_6_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_6_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_6_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_6_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_6_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + 0x0u + cf_; t14 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * a_6_i) + t25; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_6_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_6_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_6_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_6_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_6_i) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_6_i) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_6_i) >> 32) + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)t6 + t24; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t48 = (uint32_t)((uint32_t)(t47 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t48) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t48) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t48) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t48) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t33 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t48) + t47; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t48) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t48) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t48) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t48) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t48) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + 0x0u + cf_; t14 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * a_7_i) + t25; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * a_7_i) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_7_i) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_7_i) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_7_i) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_7_i) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_7_i) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_7_i) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t33 + 0x0u + cf_; t33 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * a_7_i) + t8; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_7_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_7_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_7_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_7_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_7_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_7_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_7_i) >> 32) + 0x0u + cf_; t15 = (uint32_t)w_;
    w_ = (uint64_t)t25 + t7; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t50 = (uint32_t)((uint32_t)(t49 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t50) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t50) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t50) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t50) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t16 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t50) + t49; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t50) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t50) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t50) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t50) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t50) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t33 + 0x0u + cf_; t33 = (uint32_t)w_;
    w_ = (uint64_t)t8 + t26; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t27 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t28 + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t29 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t30 + cf_; t55 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t31 + cf_; t56 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + t32 + cf_; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + t33 + cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t16 + 0x0u + cf_; t59 = (uint32_t)w_;
    w_ = (uint64_t)t51 - 0xfc632551u; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t52 - 0xf3b9cac2u - cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t53 - 0xa7179e84u - cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t54 - 0xbce6faadu - cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t55 - 0xffffffffu - cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t56 - 0xffffffffu - cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t57 - 0x0u - cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t58 - 0xffffffffu - cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t59 - 0x0u - cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t69 = (uint32_t)w_;
    t70 = (uint32_t)(t60 ^ t51);
    t71 = (uint32_t)(t70 & t69);
    t72 = (uint32_t)(t71 ^ t60);
    t73 = (uint32_t)(t61 ^ t52);
    t74 = (uint32_t)(t73 & t69);
    t75 = (uint32_t)(t74 ^ t61);
    t76 = (uint32_t)(t62 ^ t53);
    t77 = (uint32_t)(t76 & t69);
    t78 = (uint32_t)(t77 ^ t62);
    t79 = (uint32_t)(t63 ^ t54);
    t80 = (uint32_t)(t79 & t69);
    t81 = (uint32_t)(t80 ^ t63);
    t82 = (uint32_t)(t64 ^ t55);
    t83 = (uint32_t)(t82 & t69);
    t84 = (uint32_t)(t83 ^ t64);
    t85 = (uint32_t)(t65 ^ t56);
    t86 = (uint32_t)(t85 & t69);
    t87 = (uint32_t)(t86 ^ t65);
    t88 = (uint32_t)(t66 ^ t57);
    t89 = (uint32_t)(t88 & t69);
    t90 = (uint32_t)(t89 ^ t66);
    t91 = (uint32_t)(t67 ^ t58);
    t92 = (uint32_t)(t91 & t69);
    t93 = (uint32_t)(t92 ^ t67);
    r[0] = t72;
    r[1] = t75;
    r[2] = t78;
    r[3] = t81;
    r[4] = t84;
    r[5] = t87;
    r[6] = t90;
    r[7] = t93;
#endif
  }

  // c = a*b for a small integer b (pseudo.py:705-728 / monty.py:876-978)
  static MAB_DEV void mli(uint32_t (&r)[8], const uint32_t (&a)[8], uint32_t b) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<188>;\n\t"
        "mul.lo.u32 t0, 0xbe79eea2, %16;\n\t"
        "mul.hi.u32 t1, 0xbe79eea2, %16;\n\t"
        "mul.lo.u32 t2, 0x49bd6fa6, %16;\n\t"
        "mul.hi.u32 t3, 0x49bd6fa6, %16;\n\t"
        "mul.lo.u32 t4, 0x2b6bec59, %16;\n\t"
        "mul.hi.u32 t5, 0x2b6bec59, %16;\n\t"
        "mul.lo.u32 t6, 0xf3d95620, %16;\n\t"
        "mul.hi.u32 t7, 0xf3d95620, %16;\n\t"
        "mul.lo.u32 t19, 0x83244c95, %16;\n\t"
        "mul.hi.u32 t20, 0x83244c95, %16;\n\t"
        "mul.lo.u32 t21, 0x4699799c, %16;\n\t"
        "mul.hi.u32 t22, 0x4699799c, %16;\n\t"
        "mul.lo.u32 t23, 0x2845b239, %16;\n\t"
        "mul.hi.u32 t24, 0x2845b239, %16;\n\t"
        "mul.lo.u32 t25, 0x66e12d94, %16;\n\t"
        "mul.hi.u32 t26, 0x66e12d94, %16;\n\t"
        "mul.lo.u32 t36, t0, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t19, 0xf3b9cac2, t36, t19;\n\t"
        "madc.hi.cc.u32 t20, 0xf3b9cac2, t36, t20;\n\t"
        "madc.lo.cc.u32 t21, 0xbce6faad, t36, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xbce6faad, t36, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xffffffff, t36, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xffffffff, t36, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t36, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t36, t26;\n\t"
        "addc.u32 t27, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t0, 0xfc632551, t36, t0;\n\t"
        "madc.hi.cc.u32 t1, 0xfc632551, t36, t1;\n\t"
        "madc.lo.cc.u32 t2, 0xa7179e84, t36, t2;\n\t"
        "madc.hi.cc.u32 t3, 0xa7179e84, t36, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xffffffff, t36, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xffffffff, t36, t5;\n\t"
        "madc.lo.cc.u32 t6, 0x0, t36, t6;\n\t"
        "madc.hi.cc.u32 t7, 0x0, t36, t7;\n\t"
        "addc.u32 t8, 0x0, 0x0;\n\t"
        "add.cc.u32 t37, t19, t1;\n\t"
        "mul.lo.u32 t38, t37, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t2, 0xf3b9cac2, t38, t2;\n\t"
        "madc.hi.cc.u32 t3, 0xf3b9cac2, t38, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xbce6faad, t38, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xbce6faad, t38, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xffffffff, t38, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xffffffff, t38, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t38, t8;\n\t"
        "madc.hi.u32 t9, 0xffffffff, t38, 0x0;\n\t"
        "mad.lo.cc.u32 t37, 0xfc632551, t38, t37;\n\t"
        "madc.hi.cc.u32 t20, 0xfc632551, t38, t20;\n\t"
        "madc.lo.cc.u32 t21, 0xa7179e84, t38, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xa7179e84, t38, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xffffffff, t38, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xffffffff, t38, t24;\n\t"
        "madc.lo.cc.u32 t25, 0x0, t38, t25;\n\t"
        "madc.hi.cc.u32 t26, 0x0, t38, t26;\n\t"
        "addc.u32 t27, t27, 0x0;\n\t"
        "add.cc.u32 t39, t2, t20;\n\t"
        "mul.lo.u32 t40, t39, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t21, 0xf3b9cac2, t40, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xf3b9cac2, t40, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xbce6faad, t40, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xbce6faad, t40, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t40, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t40, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t40, t27;\n\t"
        "madc.hi.u32 t28, 0xffffffff, t40, 0x0;\n\t"
        "mad.lo.cc.u32 t39, 0xfc632551, t40, t39;\n\t"
        "madc.hi.cc.u32 t3, 0xfc632551, t40, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xa7179e84, t40, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xa7179e84, t40, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xffffffff, t40, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xffffffff, t40, t7;\n\t"
        "madc.lo.cc.u32 t8, 0x0, t40, t8;\n\t"
        "madc.hi.cc.u32 t9, 0x0, t40, t9;\n\t"
        "addc.u32 t10, 0x0, 0x0;\n\t"
        "add.cc.u32 t41, t21, t3;\n\t"
        "mul.lo.u32 t42, t41, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t4, 0xf3b9cac2, t42, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xf3b9cac2, t42, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xbce6faad, t42, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xbce6faad, t42, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t42, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xffffffff, t42, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t42, t10;\n\t"
        "madc.hi.u32 t11, 0xffffffff, t42, 0x0;\n\t"
        "mad.lo.cc.u32 t41, 0xfc632551, t42, t41;\n\t"
        "madc.hi.cc.u32 t22, 0xfc632551, t42, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xa7179e84, t42, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xa7179e84, t42, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t42, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t42, t26;\n\t"
        "madc.lo.cc.u32 t27, 0x0, t42, t27;\n\t"
        "madc.hi.cc.u32 t28, 0x0, t42, t28;\n\t"
        "addc.u32 t29, 0x0, 0x0;\n\t"
        "add.cc.u32 t43, t4, t22;\n\t"
        "mul.lo.u32 t44, t43, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t23, 0xf3b9cac2, t44, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xf3b9cac2, t44, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xbce6faad, t44, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xbce6faad, t44, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t44, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xffffffff, t44, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t44, t29;\n\t"
        "madc.hi.u32 t30, 0xffffffff, t44, 0x0;\n\t"
        "mad.lo.cc.u32 t43, 0xfc632551, t44, t43;\n\t"
        "madc.hi.cc.u32 t5, 0xfc632551, t44, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xa7179e84, t44, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xa7179e84, t44, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t44, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xffffffff, t44, t9;\n\t"
        "madc.lo.cc.u32 t10, 0x0, t44, t10;\n\t"
        "madc.hi.cc.u32 t11, 0x0, t44, t11;\n\t"
        "addc.u32 t12, 0x0, 0x0;\n\t"
        "add.cc.u32 t45, t23, t5;\n\t"
        "mul.lo.u32 t46, t45, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t6, 0xf3b9cac2, t46, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xf3b9cac2, t46, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xbce6faad, t46, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xbce6faad, t46, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t46, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xffffffff, t46, t11;\n\t"
        "madc.lo.cc.u32 t12, 0xffffffff, t46, t12;\n\t"
        "madc.hi.u32 t13, 0xffffffff, t46, 0x0;\n\t"
        "mad.lo.cc.u32 t45, 0xfc632551, t46, t45;\n\t"
        "madc.hi.cc.u32 t24, 0xfc632551, t46, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xa7179e84, t46, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xa7179e84, t46, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t46, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xffffffff, t46, t28;\n\t"
        "madc.lo.cc.u32 t29, 0x0, t46, t29;\n\t"
        "madc.hi.cc.u32 t30, 0x0, t46, t30;\n\t"
        "addc.u32 t31, 0x0, 0x0;\n\t"
        "add.cc.u32 t47, t6, t24;\n\t"
        "mul.lo.u32 t48, t47, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t25, 0xf3b9cac2, t48, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xf3b9cac2, t48, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xbce6faad, t48, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xbce6faad, t48, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t48, t29;\n\t"
        "madc.hi.cc.u32 t30, 0xffffffff, t48, t30;\n\t"
        "madc.lo.cc.u32 t31, 0xffffffff, t48, t31;\n\t"
        "madc.hi.u32 t32, 0xffffffff, t48, 0x0;\n\t"
        "mad.lo.cc.u32 t47, 0xfc632551, t48, t47;\n\t"
        "madc.hi.cc.u32 t7, 0xfc632551, t48, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xa7179e84, t48, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xa7179e84, t48, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t48, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xffffffff, t48, t11;\n\t"
        "madc.lo.cc.u32 t12, 0x0, t48, t12;\n\t"
        "madc.hi.cc.u32 t13, 0x0, t48, t13;\n\t"
        "addc.u32 t14, 0x0, 0x0;\n\t"
        "add.cc.u32 t49, t25, t7;\n\t"
        "mul.lo.u32 t50, t49, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t8, 0xf3b9cac2, t50, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xf3b9cac2, t50, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xbce6faad, t50, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xbce6faad, t50, t11;\n\t"
        "madc.lo.cc.u32 t12, 0xffffffff, t50, t12;\n\t"
        "madc.hi.cc.u32 t13, 0xffffffff, t50, t13;\n\t"
        "madc.lo.cc.u32 t14, 0xffffffff, t50, t14;\n\t"
        "madc.hi.u32 t15, 0xffffffff, t50, 0x0;\n\t"
        "mad.lo.cc.u32 t49, 0xfc632551, t50, t49;\n\t"
        "madc.hi.cc.u32 t26, 0xfc632551, t50, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xa7179e84, t50, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xa7179e84, t50, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t50, t29;\n\t"
        "madc.hi.cc.u32 t30, 0xffffffff, t50, t30;\n\t"
        "madc.lo.cc.u32 t31, 0x0, t50, t31;\n\t"
        "madc.hi.cc.u32 t32, 0x0, t50, t32;\n\t"
        "addc.u32 t33, 0x0, 0x0;\n\t"
        "add.cc.u32 t51, t8, t26;\n\t"
        "addc.cc.u32 t52, t9, t27;\n\t"
        "addc.cc.u32 t53, t10, t28;\n\t"
        "addc.cc.u32 t54, t11, t29;\n\t"
        "addc.cc.u32 t55, t12, t30;\n\t"
        "addc.cc.u32 t56, t13, t31;\n\t"
        "addc.cc.u32 t57, t14, t32;\n\t"
        "addc.cc.u32 t58, t15, t33;\n\t"
        "addc.u32 t59, 0x0, 0x0;\n\t"
        "sub.cc.u32 t60, t51, 0xfc632551;\n\t"
        "subc.cc.u32 t61, t52, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t62, t53, 0xa7179e84;\n\t"
        "subc.cc.u32 t63, t54, 0xbce6faad;\n\t"
        "subc.cc.u32 t64, t55, 0xffffffff;\n\t"
        "subc.cc.u32 t65, t56, 0xffffffff;\n\t"
        "subc.cc.u32 t66, t57, 0x0;\n\t"
        "subc.cc.u32 t67, t58, 0xffffffff;\n\t"
        "subc.cc.u32 t68, t59, 0x0;\n\t"
        "subc.u32 t69, 0x0, 0x0;\n\t"
        "xor.b32 t70, t60, t51;\n\t"
        "and.b32 t71, t70, t69;\n\t"
        "xor.b32 t72, t71, t60;\n\t"
        "xor.b32 t73, t61, t52;\n\t"
        "and.b32 t74, t73, t69;\n\t"
        "xor.b32 t75, t74, t61;\n\t"
        "xor.b32 t76, t62, t53;\n\t"
        "and.b32 t77, t76, t69;\n\t"
        "xor.b32 t78, t77, t62;\n\t"
        "xor.b32 t79, t63, t54;\n\t"
        "and.b32 t80, t79, t69;\n\t"
        "xor.b32 t81, t80, t63;\n\t"
        "xor.b32 t82, t64, t55;\n\t"
        "and.b32 t83, t82, t69;\n\t"
        "xor.b32 t84, t83, t64;\n\t"
        "xor.b32 t85, t65, t56;\n\t"
        "and.b32 t86, t85, t69;\n\t"
        "xor.b32 t87, t86, t65;\n\t"
        "xor.b32 t88, t66, t57;\n\t"
        "and.b32 t89, t88, t69;\n\t"
        "xor.b32 t90, t89, t66;\n\t"
        "xor.b32 t91, t67, t58;\n\t"
        "and.b32 t92, t91, t69;\n\t"
        "xor.b32 t93, t92, t67;\n\t"
        "mul.lo.u32 t94, %8, t72;\n\t"
        "mul.hi.u32 t95, %8, t72;\n\t"
        "mul.lo.u32 t96, %10, t72;\n\t"
        "mul.hi.u32 t97, %10, t72;\n\t"
        "mul.lo.u32 t98, %12, t72;\n\t"
        "mul.hi.u32 t99, %12, t72;\n\t"
        "mul.lo.u32 t100, %14, t72;\n\t"
        "mul.hi.u32 t101, %14, t72;\n\t"
        "mul.lo.u32 t113, %9, t72;\n\t"
        "mul.hi.u32 t114, %9, t72;\n\t"
        "mul.lo.u32 t115, %11, t72;\n\t"
        "mul.hi.u32 t116, %11, t72;\n\t"
        "mul.lo.u32 t117, %13, t72;\n\t"
        "mul.hi.u32 t118, %13, t72;\n\t"
        "mul.lo.u32 t119, %15, t72;\n\t"
        "mul.hi.u32 t120, %15, t72;\n\t"
        "mul.lo.u32 t130, t94, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t113, 0xf3b9cac2, t130, t113;\n\t"
        "madc.hi.cc.u32 t114, 0xf3b9cac2, t130, t114;\n\t"
        "madc.lo.cc.u32 t115, 0xbce6faad, t130, t115;\n\t"
        "madc.hi.cc.u32 t116, 0xbce6faad, t130, t116;\n\t"
        "madc.lo.cc.u32 t117, 0xffffffff, t130, t117;\n\t"
        "madc.hi.cc.u32 t118, 0xffffffff, t130, t118;\n\t"
        "madc.lo.cc.u32 t119, 0xffffffff, t130, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xffffffff, t130, t120;\n\t"
        "addc.u32 t121, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t94, 0xfc632551, t130, t94;\n\t"
        "madc.hi.cc.u32 t95, 0xfc632551, t130, t95;\n\t"
        "madc.lo.cc.u32 t96, 0xa7179e84, t130, t96;\n\t"
        "madc.hi.cc.u32 t97, 0xa7179e84, t130, t97;\n\t"
        "madc.lo.cc.u32 t98, 0xffffffff, t130, t98;\n\t"
        "madc.hi.cc.u32 t99, 0xffffffff, t130, t99;\n\t"
        "madc.lo.cc.u32 t100, 0x0, t130, t100;\n\t"
        "madc.hi.cc.u32 t101, 0x0, t130, t101;\n\t"
        "addc.u32 t102, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t113, %8, t75, t113;\n\t"
        "madc.hi.cc.u32 t114, %8, t75, t114;\n\t"
        "madc.lo.cc.u32 t115, %10, t75, t115;\n\t"
        "madc.hi.cc.u32 t116, %10, t75, t116;\n\t"
        "madc.lo.cc.u32 t117, %12, t75, t117;\n\t"
        "madc.hi.cc.u32 t118, %12, t75, t118;\n\t"
        "madc.lo.cc.u32 t119, %14, t75, t119;\n\t"
        "madc.hi.cc.u32 t120, %14, t75, t120;\n\t"
        "addc.u32 t121, t121, 0x0;\n\t"
        "mad.lo.cc.u32 t96, %9, t75, t96;\n\t"
        "madc.hi.cc.u32 t97, %9, t75, t97;\n\t"
        "madc.lo.cc.u32 t98, %11, t75, t98;\n\t"
        "madc.hi.cc.u32 t99, %11, t75, t99;\n\t"
        "madc.lo.cc.u32 t100, %13, t75, t100;\n\t"
        "madc.hi.cc.u32 t101, %13, t75, t101;\n\t"
        "madc.lo.cc.u32 t102, %15, t75, t102;\n\t"
        "madc.hi.u32 t103, %15, t75, 0x0;\n\t"
        "add.cc.u32 t131, t113, t95;\n\t"
        "mul.lo.u32 t132, t131, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t96, 0xf3b9cac2, t132, t96;\n\t"
        "madc.hi.cc.u32 t97, 0xf3b9cac2, t132, t97;\n\t"
        "madc.lo.cc.u32 t98, 0xbce6faad, t132, t98;\n\t"
        "madc.hi.cc.u32 t99, 0xbce6faad, t132, t99;\n\t"
        "madc.lo.cc.u32 t100, 0xffffffff, t132, t100;\n\t"
        "madc.hi.cc.u32 t101, 0xffffffff, t132, t101;\n\t"
        "madc.lo.cc.u32 t102, 0xffffffff, t132, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xffffffff, t132, t103;\n\t"
        "addc.u32 t104, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t131, 0xfc632551, t132, t131;\n\t"
        "madc.hi.cc.u32 t114, 0xfc632551, t132, t114;\n\t"
        "madc.lo.cc.u32 t115, 0xa7179e84, t132, t115;\n\t"
        "madc.hi.cc.u32 t116, 0xa7179e84, t132, t116;\n\t"
        "madc.lo.cc.u32 t117, 0xffffffff, t132, t117;\n\t"
        "madc.hi.cc.u32 t118, 0xffffffff, t132, t118;\n\t"
        "madc.lo.cc.u32 t119, 0x0, t132, t119;\n\t"
        "madc.hi.cc.u32 t120, 0x0, t132, t120;\n\t"
        "addc.u32 t121, t121, 0x0;\n\t"
        "mad.lo.cc.u32 t96, %8, t78, t96;\n\t"
        "madc.hi.cc.u32 t97, %8, t78, t97;\n\t"
        "madc.lo.cc.u32 t98, %10, t78, t98;\n\t"
        "madc.hi.cc.u32 t99, %10, t78, t99;\n\t"
        "madc.lo.cc.u32 t100, %12, t78, t100;\n\t"
        "madc.hi.cc.u32 t101, %12, t78, t101;\n\t"
        "madc.lo.cc.u32 t102, %14, t78, t102;\n\t"
        "madc.hi.cc.u32 t103, %14, t78, t103;\n\t"
        "addc.u32 t104, t104, 0x0;\n\t"
        "mad.lo.cc.u32 t115, %9, t78, t115;\n\t"
        "madc.hi.cc.u32 t116, %9, t78, t116;\n\t"
        "madc.lo.cc.u32 t117, %11, t78, t117;\n\t"
        "madc.hi.cc.u32 t118, %11, t78, t118;\n\t"
        "madc.lo.cc.u32 t119, %13, t78, t119;\n\t"
        "madc.hi.cc.u32 t120, %13, t78, t120;\n\t"
        "madc.lo.cc.u32 t121, %15, t78, t121;\n\t"
        "madc.hi.u32 t122, %15, t78, 0x0;\n\t"
        "add.cc.u32 t133, t96, t114;\n\t"
        "mul.lo.u32 t134, t133, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t115, 0xf3b9cac2, t134, t115;\n\t"
        "madc.hi.cc.u32 t116, 0xf3b9cac2, t134, t116;\n\t"
        "madc.lo.cc.u32 t117, 0xbce6faad, t134, t117;\n\t"
        "madc.hi.cc.u32 t118, 0xbce6faad, t134, t118;\n\t"
        "madc.lo.cc.u32 t119, 0xffffffff, t134, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xffffffff, t134, t120;\n\t"
        "madc.lo.cc.u32 t121, 0xffffffff, t134, t121;\n\t"
        "madc.hi.cc.u32 t122, 0xffffffff, t134, t122;\n\t"
        "addc.u32 t123, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t133, 0xfc632551, t134, t133;\n\t"
        "madc.hi.cc.u32 t97, 0xfc632551, t134, t97;\n\t"
        "madc.lo.cc.u32 t98, 0xa7179e84, t134, t98;\n\t"
        "madc.hi.cc.u32 t99, 0xa7179e84, t134, t99;\n\t"
        "madc.lo.cc.u32 t100, 0xffffffff, t134, t100;\n\t"
        "madc.hi.cc.u32 t101, 0xffffffff, t134, t101;\n\t"
        "madc.lo.cc.u32 t102, 0x0, t134, t102;\n\t"
        "madc.hi.cc.u32 t103, 0x0, t134, t103;\n\t"
        "addc.u32 t104, t104, 0x0;\n\t"
        "mad.lo.cc.u32 t115, %8, t81, t115;\n\t"
        "madc.hi.cc.u32 t116, %8, t81, t116;\n\t"
        "madc.lo.cc.u32 t117, %10, t81, t117;\n\t"
        "madc.hi.cc.u32 t118, %10, t81, t118;\n\t"
        "madc.lo.cc.u32 t119, %12, t81, t119;\n\t"
        "madc.hi.cc.u32 t120, %12, t81, t120;\n\t"
        "madc.lo.cc.u32 t121, %14, t81, t121;\n\t"
        "madc.hi.cc.u32 t122, %14, t81, t122;\n\t"
        "addc.u32 t123, t123, 0x0;\n\t"
        "mad.lo.cc.u32 t98, %9, t81, t98;\n\t"
        "madc.hi.cc.u32 t99, %9, t81, t99;\n\t"
        "madc.lo.cc.u32 t100, %11, t81, t100;\n\t"
        "madc.hi.cc.u32 t101, %11, t81, t101;\n\t"
        "madc.lo.cc.u32 t102, %13, t81, t102;\n\t"
        "madc.hi.cc.u32 t103, %13, t81, t103;\n\t"
        "madc.lo.cc.u32 t104, %15, t81, t104;\n\t"
        "madc.hi.u32 t105, %15, t81, 0x0;\n\t"
        "add.cc.u32 t135, t115, t97;\n\t"
        "mul.lo.u32 t136, t135, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t98, 0xf3b9cac2, t136, t98;\n\t"
        "madc.hi.cc.u32 t99, 0xf3b9cac2, t136, t99;\n\t"
        "madc.lo.cc.u32 t100, 0xbce6faad, t136, t100;\n\t"
        "madc.hi.cc.u32 t101, 0xbce6faad, t136, t101;\n\t"
        "madc.lo.cc.u32 t102, 0xffffffff, t136, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xffffffff, t136, t103;\n\t"
        "madc.lo.cc.u32 t104, 0xffffffff, t136, t104;\n\t"
        "madc.hi.cc.u32 t105, 0xffffffff, t136, t105;\n\t"
        "addc.u32 t106, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t135, 0xfc632551, t136, t135;\n\t"
        "madc.hi.cc.u32 t116, 0xfc632551, t136, t116;\n\t"
        "madc.lo.cc.u32 t117, 0xa7179e84, t136, t117;\n\t"
        "madc.hi.cc.u32 t118, 0xa7179e84, t136, t118;\n\t"
        "madc.lo.cc.u32 t119, 0xffffffff, t136, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xffffffff, t136, t120;\n\t"
        "madc.lo.cc.u32 t121, 0x0, t136, t121;\n\t"
        "madc.hi.cc.u32 t122, 0x0, t136, t122;\n\t"
        "addc.u32 t123, t123, 0x0;\n\t"
        "mad.lo.cc.u32 t98, %8, t84, t98;\n\t"
        "madc.hi.cc.u32 t99, %8, t84, t99;\n\t"
        "madc.lo.cc.u32 t100, %10, t84, t100;\n\t"
        "madc.hi.cc.u32 t101, %10, t84, t101;\n\t"
        "madc.lo.cc.u32 t102, %12, t84, t102;\n\t"
        "madc.hi.cc.u32 t103, %12, t84, t103;\n\t"
        "madc.lo.cc.u32 t104, %14, t84, t104;\n\t"
        "madc.hi.cc.u32 t105, %14, t84, t105;\n\t"
        "addc.u32 t106, t106, 0x0;\n\t"
        "mad.lo.cc.u32 t117, %9, t84, t117;\n\t"
        "madc.hi.cc.u32 t118, %9, t84, t118;\n\t"
        "madc.lo.cc.u32 t119, %11, t84, t119;\n\t"
        "madc.hi.cc.u32 t120, %11, t84, t120;\n\t"
        "madc.lo.cc.u32 t121, %13, t84, t121;\n\t"
        "madc.hi.cc.u32 t122, %13, t84, t122;\n\t"
        "madc.lo.cc.u32 t123, %15, t84, t123;\n\t"
        "madc.hi.u32 t124, %15, t84, 0x0;\n\t"
        "add.cc.u32 t137, t98, t116;\n\t"
        "mul.lo.u32 t138, t137, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t117, 0xf3b9cac2, t138, t117;\n\t"
        "madc.hi.cc.u32 t118, 0xf3b9cac2, t138, t118;\n\t"
        "madc.lo.cc.u32 t119, 0xbce6faad, t138, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xbce6faad, t138, t120;\n\t"
        "madc.lo.cc.u32 t121, 0xffffffff, t138, t121;\n\t"
        "madc.hi.cc.u32 t122, 0xffffffff, t138, t122;\n\t"
        "madc.lo.cc.u32 t123, 0xffffffff, t138, t123;\n\t"
        "madc.hi.cc.u32 t124, 0xffffffff, t138, t124;\n\t"
        "addc.u32 t125, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t137, 0xfc632551, t138, t137;\n\t"
        "madc.hi.cc.u32 t99, 0xfc632551, t138, t99;\n\t"
        "madc.lo.cc.u32 t100, 0xa7179e84, t138, t100;\n\t"
        "madc.hi.cc.u32 t101, 0xa7179e84, t138, t101;\n\t"
        "madc.lo.cc.u32 t102, 0xffffffff, t138, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xffffffff, t138, t103;\n\t"
        "madc.lo.cc.u32 t104, 0x0, t138, t104;\n\t"
        "madc.hi.cc.u32 t105, 0x0, t138, t105;\n\t"
        "addc.u32 t106, t106, 0x0;\n\t"
        "mad.lo.cc.u32 t117, %8, t87, t117;\n\t"
        "madc.hi.cc.u32 t118, %8, t87, t118;\n\t"
        "madc.lo.cc.u32 t119, %10, t87, t119;\n\t"
        "madc.hi.cc.u32 t120, %10, t87, t120;\n\t"
        "madc.lo.cc.u32 t121, %12, t87, t121;\n\t"
        "madc.hi.cc.u32 t122, %12, t87, t122;\n\t"
        "madc.lo.cc.u32 t123, %14, t87, t123;\n\t"
        "madc.hi.cc.u32 t124, %14, t87, t124;\n\t"
        "addc.u32 t125, t125, 0x0;\n\t"
        "mad.lo.cc.u32 t100, %9, t87, t100;\n\t"
        "madc.hi.cc.u32 t101, %9, t87, t101;\n\t"
        "madc.lo.cc.u32 t102, %11, t87, t102;\n\t"
        "madc.hi.cc.u32 t103, %11, t87, t103;\n\t"
        "madc.lo.cc.u32 t104, %13, t87, t104;\n\t"
        "madc.hi.cc.u32 t105, %13, t87, t105;\n\t"
        "madc.lo.cc.u32 t106, %15, t87, t106;\n\t"
        "madc.hi.u32 t107, %15, t87, 0x0;\n\t"
        "add.cc.u32 t139, t117, t99;\n\t"
        "mul.lo.u32 t140, t139, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t100, 0xf3b9cac2, t140, t100;\n\t"
        "madc.hi.cc.u32 t101, 0xf3b9cac2, t140, t101;\n\t"
        "madc.lo.cc.u32 t102, 0xbce6faad, t140, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xbce6faad, t140, t103;\n\t"
        "madc.lo.cc.u32 t104, 0xffffffff, t140, t104;\n\t"
        "madc.hi.cc.u32 t105, 0xffffffff, t140, t105;\n\t"
        "madc.lo.cc.u32 t106, 0xffffffff, t140, t106;\n\t"
        "madc.hi.cc.u32 t107, 0xffffffff, t140, t107;\n\t"
        "addc.u32 t108, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t139, 0xfc632551, t140, t139;\n\t"
        "madc.hi.cc.u32 t118, 0xfc632551, t140, t118;\n\t"
        "madc.lo.cc.u32 t119, 0xa7179e84, t140, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xa7179e84, t140, t120;\n\t"
        "madc.lo.cc.u32 t121, 0xffffffff, t140, t121;\n\t"
        "madc.hi.cc.u32 t122, 0xffffffff, t140, t122;\n\t"
        "madc.lo.cc.u32 t123, 0x0, t140, t123;\n\t"
        "madc.hi.cc.u32 t124, 0x0, t140, t124;\n\t"
        "addc.u32 t125, t125, 0x0;\n\t"
        "mad.lo.cc.u32 t100, %8, t90, t100;\n\t"
        "madc.hi.cc.u32 t101, %8, t90, t101;\n\t"
        "madc.lo.cc.u32 t102, %10, t90, t102;\n\t"
        "madc.hi.cc.u32 t103, %10, t90, t103;\n\t"
        "madc.lo.cc.u32 t104, %12, t90, t104;\n\t"
        "madc.hi.cc.u32 t105, %12, t90, t105;\n\t"
        "madc.lo.cc.u32 t106, %14, t90, t106;\n\t"
        "madc.hi.cc.u32 t107, %14, t90, t107;\n\t"
        "addc.u32 t108, t108, 0x0;\n\t"
        "mad.lo.cc.u32 t119, %9, t90, t119;\n\t"
        "madc.hi.cc.u32 t120, %9, t90, t120;\n\t"
        "madc.lo.cc.u32 t121, %11, t90, t121;\n\t"
        "madc.hi.cc.u32 t122, %11, t90, t122;\n\t"
        "madc.lo.cc.u32 t123, %13, t90, t123;\n\t"
        "madc.hi.cc.u32 t124, %13, t90, t124;\n\t"
        "madc.lo.cc.u32 t125, %15, t90, t125;\n\t"
        "madc.hi.u32 t126, %15, t90, 0x0;\n\t"
        "add.cc.u32 t141, t100, t118;\n\t"
        "mul.lo.u32 t142, t141, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t119, 0xf3b9cac2, t142, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xf3b9cac2, t142, t120;\n\t"
        "madc.lo.cc.u32 t121, 0xbce6faad, t142, t121;\n\t"
        "madc.hi.cc.u32 t122, 0xbce6faad, t142, t122;\n\t"
        "madc.lo.cc.u32 t123, 0xffffffff, t142, t123;\n\t"
        "madc.hi.cc.u32 t124, 0xffffffff, t142, t124;\n\t"
        "madc.lo.cc.u32 t125, 0xffffffff, t142, t125;\n\t"
        "madc.hi.cc.u32 t126, 0xffffffff, t142, t126;\n\t"
        "addc.u32 t127, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t141, 0xfc632551, t142, t141;\n\t"
        "madc.hi.cc.u32 t101, 0xfc632551, t142, t101;\n\t"
        "madc.lo.cc.u32 t102, 0xa7179e84, t142, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xa7179e84, t142, t103;\n\t"
        "madc.lo.cc.u32 t104, 0xffffffff, t142, t104;\n\t"
        "madc.hi.cc.u32 t105, 0xffffffff, t142, t105;\n\t"
        "madc.lo.cc.u32 t106, 0x0, t142, t106;\n\t"
        "madc.hi.cc.u32 t107, 0x0, t142, t107;\n\t"
        "addc.u32 t108, t108, 0x0;\n\t"
        "mad.lo.cc.u32 t119, %8, t93, t119;\n\t"
        "madc.hi.cc.u32 t120, %8, t93, t120;\n\t"
        "madc.lo.cc.u32 t121, %10, t93, t121;\n\t"
        "madc.hi.cc.u32 t122, %10, t93, t122;\n\t"
        "madc.lo.cc.u32 t123, %12, t93, t123;\n\t"
        "madc.hi.cc.u32 t124, %12, t93, t124;\n\t"
        "madc.lo.cc.u32 t125, %14, t93, t125;\n\t"
        "madc.hi.cc.u32 t126, %14, t93, t126;\n\t"
        "addc.u32 t127, t127, 0x0;\n\t"
        "mad.lo.cc.u32 t102, %9, t93, t102;\n\t"
        "madc.hi.cc.u32 t103, %9, t93, t103;\n\t"
        "madc.lo.cc.u32 t104, %11, t93, t104;\n\t"
        "madc.hi.cc.u32 t105, %11, t93, t105;\n\t"
        "madc.lo.cc.u32 t106, %13, t93, t106;\n\t"
        "madc.hi.cc.u32 t107, %13, t93, t107;\n\t"
        "madc.lo.cc.u32 t108, %15, t93, t108;\n\t"
        "madc.hi.u32 t109, %15, t93, 0x0;\n\t"
        "add.cc.u32 t143, t119, t101;\n\t"
        "mul.lo.u32 t144, t143, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t102, 0xf3b9cac2, t144, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xf3b9cac2, t144, t103;\n\t"
        "madc.lo.cc.u32 t104, 0xbce6faad, t144, t104;\n\t"
        "madc.hi.cc.u32 t105, 0xbce6faad, t144, t105;\n\t"
        "madc.lo.cc.u32 t106, 0xffffffff, t144, t106;\n\t"
        "madc.hi.cc.u32 t107, 0xffffffff, t144, t107;\n\t"
        "madc.lo.cc.u32 t108, 0xffffffff, t144, t108;\n\t"
        "madc.hi.cc.u32 t109, 0xffffffff, t144, t109;\n\t"
        "addc.u32 t110, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t143, 0xfc632551, t144, t143;\n\t"
        "madc.hi.cc.u32 t120, 0xfc632551, t144, t120;\n\t"
        "madc.lo.cc.u32 t121, 0xa7179e84, t144, t121;\n\t"
        "madc.hi.cc.u32 t122, 0xa7179e84, t144, t122;\n\t"
        "madc.lo.cc.u32 t123, 0xffffffff, t144, t123;\n\t"
        "madc.hi.cc.u32 t124, 0xffffffff, t144, t124;\n\t"
        "madc.lo.cc.u32 t125, 0x0, t144, t125;\n\t"
        "madc.hi.cc.u32 t126, 0x0, t144, t126;\n\t"
        "addc.u32 t127, t127, 0x0;\n\t"
        "add.cc.u32 t145, t102, t120;\n\t"
        "addc.cc.u32 t146, t103, t121;\n\t"
        "addc.cc.u32 t147, t104, t122;\n\t"
        "addc.cc.u32 t148, t105, t123;\n\t"
        "addc.cc.u32 t149, t106, t124;\n\t"
        "addc.cc.u32 t150, t107, t125;\n\t"
        "addc.cc.u32 t151, t108, t126;\n\t"
        "addc.cc.u32 t152, t109, t127;\n\t"
        "addc.u32 t153, t110, 0x0;\n\t"
        "sub.cc.u32 t154, t145, 0xfc632551;\n\t"
        "subc.cc.u32 t155, t146, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t156, t147, 0xa7179e84;\n\t"
        "subc.cc.u32 t157, t148, 0xbce6faad;\n\t"
        "subc.cc.u32 t158, t149, 0xffffffff;\n\t"
        "subc.cc.u32 t159, t150, 0xffffffff;\n\t"
        "subc.cc.u32 t160, t151, 0x0;\n\t"
        "subc.cc.u32 t161, t152, 0xffffffff;\n\t"
        "subc.cc.u32 t162, t153, 0x0;\n\t"
        "subc.u32 t163, 0x0, 0x0;\n\t"
        "xor.b32 t164, t154, t145;\n\t"
        "and.b32 t165, t164, t163;\n\t"
        "xor.b32 t166, t165, t154;\n\t"
        "xor.b32 t167, t155, t146;\n\t"
        "and.b32 t168, t167, t163;\n\t"
        "xor.b32 t169, t168, t155;\n\t"
        "xor.b32 t170, t156, t147;\n\t"
        "and.b32 t171, t170, t163;\n\t"
        "xor.b32 t172, t171, t156;\n\t"
        "xor.b32 t173, t157, t148;\n\t"
        "and.b32 t174, t173, t163;\n\t"
        "xor.b32 t175, t174, t157;\n\t"
        "xor.b32 t176, t158, t149;\n\t"
        "and.b32 t177, t176, t163;\n\t"
        "xor.b32 t178, t177, t158;\n\t"
        "xor.b32 t179, t159, t150;\n\t"
        "and.b32 t180, t179, t163;\n\t"
        "xor.b32 t181, t180, t159;\n\t"
        "xor.b32 t182, t160, t151;\n\t"
        "and.b32 t183, t182, t163;\n\t"
        "xor.b32 t184, t183, t160;\n\t"
        "xor.b32 t185, t161, t152;\n\t"
        "and.b32 t186, t185, t163;\n\t"
        "xor.b32 t187, t186, t161;\n\t"
        "mov.u32 %0, t166;\n\t"
        "mov.u32 %1, t169;\n\t"
        "mov.u32 %2, t172;\n\t"
        "mov.u32 %3, t175;\n\t"
        "mov.u32 %4, t178;\n\t"
        "mov.u32 %5, t181;\n\t"
        "mov.u32 %6, t184;\n\t"
        "mov.u32 %7, t187;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127, t128, t129, t130, t131, t132, t133, t134, t135, t136, t137, t138, t139, t140, t141, t142, t143, t144, t145, t146, t147, t148, t149, t150, t151, t152, t153, t154, t155, t156, t157, t158, t159, t160, t161, t162, t163, t164, t165, t166, t167, t168, t169, t170, t171, t172, t173, t174, t175, t176, t177, t178, t179, t180, t181, t182, t183, t184, t185, t186, t187;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t0 = (uint32_t)((uint32_t)(0xbe79eea2u * b_i));
    t1 = (uint32_t)(((uint64_t)0xbe79eea2u * b_i) >> 32);
    t2 = (uint32_t)((uint32_t)(0x49bd6fa6u * b_i));
    t3 = (uint32_t)(((uint64_t)0x49bd6fa6u * b_i) >> 32);
    t4 = (uint32_t)((uint32_t)(0x2b6bec59u * b_i));
    t5 = (uint32_t)(((uint64_t)0x2b6bec59u * b_i) >> 32);
    t6 = (uint32_t)((uint32_t)(0xf3d95620u * b_i));
    t7 = (uint32_t)(((uint64_t)0xf3d95620u * b_i) >> 32);
    t19 = (uint32_t)((uint32_t)(0x83244c95u * b_i));
    t20 = (uint32_t)(((uint64_t)0x83244c95u * b_i) >> 32);
    t21 = (uint32_t)((uint32_t)(0x4699799cu * b_i));
    t22 = (uint32_t)(((uint64_t)0x4699799cu * b_i) >> 32);
    t23 = (uint32_t)((uint32_t)(0x2845b239u * b_i));
    t24 = (uint32_t)(((uint64_t)0x2845b239u * b_i) >> 32);
    t25 = (uint32_t)((uint32_t)(0x66e12d94u * b_i));
    t26 = (uint32_t)(((uint64_t)0x66e12d94u * b_i) >> 32);
    t36 = (uint32_t)((uint32_t)(t0 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t36) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t36) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t36) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t36) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t36) + t0; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t36) >> 32) + t1 + cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t36) + t2 + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t36) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t36) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t36) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t8 = (uint32_t)w_;
    w_ = (uint64_t)t19 + t1; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t38 = (uint32_t)((uint32_t)(t37 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t38) + t2 + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t38) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t38) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t38) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + 0x0u + cf_; t9 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t38) + t37; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t38) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t38) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t38) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t38) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t38) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t27 + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)t2 + t20; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t40 = (uint32_t)((uint32_t)(t39 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t40) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t40) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t40) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t40) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + 0x0u + cf_; t28 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t40) + t39; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t40) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t40) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t40) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t40) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t40) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)t21 + t3; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t42 = (uint32_t)((uint32_t)(t41 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t42) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t42) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t42) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t42) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + 0x0u + cf_; t11 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t42) + t41; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t42) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t42) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t42) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t42) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t42) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)t4 + t22; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t44 = (uint32_t)((uint32_t)(t43 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t44) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t44) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t44) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t44) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + 0x0u + cf_; t30 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t44) + t43; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t44) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t44) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t44) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t44) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t44) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)t23 + t5; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t46 = (uint32_t)((uint32_t)(t45 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t46) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t46) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t46) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t46) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + 0x0u + cf_; t13 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t46) + t45; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t46) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t46) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t46) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t46) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t46) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t31 = (uint32_t)w_;
    w_ = (uint64_t)t6 + t24; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t48 = (uint32_t)((uint32_t)(t47 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t48) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t48) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t48) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t48) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t48) + t47; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t48) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t48) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t48) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t48) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t48) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t14 = (uint32_t)w_;
    w_ = (uint64_t)t25 + t7; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t50 = (uint32_t)((uint32_t)(t49 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t50) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t50) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t50) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t50) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + 0x0u + cf_; t15 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t50) + t49; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t50) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t50) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t50) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t50) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t50) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t33 = (uint32_t)w_;
    w_ = (uint64_t)t8 + t26; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t27 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t28 + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t29 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t30 + cf_; t55 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t31 + cf_; t56 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + t32 + cf_; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + t33 + cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t59 = (uint32_t)w_;
    w_ = (uint64_t)t51 - 0xfc632551u; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t52 - 0xf3b9cac2u - cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t53 - 0xa7179e84u - cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t54 - 0xbce6faadu - cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t55 - 0xffffffffu - cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t56 - 0xffffffffu - cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t57 - 0x0u - cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t58 - 0xffffffffu - cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t59 - 0x0u - cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t69 = (uint32_t)w_;
    t70 = (uint32_t)(t60 ^ t51);
    t71 = (uint32_t)(t70 & t69);
    t72 = (uint32_t)(t71 ^ t60);
    t73 = (uint32_t)(t61 ^ t52);
    t74 = (uint32_t)(t73 & t69);
    t75 = (uint32_t)(t74 ^ t61);
    t76 = (uint32_t)(t62 ^ t53);
    t77 = (uint32_t)(t76 & t69);
    t78 = (uint32_t)(t77 ^ t62);
    t79 = (uint32_t)(t63 ^ t54);
    t80 = (uint32_t)(t79 & t69);
    t81 = (uint32_t)(t80 ^ t63);
    t82 = (uint32_t)(t64 ^ t55);
    t83 = (uint32_t)(t82 & t69);
    t84 = (uint32_t)(t83 ^ t64);
    t85 = (uint32_t)(t65 ^ t56);
    t86 = (uint32_t)(t85 & t69);
    t87 = (uint32_t)(t86 ^ t65);
    t88 = (uint32_t)(t66 ^ t57);
    t89 = (uint32_t)(t88 & t69);
    t90 = (uint32_t)(t89 ^ t66);
    t91 = (uint32_t)(t67 ^ t58);
    t92 = (uint32_t)(t91 & t69);
    t93 = (uint32_t)(t92 ^ t67);
    t94 = (uint32_t)((uint32_t)(a_0_i * t72));
    t95 = (uint32_t)(((uint64_t)a_0_i * t72) >> 32);
    t96 = (uint32_t)((uint32_t)(a_2_i * t72));
    t97 = (uint32_t)(((uint64_t)a_2_i * t72) >> 32);
    t98 = (uint32_t)((uint32_t)(a_4_i * t72));
    t99 = (uint32_t)(((uint64_t)a_4_i * t72) >> 32);
    t100 = (uint32_t)((uint32_t)(a_6_i * t72));
    t101 = (uint32_t)(((uint64_t)a_6_i * t72) >> 32);
    t113 = (uint32_t)((uint32_t)(a_1_i * t72));
    t114 = (uint32_t)(((uint64_t)a_1_i * t72) >> 32);
    t115 = (uint32_t)((uint32_t)(a_3_i * t72));
    t116 = (uint32_t)(((uint64_t)a_3_i * t72) >> 32);
    t117 = (uint32_t)((uint32_t)(a_5_i * t72));
    t118 = (uint32_t)(((uint64_t)a_5_i * t72) >> 32);
    t119 = (uint32_t)((uint32_t)(a_7_i * t72));
    t120 = (uint32_t)(((uint64_t)a_7_i * t72) >> 32);
    t130 = (uint32_t)((uint32_t)(t94 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t130) + t113; t113 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t130) >> 32) + t114 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t130) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t130) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t130) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t130) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t130) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t130) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t121 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t130) + t94; t94 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t130) >> 32) + t95 + cf_; t95 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t130) + t96 + cf_; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t130) >> 32) + t97 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t130) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t130) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t130) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t130) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t102 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t75) + t113; t113 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t75) >> 32) + t114 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t75) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t75) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t75) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t75) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t75) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t75) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t121 + 0x0u + cf_; t121 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t75) + t96; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t75) >> 32) + t97 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t75) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t75) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t75) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t75) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t75) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t75) >> 32) + 0x0u + cf_; t103 = (uint32_t)w_;
    w_ = (uint64_t)t113 + t95; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t132 = (uint32_t)((uint32_t)(t131 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t132) + t96 + cf_; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t132) >> 32) + t97 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t132) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t132) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t132) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t132) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t132) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t132) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t104 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t132) + t131; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t132) >> 32) + t114 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t132) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t132) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t132) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t132) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t132) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t132) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t121 + 0x0u + cf_; t121 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t78) + t96; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t78) >> 32) + t97 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t78) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t78) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t78) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t78) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t78) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t78) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t104 + 0x0u + cf_; t104 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t78) + t115; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t78) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t78) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t78) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t78) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t78) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t78) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t78) >> 32) + 0x0u + cf_; t122 = (uint32_t)w_;
    w_ = (uint64_t)t96 + t114; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t134 = (uint32_t)((uint32_t)(t133 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t134) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t134) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t134) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t134) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t134) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t134) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t134) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t134) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t123 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t134) + t133; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t134) >> 32) + t97 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t134) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t134) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t134) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t134) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t134) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t134) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t104 + 0x0u + cf_; t104 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t81) + t115; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t81) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t81) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t81) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t81) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t81) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t81) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t81) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t123 + 0x0u + cf_; t123 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t81) + t98; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t81) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t81) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t81) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t81) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t81) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t81) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t81) >> 32) + 0x0u + cf_; t105 = (uint32_t)w_;
    w_ = (uint64_t)t115 + t97; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t136 = (uint32_t)((uint32_t)(t135 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t136) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t136) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t136) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t136) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t136) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t136) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t136) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t136) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t106 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t136) + t135; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t136) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t136) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t136) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t136) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t136) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t136) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t136) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t123 + 0x0u + cf_; t123 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t84) + t98; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t84) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t84) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t84) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t84) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t84) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t84) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t84) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t106 + 0x0u + cf_; t106 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t84) + t117; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t84) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t84) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t84) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t84) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t84) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t84) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t84) >> 32) + 0x0u + cf_; t124 = (uint32_t)w_;
    w_ = (uint64_t)t98 + t116; t137 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t138 = (uint32_t)((uint32_t)(t137 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t138) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t138) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t138) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t138) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t138) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t138) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t138) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t138) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t125 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t138) + t137; t137 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t138) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t138) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t138) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t138) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t138) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t138) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t138) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t106 + 0x0u + cf_; t106 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t87) + t117; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t87) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t87) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t87) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t87) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t87) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t87) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t87) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t125 + 0x0u + cf_; t125 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t87) + t100; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t87) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t87) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t87) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t87) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t87) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t87) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t87) >> 32) + 0x0u + cf_; t107 = (uint32_t)w_;
    w_ = (uint64_t)t117 + t99; t139 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t140 = (uint32_t)((uint32_t)(t139 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t140) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t140) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t140) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t140) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t140) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t140) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t140) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t140) >> 32) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t108 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t140) + t139; t139 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t140) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t140) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t140) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t140) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t140) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t140) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t140) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t125 + 0x0u + cf_; t125 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t90) + t100; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t90) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t90) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t90) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t90) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t90) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t90) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t90) >> 32) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t108 + 0x0u + cf_; t108 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t90) + t119; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t90) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t90) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t90) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t90) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t90) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t90) + t125 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t90) >> 32) + 0x0u + cf_; t126 = (uint32_t)w_;
    w_ = (uint64_t)t100 + t118; t141 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t142 = (uint32_t)((uint32_t)(t141 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t142) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t142) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t142) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t142) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t142) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t142) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t142) + t125 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t142) >> 32) + t126 + cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t127 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t142) + t141; t141 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t142) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t142) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t142) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t142) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t142) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t142) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t142) >> 32) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t108 + 0x0u + cf_; t108 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t93) + t119; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t93) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t93) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t93) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t93) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t93) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t93) + t125 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t93) >> 32) + t126 + cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t127 + 0x0u + cf_; t127 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t93) + t102; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t93) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t93) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t93) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t93) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t93) >> 32) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t93) + t108 + cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t93) >> 32) + 0x0u + cf_; t109 = (uint32_t)w_;
    w_ = (uint64_t)t119 + t101; t143 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t144 = (uint32_t)((uint32_t)(t143 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t144) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t144) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t144) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t144) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t144) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t144) >> 32) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t144) + t108 + cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t144) >> 32) + t109 + cf_; t109 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t110 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t144) + t143; t143 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t144) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t144) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t144) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t144) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t144) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t144) + t125 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t144) >> 32) + t126 + cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t127 + 0x0u + cf_; t127 = (uint32_t)w_;
    w_ = (uint64_t)t102 + t120; t145 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t103 + t121 + cf_; t146 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t104 + t122 + cf_; t147 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t105 + t123 + cf_; t148 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t106 + t124 + cf_; t149 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t107 + t125 + cf_; t150 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t108 + t126 + cf_; t151 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t109 + t127 + cf_; t152 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t110 + 0x0u + cf_; t153 = (uint32_t)w_;
    w_ = (uint64_t)t145 - 0xfc632551u; t154 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t146 - 0xf3b9cac2u - cf_; t155 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t147 - 0xa7179e84u - cf_; t156 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t148 - 0xbce6faadu - cf_; t157 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t149 - 0xffffffffu - cf_; t158 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t150 - 0xffffffffu - cf_; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t151 - 0x0u - cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t152 - 0xffffffffu - cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t153 - 0x0u - cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t163 = (uint32_t)w_;
    t164 = (uint32_t)(t154 ^ t145);
    t165 = (uint32_t)(t164 & t163);
    t166 = (uint32_t)(t165 ^ t154);
    t167 = (uint32_t)(t155 ^ t146);
    t168 = (uint32_t)(t167 & t163);
    t169 = (uint32_t)(t168 ^ t155);
    t170 = (uint32_t)(t156 ^ t147);
    t171 = (uint32_t)(t170 & t163);
    t172 = (uint32_t)(t171 ^ t156);
    t173 = (uint32_t)(t157 ^ t148);
    t174 = (uint32_t)(t173 & t163);
    t175 = (uint32_t)(t174 ^ t157);
    t176 = (uint32_t)(t158 ^ t149);
    t177 = (uint32_t)(t176 & t163);
    t178 = (uint32_t)(t177 ^ t158);
    t179 = (uint32_t)(t159 ^ t150);
    t180 = (uint32_t)(t179 & t163);
    t181 = (uint32_t)(t180 ^ t159);
    t182 = (uint32_t)(t160 ^ t151);
    t183 = (uint32_t)(t182 & t163);
    t184 = (uint32_t)(t183 ^ t160);
    t185 = (uint32_t)(t161 ^ t152);
    t186 = (uint32_t)(t185 & t163);
    t187 = (uint32_t)(t186 ^ t161);
    r[0] = t166;
    r[1] = t169;
    r[2] = t172;
    r[3] = t175;
    r[4] = t178;
    r[5] = t181;
    r[6] = t184;
    r[7] = t187;
#endif
  }

  // r = a*b + c, small integer b: modmli + modadd fused (rfc7748.c:209,212)
  static MAB_DEV void mla(uint32_t (&r)[8], const uint32_t (&a)[8], uint32_t b, const uint32_t (&c)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<231>;\n\t"
        "mul.lo.u32 t0, 0xbe79eea2, %24;\n\t"
        "mul.hi.u32 t1, 0xbe79eea2, %24;\n\t"
        "mul.lo.u32 t2, 0x49bd6fa6, %24;\n\t"
        "mul.hi.u32 t3, 0x49bd6fa6, %24;\n\t"
        "mul.lo.u32 t4, 0x2b6bec59, %24;\n\t"
        "mul.hi.u32 t5, 0x2b6bec59, %24;\n\t"
        "mul.lo.u32 t6, 0xf3d95620, %24;\n\t"
        "mul.hi.u32 t7, 0xf3d95620, %24;\n\t"
        "mul.lo.u32 t19, 0x83244c95, %24;\n\t"
        "mul.hi.u32 t20, 0x83244c95, %24;\n\t"
        "mul.lo.u32 t21, 0x4699799c, %24;\n\t"
        "mul.hi.u32 t22, 0x4699799c, %24;\n\t"
        "mul.lo.u32 t23, 0x2845b239, %24;\n\t"
        "mul.hi.u32 t24, 0x2845b239, %24;\n\t"
        "mul.lo.u32 t25, 0x66e12d94, %24;\n\t"
        "mul.hi.u32 t26, 0x66e12d94, %24;\n\t"
        "mul.lo.u32 t36, t0, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t19, 0xf3b9cac2, t36, t19;\n\t"
        "madc.hi.cc.u32 t20, 0xf3b9cac2, t36, t20;\n\t"
        "madc.lo.cc.u32 t21, 0xbce6faad, t36, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xbce6faad, t36, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xffffffff, t36, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xffffffff, t36, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t36, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t36, t26;\n\t"
        "addc.u32 t27, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t0, 0xfc632551, t36, t0;\n\t"
        "madc.hi.cc.u32 t1, 0xfc632551, t36, t1;\n\t"
        "madc.lo.cc.u32 t2, 0xa7179e84, t36, t2;\n\t"
        "madc.hi.cc.u32 t3, 0xa7179e84, t36, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xffffffff, t36, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xffffffff, t36, t5;\n\t"
        "madc.lo.cc.u32 t6, 0x0, t36, t6;\n\t"
        "madc.hi.cc.u32 t7, 0x0, t36, t7;\n\t"
        "addc.u32 t8, 0x0, 0x0;\n\t"
        "add.cc.u32 t37, t19, t1;\n\t"
        "mul.lo.u32 t38, t37, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t2, 0xf3b9cac2, t38, t2;\n\t"
        "madc.hi.cc.u32 t3, 0xf3b9cac2, t38, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xbce6faad, t38, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xbce6faad, t38, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xffffffff, t38, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xffffffff, t38, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t38, t8;\n\t"
        "madc.hi.u32 t9, 0xffffffff, t38, 0x0;\n\t"
        "mad.lo.cc.u32 t37, 0xfc632551, t38, t37;\n\t"
        "madc.hi.cc.u32 t20, 0xfc632551, t38, t20;\n\t"
        "madc.lo.cc.u32 t21, 0xa7179e84, t38, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xa7179e84, t38, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xffffffff, t38, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xffffffff, t38, t24;\n\t"
        "madc.lo.cc.u32 t25, 0x0, t38, t25;\n\t"
        "madc.hi.cc.u32 t26, 0x0, t38, t26;\n\t"
        "addc.u32 t27, t27, 0x0;\n\t"
        "add.cc.u32 t39, t2, t20;\n\t"
        "mul.lo.u32 t40, t39, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t21, 0xf3b9cac2, t40, t21;\n\t"
        "madc.hi.cc.u32 t22, 0xf3b9cac2, t40, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xbce6faad, t40, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xbce6faad, t40, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t40, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t40, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t40, t27;\n\t"
        "madc.hi.u32 t28, 0xffffffff, t40, 0x0;\n\t"
        "mad.lo.cc.u32 t39, 0xfc632551, t40, t39;\n\t"
        "madc.hi.cc.u32 t3, 0xfc632551, t40, t3;\n\t"
        "madc.lo.cc.u32 t4, 0xa7179e84, t40, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xa7179e84, t40, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xffffffff, t40, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xffffffff, t40, t7;\n\t"
        "madc.lo.cc.u32 t8, 0x0, t40, t8;\n\t"
        "madc.hi.cc.u32 t9, 0x0, t40, t9;\n\t"
        "addc.u32 t10, 0x0, 0x0;\n\t"
        "add.cc.u32 t41, t21, t3;\n\t"
        "mul.lo.u32 t42, t41, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t4, 0xf3b9cac2, t42, t4;\n\t"
        "madc.hi.cc.u32 t5, 0xf3b9cac2, t42, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xbce6faad, t42, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xbce6faad, t42, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t42, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xffffffff, t42, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t42, t10;\n\t"
        "madc.hi.u32 t11, 0xffffffff, t42, 0x0;\n\t"
        "mad.lo.cc.u32 t41, 0xfc632551, t42, t41;\n\t"
        "madc.hi.cc.u32 t22, 0xfc632551, t42, t22;\n\t"
        "madc.lo.cc.u32 t23, 0xa7179e84, t42, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xa7179e84, t42, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xffffffff, t42, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xffffffff, t42, t26;\n\t"
        "madc.lo.cc.u32 t27, 0x0, t42, t27;\n\t"
        "madc.hi.cc.u32 t28, 0x0, t42, t28;\n\t"
        "addc.u32 t29, 0x0, 0x0;\n\t"
        "add.cc.u32 t43, t4, t22;\n\t"
        "mul.lo.u32 t44, t43, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t23, 0xf3b9cac2, t44, t23;\n\t"
        "madc.hi.cc.u32 t24, 0xf3b9cac2, t44, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xbce6faad, t44, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xbce6faad, t44, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t44, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xffffffff, t44, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t44, t29;\n\t"
        "madc.hi.u32 t30, 0xffffffff, t44, 0x0;\n\t"
        "mad.lo.cc.u32 t43, 0xfc632551, t44, t43;\n\t"
        "madc.hi.cc.u32 t5, 0xfc632551, t44, t5;\n\t"
        "madc.lo.cc.u32 t6, 0xa7179e84, t44, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xa7179e84, t44, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xffffffff, t44, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xffffffff, t44, t9;\n\t"
        "madc.lo.cc.u32 t10, 0x0, t44, t10;\n\t"
        "madc.hi.cc.u32 t11, 0x0, t44, t11;\n\t"
        "addc.u32 t12, 0x0, 0x0;\n\t"
        "add.cc.u32 t45, t23, t5;\n\t"
        "mul.lo.u32 t46, t45, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t6, 0xf3b9cac2, t46, t6;\n\t"
        "madc.hi.cc.u32 t7, 0xf3b9cac2, t46, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xbce6faad, t46, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xbce6faad, t46, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t46, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xffffffff, t46, t11;\n\t"
        "madc.lo.cc.u32 t12, 0xffffffff, t46, t12;\n\t"
        "madc.hi.u32 t13, 0xffffffff, t46, 0x0;\n\t"
        "mad.lo.cc.u32 t45, 0xfc632551, t46, t45;\n\t"
        "madc.hi.cc.u32 t24, 0xfc632551, t46, t24;\n\t"
        "madc.lo.cc.u32 t25, 0xa7179e84, t46, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xa7179e84, t46, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xffffffff, t46, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xffffffff, t46, t28;\n\t"
        "madc.lo.cc.u32 t29, 0x0, t46, t29;\n\t"
        "madc.hi.cc.u32 t30, 0x0, t46, t30;\n\t"
        "addc.u32 t31, 0x0, 0x0;\n\t"
        "add.cc.u32 t47, t6, t24;\n\t"
        "mul.lo.u32 t48, t47, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t25, 0xf3b9cac2, t48, t25;\n\t"
        "madc.hi.cc.u32 t26, 0xf3b9cac2, t48, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xbce6faad, t48, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xbce6faad, t48, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t48, t29;\n\t"
        "madc.hi.cc.u32 t30, 0xffffffff, t48, t30;\n\t"
        "madc.lo.cc.u32 t31, 0xffffffff, t48, t31;\n\t"
        "madc.hi.u32 t32, 0xffffffff, t48, 0x0;\n\t"
        "mad.lo.cc.u32 t47, 0xfc632551, t48, t47;\n\t"
        "madc.hi.cc.u32 t7, 0xfc632551, t48, t7;\n\t"
        "madc.lo.cc.u32 t8, 0xa7179e84, t48, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xa7179e84, t48, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xffffffff, t48, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xffffffff, t48, t11;\n\t"
        "madc.lo.cc.u32 t12, 0x0, t48, t12;\n\t"
        "madc.hi.cc.u32 t13, 0x0, t48, t13;\n\t"
        "addc.u32 t14, 0x0, 0x0;\n\t"
        "add.cc.u32 t49, t25, t7;\n\t"
        "mul.lo.u32 t50, t49, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t8, 0xf3b9cac2, t50, t8;\n\t"
        "madc.hi.cc.u32 t9, 0xf3b9cac2, t50, t9;\n\t"
        "madc.lo.cc.u32 t10, 0xbce6faad, t50, t10;\n\t"
        "madc.hi.cc.u32 t11, 0xbce6faad, t50, t11;\n\t"
        "madc.lo.cc.u32 t12, 0xffffffff, t50, t12;\n\t"
        "madc.hi.cc.u32 t13, 0xffffffff, t50, t13;\n\t"
        "madc.lo.cc.u32 t14, 0xffffffff, t50, t14;\n\t"
        "madc.hi.u32 t15, 0xffffffff, t50, 0x0;\n\t"
        "mad.lo.cc.u32 t49, 0xfc632551, t50, t49;\n\t"
        "madc.hi.cc.u32 t26, 0xfc632551, t50, t26;\n\t"
        "madc.lo.cc.u32 t27, 0xa7179e84, t50, t27;\n\t"
        "madc.hi.cc.u32 t28, 0xa7179e84, t50, t28;\n\t"
        "madc.lo.cc.u32 t29, 0xffffffff, t50, t29;\n\t"
        "madc.hi.cc.u32 t30, 0xffffffff, t50, t30;\n\t"
        "madc.lo.cc.u32 t31, 0x0, t50, t31;\n\t"
        "madc.hi.cc.u32 t32, 0x0, t50, t32;\n\t"
        "addc.u32 t33, 0x0, 0x0;\n\t"
        "add.cc.u32 t51, t8, t26;\n\t"
        "addc.cc.u32 t52, t9, t27;\n\t"
        "addc.cc.u32 t53, t10, t28;\n\t"
        "addc.cc.u32 t54, t11, t29;\n\t"
        "addc.cc.u32 t55, t12, t30;\n\t"
        "addc.cc.u32 t56, t13, t31;\n\t"
        "addc.cc.u32 t57, t14, t32;\n\t"
        "addc.cc.u32 t58, t15, t33;\n\t"
        "addc.u32 t59, 0x0, 0x0;\n\t"
        "sub.cc.u32 t60, t51, 0xfc632551;\n\t"
        "subc.cc.u32 t61, t52, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t62, t53, 0xa7179e84;\n\t"
        "subc.cc.u32 t63, t54, 0xbce6faad;\n\t"
        "subc.cc.u32 t64, t55, 0xffffffff;\n\t"
        "subc.cc.u32 t65, t56, 0xffffffff;\n\t"
        "subc.cc.u32 t66, t57, 0x0;\n\t"
        "subc.cc.u32 t67, t58, 0xffffffff;\n\t"
        "subc.cc.u32 t68, t59, 0x0;\n\t"
        "subc.u32 t69, 0x0, 0x0;\n\t"
        "xor.b32 t70, t60, t51;\n\t"
        "and.b32 t71, t70, t69;\n\t"
        "xor.b32 t72, t71, t60;\n\t"
        "xor.b32 t73, t61, t52;\n\t"
        "and.b32 t74, t73, t69;\n\t"
        "xor.b32 t75, t74, t61;\n\t"
        "xor.b32 t76, t62, t53;\n\t"
        "and.b32 t77, t76, t69;\n\t"
        "xor.b32 t78, t77, t62;\n\t"
        "xor.b32 t79, t63, t54;\n\t"
        "and.b32 t80, t79, t69;\n\t"
        "xor.b32 t81, t80, t63;\n\t"
        "xor.b32 t82, t64, t55;\n\t"
        "and.b32 t83, t82, t69;\n\t"
        "xor.b32 t84, t83, t64;\n\t"
        "xor.b32 t85, t65, t56;\n\t"
        "and.b32 t86, t85, t69;\n\t"
        "xor.b32 t87, t86, t65;\n\t"
        "xor.b32 t88, t66, t57;\n\t"
        "and.b32 t89, t88, t69;\n\t"
        "xor.b32 t90, t89, t66;\n\t"
        "xor.b32 t91, t67, t58;\n\t"
        "and.b32 t92, t91, t69;\n\t"
        "xor.b32 t93, t92, t67;\n\t"
        "mul.lo.u32 t94, %8, t72;\n\t"
        "mul.hi.u32 t95, %8, t72;\n\t"
        "mul.lo.u32 t96, %10, t72;\n\t"
        "mul.hi.u32 t97, %10, t72;\n\t"
        "mul.lo.u32 t98, %12, t72;\n\t"
        "mul.hi.u32 t99, %12, t72;\n\t"
        "mul.lo.u32 t100, %14, t72;\n\t"
        "mul.hi.u32 t101, %14, t72;\n\t"
        "mul.lo.u32 t113, %9, t72;\n\t"
        "mul.hi.u32 t114, %9, t72;\n\t"
        "mul.lo.u32 t115, %11, t72;\n\t"
        "mul.hi.u32 t116, %11, t72;\n\t"
        "mul.lo.u32 t117, %13, t72;\n\t"
        "mul.hi.u32 t118, %13, t72;\n\t"
        "mul.lo.u32 t119, %15, t72;\n\t"
        "mul.hi.u32 t120, %15, t72;\n\t"
        "mul.lo.u32 t130, t94, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t113, 0xf3b9cac2, t130, t113;\n\t"
        "madc.hi.cc.u32 t114, 0xf3b9cac2, t130, t114;\n\t"
        "madc.lo.cc.u32 t115, 0xbce6faad, t130, t115;\n\t"
        "madc.hi.cc.u32 t116, 0xbce6faad, t130, t116;\n\t"
        "madc.lo.cc.u32 t117, 0xffffffff, t130, t117;\n\t"
        "madc.hi.cc.u32 t118, 0xffffffff, t130, t118;\n\t"
        "madc.lo.cc.u32 t119, 0xffffffff, t130, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xffffffff, t130, t120;\n\t"
        "addc.u32 t121, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t94, 0xfc632551, t130, t94;\n\t"
        "madc.hi.cc.u32 t95, 0xfc632551, t130, t95;\n\t"
        "madc.lo.cc.u32 t96, 0xa7179e84, t130, t96;\n\t"
        "madc.hi.cc.u32 t97, 0xa7179e84, t130, t97;\n\t"
        "madc.lo.cc.u32 t98, 0xffffffff, t130, t98;\n\t"
        "madc.hi.cc.u32 t99, 0xffffffff, t130, t99;\n\t"
        "madc.lo.cc.u32 t100, 0x0, t130, t100;\n\t"
        "madc.hi.cc.u32 t101, 0x0, t130, t101;\n\t"
        "addc.u32 t102, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t113, %8, t75, t113;\n\t"
        "madc.hi.cc.u32 t114, %8, t75, t114;\n\t"
        "madc.lo.cc.u32 t115, %10, t75, t115;\n\t"
        "madc.hi.cc.u32 t116, %10, t75, t116;\n\t"
        "madc.lo.cc.u32 t117, %12, t75, t117;\n\t"
        "madc.hi.cc.u32 t118, %12, t75, t118;\n\t"
        "madc.lo.cc.u32 t119, %14, t75, t119;\n\t"
        "madc.hi.cc.u32 t120, %14, t75, t120;\n\t"
        "addc.u32 t121, t121, 0x0;\n\t"
        "mad.lo.cc.u32 t96, %9, t75, t96;\n\t"
        "madc.hi.cc.u32 t97, %9, t75, t97;\n\t"
        "madc.lo.cc.u32 t98, %11, t75, t98;\n\t"
        "madc.hi.cc.u32 t99, %11, t75, t99;\n\t"
        "madc.lo.cc.u32 t100, %13, t75, t100;\n\t"
        "madc.hi.cc.u32 t101, %13, t75, t101;\n\t"
        "madc.lo.cc.u32 t102, %15, t75, t102;\n\t"
        "madc.hi.u32 t103, %15, t75, 0x0;\n\t"
        "add.cc.u32 t131, t113, t95;\n\t"
        "mul.lo.u32 t132, t131, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t96, 0xf3b9cac2, t132, t96;\n\t"
        "madc.hi.cc.u32 t97, 0xf3b9cac2, t132, t97;\n\t"
        "madc.lo.cc.u32 t98, 0xbce6faad, t132, t98;\n\t"
        "madc.hi.cc.u32 t99, 0xbce6faad, t132, t99;\n\t"
        "madc.lo.cc.u32 t100, 0xffffffff, t132, t100;\n\t"
        "madc.hi.cc.u32 t101, 0xffffffff, t132, t101;\n\t"
        "madc.lo.cc.u32 t102, 0xffffffff, t132, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xffffffff, t132, t103;\n\t"
        "addc.u32 t104, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t131, 0xfc632551, t132, t131;\n\t"
        "madc.hi.cc.u32 t114, 0xfc632551, t132, t114;\n\t"
        "madc.lo.cc.u32 t115, 0xa7179e84, t132, t115;\n\t"
        "madc.hi.cc.u32 t116, 0xa7179e84, t132, t116;\n\t"
        "madc.lo.cc.u32 t117, 0xffffffff, t132, t117;\n\t"
        "madc.hi.cc.u32 t118, 0xffffffff, t132, t118;\n\t"
        "madc.lo.cc.u32 t119, 0x0, t132, t119;\n\t"
        "madc.hi.cc.u32 t120, 0x0, t132, t120;\n\t"
        "addc.u32 t121, t121, 0x0;\n\t"
        "mad.lo.cc.u32 t96, %8, t78, t96;\n\t"
        "madc.hi.cc.u32 t97, %8, t78, t97;\n\t"
        "madc.lo.cc.u32 t98, %10, t78, t98;\n\t"
        "madc.hi.cc.u32 t99, %10, t78, t99;\n\t"
        "madc.lo.cc.u32 t100, %12, t78, t100;\n\t"
        "madc.hi.cc.u32 t101, %12, t78, t101;\n\t"
        "madc.lo.cc.u32 t102, %14, t78, t102;\n\t"
        "madc.hi.cc.u32 t103, %14, t78, t103;\n\t"
        "addc.u32 t104, t104, 0x0;\n\t"
        "mad.lo.cc.u32 t115, %9, t78, t115;\n\t"
        "madc.hi.cc.u32 t116, %9, t78, t116;\n\t"
        "madc.lo.cc.u32 t117, %11, t78, t117;\n\t"
        "madc.hi.cc.u32 t118, %11, t78, t118;\n\t"
        "madc.lo.cc.u32 t119, %13, t78, t119;\n\t"
        "madc.hi.cc.u32 t120, %13, t78, t120;\n\t"
        "madc.lo.cc.u32 t121, %15, t78, t121;\n\t"
        "madc.hi.u32 t122, %15, t78, 0x0;\n\t"
        "add.cc.u32 t133, t96, t114;\n\t"
        "mul.lo.u32 t134, t133, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t115, 0xf3b9cac2, t134, t115;\n\t"
        "madc.hi.cc.u32 t116, 0xf3b9cac2, t134, t116;\n\t"
        "madc.lo.cc.u32 t117, 0xbce6faad, t134, t117;\n\t"
        "madc.hi.cc.u32 t118, 0xbce6faad, t134, t118;\n\t"
        "madc.lo.cc.u32 t119, 0xffffffff, t134, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xffffffff, t134, t120;\n\t"
        "madc.lo.cc.u32 t121, 0xffffffff, t134, t121;\n\t"
        "madc.hi.cc.u32 t122, 0xffffffff, t134, t122;\n\t"
        "addc.u32 t123, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t133, 0xfc632551, t134, t133;\n\t"
        "madc.hi.cc.u32 t97, 0xfc632551, t134, t97;\n\t"
        "madc.lo.cc.u32 t98, 0xa7179e84, t134, t98;\n\t"
        "madc.hi.cc.u32 t99, 0xa7179e84, t134, t99;\n\t"
        "madc.lo.cc.u32 t100, 0xffffffff, t134, t100;\n\t"
        "madc.hi.cc.u32 t101, 0xffffffff, t134, t101;\n\t"
        "madc.lo.cc.u32 t102, 0x0, t134, t102;\n\t"
        "madc.hi.cc.u32 t103, 0x0, t134, t103;\n\t"
        "addc.u32 t104, t104, 0x0;\n\t"
        "mad.lo.cc.u32 t115, %8, t81, t115;\n\t"
        "madc.hi.cc.u32 t116, %8, t81, t116;\n\t"
        "madc.lo.cc.u32 t117, %10, t81, t117;\n\t"
        "madc.hi.cc.u32 t118, %10, t81, t118;\n\t"
        "madc.lo.cc.u32 t119, %12, t81, t119;\n\t"
        "madc.hi.cc.u32 t120, %12, t81, t120;\n\t"
        "madc.lo.cc.u32 t121, %14, t81, t121;\n\t"
        "madc.hi.cc.u32 t122, %14, t81, t122;\n\t"
        "addc.u32 t123, t123, 0x0;\n\t"
        "mad.lo.cc.u32 t98, %9, t81, t98;\n\t"
        "madc.hi.cc.u32 t99, %9, t81, t99;\n\t"
        "madc.lo.cc.u32 t100, %11, t81, t100;\n\t"
        "madc.hi.cc.u32 t101, %11, t81, t101;\n\t"
        "madc.lo.cc.u32 t102, %13, t81, t102;\n\t"
        "madc.hi.cc.u32 t103, %13, t81, t103;\n\t"
        "madc.lo.cc.u32 t104, %15, t81, t104;\n\t"
        "madc.hi.u32 t105, %15, t81, 0x0;\n\t"
        "add.cc.u32 t135, t115, t97;\n\t"
        "mul.lo.u32 t136, t135, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t98, 0xf3b9cac2, t136, t98;\n\t"
        "madc.hi.cc.u32 t99, 0xf3b9cac2, t136, t99;\n\t"
        "madc.lo.cc.u32 t100, 0xbce6faad, t136, t100;\n\t"
        "madc.hi.cc.u32 t101, 0xbce6faad, t136, t101;\n\t"
        "madc.lo.cc.u32 t102, 0xffffffff, t136, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xffffffff, t136, t103;\n\t"
        "madc.lo.cc.u32 t104, 0xffffffff, t136, t104;\n\t"
        "madc.hi.cc.u32 t105, 0xffffffff, t136, t105;\n\t"
        "addc.u32 t106, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t135, 0xfc632551, t136, t135;\n\t"
        "madc.hi.cc.u32 t116, 0xfc632551, t136, t116;\n\t"
        "madc.lo.cc.u32 t117, 0xa7179e84, t136, t117;\n\t"
        "madc.hi.cc.u32 t118, 0xa7179e84, t136, t118;\n\t"
        "madc.lo.cc.u32 t119, 0xffffffff, t136, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xffffffff, t136, t120;\n\t"
        "madc.lo.cc.u32 t121, 0x0, t136, t121;\n\t"
        "madc.hi.cc.u32 t122, 0x0, t136, t122;\n\t"
        "addc.u32 t123, t123, 0x0;\n\t"
        "mad.lo.cc.u32 t98, %8, t84, t98;\n\t"
        "madc.hi.cc.u32 t99, %8, t84, t99;\n\t"
        "madc.lo.cc.u32 t100, %10, t84, t100;\n\t"
        "madc.hi.cc.u32 t101, %10, t84, t101;\n\t"
        "madc.lo.cc.u32 t102, %12, t84, t102;\n\t"
        "madc.hi.cc.u32 t103, %12, t84, t103;\n\t"
        "madc.lo.cc.u32 t104, %14, t84, t104;\n\t"
        "madc.hi.cc.u32 t105, %14, t84, t105;\n\t"
        "addc.u32 t106, t106, 0x0;\n\t"
        "mad.lo.cc.u32 t117, %9, t84, t117;\n\t"
        "madc.hi.cc.u32 t118, %9, t84, t118;\n\t"
        "madc.lo.cc.u32 t119, %11, t84, t119;\n\t"
        "madc.hi.cc.u32 t120, %11, t84, t120;\n\t"
        "madc.lo.cc.u32 t121, %13, t84, t121;\n\t"
        "madc.hi.cc.u32 t122, %13, t84, t122;\n\t"
        "madc.lo.cc.u32 t123, %15, t84, t123;\n\t"
        "madc.hi.u32 t124, %15, t84, 0x0;\n\t"
        "add.cc.u32 t137, t98, t116;\n\t"
        "mul.lo.u32 t138, t137, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t117, 0xf3b9cac2, t138, t117;\n\t"
        "madc.hi.cc.u32 t118, 0xf3b9cac2, t138, t118;\n\t"
        "madc.lo.cc.u32 t119, 0xbce6faad, t138, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xbce6faad, t138, t120;\n\t"
        "madc.lo.cc.u32 t121, 0xffffffff, t138, t121;\n\t"
        "madc.hi.cc.u32 t122, 0xffffffff, t138, t122;\n\t"
        "madc.lo.cc.u32 t123, 0xffffffff, t138, t123;\n\t"
        "madc.hi.cc.u32 t124, 0xffffffff, t138, t124;\n\t"
        "addc.u32 t125, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t137, 0xfc632551, t138, t137;\n\t"
        "madc.hi.cc.u32 t99, 0xfc632551, t138, t99;\n\t"
        "madc.lo.cc.u32 t100, 0xa7179e84, t138, t100;\n\t"
        "madc.hi.cc.u32 t101, 0xa7179e84, t138, t101;\n\t"
        "madc.lo.cc.u32 t102, 0xffffffff, t138, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xffffffff, t138, t103;\n\t"
        "madc.lo.cc.u32 t104, 0x0, t138, t104;\n\t"
        "madc.hi.cc.u32 t105, 0x0, t138, t105;\n\t"
        "addc.u32 t106, t106, 0x0;\n\t"
        "mad.lo.cc.u32 t117, %8, t87, t117;\n\t"
        "madc.hi.cc.u32 t118, %8, t87, t118;\n\t"
        "madc.lo.cc.u32 t119, %10, t87, t119;\n\t"
        "madc.hi.cc.u32 t120, %10, t87, t120;\n\t"
        "madc.lo.cc.u32 t121, %12, t87, t121;\n\t"
        "madc.hi.cc.u32 t122, %12, t87, t122;\n\t"
        "madc.lo.cc.u32 t123, %14, t87, t123;\n\t"
        "madc.hi.cc.u32 t124, %14, t87, t124;\n\t"
        "addc.u32 t125, t125, 0x0;\n\t"
        "mad.lo.cc.u32 t100, %9, t87, t100;\n\t"
        "madc.hi.cc.u32 t101, %9, t87, t101;\n\t"
        "madc.lo.cc.u32 t102, %11, t87, t102;\n\t"
        "madc.hi.cc.u32 t103, %11, t87, t103;\n\t"
        "madc.lo.cc.u32 t104, %13, t87, t104;\n\t"
        "madc.hi.cc.u32 t105, %13, t87, t105;\n\t"
        "madc.lo.cc.u32 t106, %15, t87, t106;\n\t"
        "madc.hi.u32 t107, %15, t87, 0x0;\n\t"
        "add.cc.u32 t139, t117, t99;\n\t"
        "mul.lo.u32 t140, t139, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t100, 0xf3b9cac2, t140, t100;\n\t"
        "madc.hi.cc.u32 t101, 0xf3b9cac2, t140, t101;\n\t"
        "madc.lo.cc.u32 t102, 0xbce6faad, t140, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xbce6faad, t140, t103;\n\t"
        "madc.lo.cc.u32 t104, 0xffffffff, t140, t104;\n\t"
        "madc.hi.cc.u32 t105, 0xffffffff, t140, t105;\n\t"
        "madc.lo.cc.u32 t106, 0xffffffff, t140, t106;\n\t"
        "madc.hi.cc.u32 t107, 0xffffffff, t140, t107;\n\t"
        "addc.u32 t108, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t139, 0xfc632551, t140, t139;\n\t"
        "madc.hi.cc.u32 t118, 0xfc632551, t140, t118;\n\t"
        "madc.lo.cc.u32 t119, 0xa7179e84, t140, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xa7179e84, t140, t120;\n\t"
        "madc.lo.cc.u32 t121, 0xffffffff, t140, t121;\n\t"
        "madc.hi.cc.u32 t122, 0xffffffff, t140, t122;\n\t"
        "madc.lo.cc.u32 t123, 0x0, t140, t123;\n\t"
        "madc.hi.cc.u32 t124, 0x0, t140, t124;\n\t"
        "addc.u32 t125, t125, 0x0;\n\t"
        "mad.lo.cc.u32 t100, %8, t90, t100;\n\t"
        "madc.hi.cc.u32 t101, %8, t90, t101;\n\t"
        "madc.lo.cc.u32 t102, %10, t90, t102;\n\t"
        "madc.hi.cc.u32 t103, %10, t90, t103;\n\t"
        "madc.lo.cc.u32 t104, %12, t90, t104;\n\t"
        "madc.hi.cc.u32 t105, %12, t90, t105;\n\t"
        "madc.lo.cc.u32 t106, %14, t90, t106;\n\t"
        "madc.hi.cc.u32 t107, %14, t90, t107;\n\t"
        "addc.u32 t108, t108, 0x0;\n\t"
        "mad.lo.cc.u32 t119, %9, t90, t119;\n\t"
        "madc.hi.cc.u32 t120, %9, t90, t120;\n\t"
        "madc.lo.cc.u32 t121, %11, t90, t121;\n\t"
        "madc.hi.cc.u32 t122, %11, t90, t122;\n\t"
        "madc.lo.cc.u32 t123, %13, t90, t123;\n\t"
        "madc.hi.cc.u32 t124, %13, t90, t124;\n\t"
        "madc.lo.cc.u32 t125, %15, t90, t125;\n\t"
        "madc.hi.u32 t126, %15, t90, 0x0;\n\t"
        "add.cc.u32 t141, t100, t118;\n\t"
        "mul.lo.u32 t142, t141, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t119, 0xf3b9cac2, t142, t119;\n\t"
        "madc.hi.cc.u32 t120, 0xf3b9cac2, t142, t120;\n\t"
        "madc.lo.cc.u32 t121, 0xbce6faad, t142, t121;\n\t"
        "madc.hi.cc.u32 t122, 0xbce6faad, t142, t122;\n\t"
        "madc.lo.cc.u32 t123, 0xffffffff, t142, t123;\n\t"
        "madc.hi.cc.u32 t124, 0xffffffff, t142, t124;\n\t"
        "madc.lo.cc.u32 t125, 0xffffffff, t142, t125;\n\t"
        "madc.hi.cc.u32 t126, 0xffffffff, t142, t126;\n\t"
        "addc.u32 t127, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t141, 0xfc632551, t142, t141;\n\t"
        "madc.hi.cc.u32 t101, 0xfc632551, t142, t101;\n\t"
        "madc.lo.cc.u32 t102, 0xa7179e84, t142, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xa7179e84, t142, t103;\n\t"
        "madc.lo.cc.u32 t104, 0xffffffff, t142, t104;\n\t"
        "madc.hi.cc.u32 t105, 0xffffffff, t142, t105;\n\t"
        "madc.lo.cc.u32 t106, 0x0, t142, t106;\n\t"
        "madc.hi.cc.u32 t107, 0x0, t142, t107;\n\t"
        "addc.u32 t108, t108, 0x0;\n\t"
        "mad.lo.cc.u32 t119, %8, t93, t119;\n\t"
        "madc.hi.cc.u32 t120, %8, t93, t120;\n\t"
        "madc.lo.cc.u32 t121, %10, t93, t121;\n\t"
        "madc.hi.cc.u32 t122, %10, t93, t122;\n\t"
        "madc.lo.cc.u32 t123, %12, t93, t123;\n\t"
        "madc.hi.cc.u32 t124, %12, t93, t124;\n\t"
        "madc.lo.cc.u32 t125, %14, t93, t125;\n\t"
        "madc.hi.cc.u32 t126, %14, t93, t126;\n\t"
        "addc.u32 t127, t127, 0x0;\n\t"
        "mad.lo.cc.u32 t102, %9, t93, t102;\n\t"
        "madc.hi.cc.u32 t103, %9, t93, t103;\n\t"
        "madc.lo.cc.u32 t104, %11, t93, t104;\n\t"
        "madc.hi.cc.u32 t105, %11, t93, t105;\n\t"
        "madc.lo.cc.u32 t106, %13, t93, t106;\n\t"
        "madc.hi.cc.u32 t107, %13, t93, t107;\n\t"
        "madc.lo.cc.u32 t108, %15, t93, t108;\n\t"
        "madc.hi.u32 t109, %15, t93, 0x0;\n\t"
        "add.cc.u32 t143, t119, t101;\n\t"
        "mul.lo.u32 t144, t143, 0xee00bc4f;\n\t"
        "madc.lo.cc.u32 t102, 0xf3b9cac2, t144, t102;\n\t"
        "madc.hi.cc.u32 t103, 0xf3b9cac2, t144, t103;\n\t"
        "madc.lo.cc.u32 t104, 0xbce6faad, t144, t104;\n\t"
        "madc.hi.cc.u32 t105, 0xbce6faad, t144, t105;\n\t"
        "madc.lo.cc.u32 t106, 0xffffffff, t144, t106;\n\t"
        "madc.hi.cc.u32 t107, 0xffffffff, t144, t107;\n\t"
        "madc.lo.cc.u32 t108, 0xffffffff, t144, t108;\n\t"
        "madc.hi.cc.u32 t109, 0xffffffff, t144, t109;\n\t"
        "addc.u32 t110, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t143, 0xfc632551, t144, t143;\n\t"
        "madc.hi.cc.u32 t120, 0xfc632551, t144, t120;\n\t"
        "madc.lo.cc.u32 t121, 0xa7179e84, t144, t121;\n\t"
        "madc.hi.cc.u32 t122, 0xa7179e84, t144, t122;\n\t"
        "madc.lo.cc.u32 t123, 0xffffffff, t144, t123;\n\t"
        "madc.hi.cc.u32 t124, 0xffffffff, t144, t124;\n\t"
        "madc.lo.cc.u32 t125, 0x0, t144, t125;\n\t"
        "madc.hi.cc.u32 t126, 0x0, t144, t126;\n\t"
        "addc.u32 t127, t127, 0x0;\n\t"
        "add.cc.u32 t145, t102, t120;\n\t"
        "addc.cc.u32 t146, t103, t121;\n\t"
        "addc.cc.u32 t147, t104, t122;\n\t"
        "addc.cc.u32 t148, t105, t123;\n\t"
        "addc.cc.u32 t149, t106, t124;\n\t"
        "addc.cc.u32 t150, t107, t125;\n\t"
        "addc.cc.u32 t151, t108, t126;\n\t"
        "addc.cc.u32 t152, t109, t127;\n\t"
        "addc.u32 t153, t110, 0x0;\n\t"
        "sub.cc.u32 t154, t145, 0xfc632551;\n\t"
        "subc.cc.u32 t155, t146, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t156, t147, 0xa7179e84;\n\t"
        "subc.cc.u32 t157, t148, 0xbce6faad;\n\t"
        "subc.cc.u32 t158, t149, 0xffffffff;\n\t"
        "subc.cc.u32 t159, t150, 0xffffffff;\n\t"
        "subc.cc.u32 t160, t151, 0x0;\n\t"
        "subc.cc.u32 t161, t152, 0xffffffff;\n\t"
        "subc.cc.u32 t162, t153, 0x0;\n\t"
        "subc.u32 t163, 0x0, 0x0;\n\t"
        "xor.b32 t164, t154, t145;\n\t"
        "and.b32 t165, t164, t163;\n\t"
        "xor.b32 t166, t165, t154;\n\t"
        "xor.b32 t167, t155, t146;\n\t"
        "and.b32 t168, t167, t163;\n\t"
        "xor.b32 t169, t168, t155;\n\t"
        "xor.b32 t170, t156, t147;\n\t"
        "and.b32 t171, t170, t163;\n\t"
        "xor.b32 t172, t171, t156;\n\t"
        "xor.b32 t173, t157, t148;\n\t"
        "and.b32 t174, t173, t163;\n\t"
        "xor.b32 t175, t174, t157;\n\t"
        "xor.b32 t176, t158, t149;\n\t"
        "and.b32 t177, t176, t163;\n\t"
        "xor.b32 t178, t177, t158;\n\t"
        "xor.b32 t179, t159, t150;\n\t"
        "and.b32 t180, t179, t163;\n\t"
        "xor.b32 t181, t180, t159;\n\t"
        "xor.b32 t182, t160, t151;\n\t"
        "and.b32 t183, t182, t163;\n\t"
        "xor.b32 t184, t183, t160;\n\t"
        "xor.b32 t185, t161, t152;\n\t"
        "and.b32 t186, t185, t163;\n\t"
        "xor.b32 t187, t186, t161;\n\t"
        "add.cc.u32 t188, t166, %16;\n\t"
        "addc.cc.u32 t189, t169, %17;\n\t"
        "addc.cc.u32 t190, t172, %18;\n\t"
        "addc.cc.u32 t191, t175, %19;\n\t"
        "addc.cc.u32 t192, t178, %20;\n\t"
        "addc.cc.u32 t193, t181, %21;\n\t"
        "addc.cc.u32 t194, t184, %22;\n\t"
        "addc.cc.u32 t195, t187, %23;\n\t"
        "addc.u32 t196, 0x0, 0x0;\n\t"
        "sub.cc.u32 t197, t188, 0xfc632551;\n\t"
        "subc.cc.u32 t198, t189, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t199, t190, 0xa7179e84;\n\t"
        "subc.cc.u32 t200, t191, 0xbce6faad;\n\t"
        "subc.cc.u32 t201, t192, 0xffffffff;\n\t"
        "subc.cc.u32 t202, t193, 0xffffffff;\n\t"
        "subc.cc.u32 t203, t194, 0x0;\n\t"
        "subc.cc.u32 t204, t195, 0xffffffff;\n\t"
        "subc.cc.u32 t205, t196, 0x0;\n\t"
        "subc.u32 t206, 0x0, 0x0;\n\t"
        "xor.b32 t207, t197, t188;\n\t"
        "and.b32 t208, t207, t206;\n\t"
        "xor.b32 t209, t208, t197;\n\t"
        "xor.b32 t210, t198, t189;\n\t"
        "and.b32 t211, t210, t206;\n\t"
        "xor.b32 t212, t211, t198;\n\t"
        "xor.b32 t213, t199, t190;\n\t"
        "and.b32 t214, t213, t206;\n\t"
        "xor.b32 t215, t214, t199;\n\t"
        "xor.b32 t216, t200, t191;\n\t"
        "and.b32 t217, t216, t206;\n\t"
        "xor.b32 t218, t217, t200;\n\t"
        "xor.b32 t219, t201, t192;\n\t"
        "and.b32 t220, t219, t206;\n\t"
        "xor.b32 t221, t220, t201;\n\t"
        "xor.b32 t222, t202, t193;\n\t"
        "and.b32 t223, t222, t206;\n\t"
        "xor.b32 t224, t223, t202;\n\t"
        "xor.b32 t225, t203, t194;\n\t"
        "and.b32 t226, t225, t206;\n\t"
        "xor.b32 t227, t226, t203;\n\t"
        "xor.b32 t228, t204, t195;\n\t"
        "and.b32 t229, t228, t206;\n\t"
        "xor.b32 t230, t229, t204;\n\t"
        "mov.u32 %0, t209;\n\t"
        "mov.u32 %1, t212;\n\t"
        "mov.u32 %2, t215;\n\t"
        "mov.u32 %3, t218;\n\t"
        "mov.u32 %4, t221;\n\t"
        "mov.u32 %5, t224;\n\t"
        "mov.u32 %6, t227;\n\t"
        "mov.u32 %7, t230;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127, t128, t129, t130, t131, t132, t133, t134, t135, t136, t137, t138, t139, t140, t141, t142, t143, t144, t145, t146, t147, t148, t149, t150, t151, t152, t153, t154, t155, t156, t157, t158, t159, t160, t161, t162, t163, t164, t165, t166, t167, t168, t169, t170, t171, t172, t173, t174, t175, t176, t177, t178, t179, t180, t181, t182, t183, t184, t185, t186, t187, t188, t189, t190, t191, t192, t193, t194, t195, t196, t197, t198, t199, t200, t201, t202, t203, t204, t205, t206, t207, t208, t209, t210, t211, t212, t213, t214, t215, t216, t217, t218, t219, t220, t221, t222, t223, t224, t225, t226, t227, t228, t229, t230;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t0 = (uint32_t)((uint32_t)(0xbe79eea2u * b_i));
    t1 = (uint32_t)(((uint64_t)0xbe79eea2u * b_i) >> 32);
    t2 = (uint32_t)((uint32_t)(0x49bd6fa6u * b_i));
    t3 = (uint32_t)(((uint64_t)0x49bd6fa6u * b_i) >> 32);
    t4 = (uint32_t)((uint32_t)(0x2b6bec59u * b_i));
    t5 = (uint32_t)(((uint64_t)0x2b6bec59u * b_i) >> 32);
    t6 = (uint32_t)((uint32_t)(0xf3d95620u * b_i));
    t7 = (uint32_t)(((uint64_t)0xf3d95620u * b_i) >> 32);
    t19 = (uint32_t)((uint32_t)(0x83244c95u * b_i));
    t20 = (uint32_t)(((uint64_t)0x83244c95u * b_i) >> 32);
    t21 = (uint32_t)((uint32_t)(0x4699799cu * b_i));
    t22 = (uint32_t)(((uint64_t)0x4699799cu * b_i) >> 32);
    t23 = (uint32_t)((uint32_t)(0x2845b239u * b_i));
    t24 = (uint32_t)(((uint64_t)0x2845b239u * b_i) >> 32);
    t25 = (uint32_t)((uint32_t)(0x66e12d94u * b_i));
    t26 = (uint32_t)(((uint64_t)0x66e12d94u * b_i) >> 32);
    t36 = (uint32_t)((uint32_t)(t0 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t36) + t19; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t36) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t36) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t36) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t36) + t0; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t36) >> 32) + t1 + cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t36) + t2 + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t36) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t36) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t36) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t36) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t36) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t8 = (uint32_t)w_;
    w_ = (uint64_t)t19 + t1; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t38 = (uint32_t)((uint32_t)(t37 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t38) + t2 + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t38) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t38) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t38) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + 0x0u + cf_; t9 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t38) + t37; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t38) >> 32) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t38) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t38) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t38) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t38) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t38) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t38) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t27 + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)t2 + t20; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t40 = (uint32_t)((uint32_t)(t39 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t40) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t40) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t40) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t40) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + 0x0u + cf_; t28 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t40) + t39; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t40) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t40) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t40) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t40) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t40) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t40) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t40) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t10 = (uint32_t)w_;
    w_ = (uint64_t)t21 + t3; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t42 = (uint32_t)((uint32_t)(t41 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t42) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t42) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t42) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t42) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + 0x0u + cf_; t11 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t42) + t41; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t42) >> 32) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t42) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t42) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t42) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t42) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t42) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t42) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)t4 + t22; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t44 = (uint32_t)((uint32_t)(t43 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t44) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t44) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t44) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t44) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + 0x0u + cf_; t30 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t44) + t43; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t44) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t44) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t44) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t44) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t44) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t44) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t44) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t12 = (uint32_t)w_;
    w_ = (uint64_t)t23 + t5; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t46 = (uint32_t)((uint32_t)(t45 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t46) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t46) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t46) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t46) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + 0x0u + cf_; t13 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t46) + t45; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t46) >> 32) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t46) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t46) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t46) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t46) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t46) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t46) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t31 = (uint32_t)w_;
    w_ = (uint64_t)t6 + t24; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t48 = (uint32_t)((uint32_t)(t47 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t48) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t48) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t48) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t48) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t48) + t47; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t48) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t48) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t48) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t48) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t48) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t48) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t48) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t14 = (uint32_t)w_;
    w_ = (uint64_t)t25 + t7; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t50 = (uint32_t)((uint32_t)(t49 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t50) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t50) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t50) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t50) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + 0x0u + cf_; t15 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t50) + t49; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t50) >> 32) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t50) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t50) >> 32) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t50) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t50) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t50) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t50) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t33 = (uint32_t)w_;
    w_ = (uint64_t)t8 + t26; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t27 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t28 + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t29 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t30 + cf_; t55 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t31 + cf_; t56 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + t32 + cf_; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + t33 + cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t59 = (uint32_t)w_;
    w_ = (uint64_t)t51 - 0xfc632551u; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t52 - 0xf3b9cac2u - cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t53 - 0xa7179e84u - cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t54 - 0xbce6faadu - cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t55 - 0xffffffffu - cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t56 - 0xffffffffu - cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t57 - 0x0u - cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t58 - 0xffffffffu - cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t59 - 0x0u - cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t69 = (uint32_t)w_;
    t70 = (uint32_t)(t60 ^ t51);
    t71 = (uint32_t)(t70 & t69);
    t72 = (uint32_t)(t71 ^ t60);
    t73 = (uint32_t)(t61 ^ t52);
    t74 = (uint32_t)(t73 & t69);
    t75 = (uint32_t)(t74 ^ t61);
    t76 = (uint32_t)(t62 ^ t53);
    t77 = (uint32_t)(t76 & t69);
    t78 = (uint32_t)(t77 ^ t62);
    t79 = (uint32_t)(t63 ^ t54);
    t80 = (uint32_t)(t79 & t69);
    t81 = (uint32_t)(t80 ^ t63);
    t82 = (uint32_t)(t64 ^ t55);
    t83 = (uint32_t)(t82 & t69);
    t84 = (uint32_t)(t83 ^ t64);
    t85 = (uint32_t)(t65 ^ t56);
    t86 = (uint32_t)(t85 & t69);
    t87 = (uint32_t)(t86 ^ t65);
    t88 = (uint32_t)(t66 ^ t57);
    t89 = (uint32_t)(t88 & t69);
    t90 = (uint32_t)(t89 ^ t66);
    t91 = (uint32_t)(t67 ^ t58);
    t92 = (uint32_t)(t91 & t69);
    t93 = (uint32_t)(t92 ^ t67);
    t94 = (uint32_t)((uint32_t)(a_0_i * t72));
    t95 = (uint32_t)(((uint64_t)a_0_i * t72) >> 32);
    t96 = (uint32_t)((uint32_t)(a_2_i * t72));
    t97 = (uint32_t)(((uint64_t)a_2_i * t72) >> 32);
    t98 = (uint32_t)((uint32_t)(a_4_i * t72));
    t99 = (uint32_t)(((uint64_t)a_4_i * t72) >> 32);
    t100 = (uint32_t)((uint32_t)(a_6_i * t72));
    t101 = (uint32_t)(((uint64_t)a_6_i * t72) >> 32);
    t113 = (uint32_t)((uint32_t)(a_1_i * t72));
    t114 = (uint32_t)(((uint64_t)a_1_i * t72) >> 32);
    t115 = (uint32_t)((uint32_t)(a_3_i * t72));
    t116 = (uint32_t)(((uint64_t)a_3_i * t72) >> 32);
    t117 = (uint32_t)((uint32_t)(a_5_i * t72));
    t118 = (uint32_t)(((uint64_t)a_5_i * t72) >> 32);
    t119 = (uint32_t)((uint32_t)(a_7_i * t72));
    t120 = (uint32_t)(((uint64_t)a_7_i * t72) >> 32);
    t130 = (uint32_t)((uint32_t)(t94 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t130) + t113; t113 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t130) >> 32) + t114 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t130) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t130) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t130) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t130) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t130) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t130) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t121 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t130) + t94; t94 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t130) >> 32) + t95 + cf_; t95 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t130) + t96 + cf_; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t130) >> 32) + t97 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t130) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t130) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t130) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t130) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t102 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t75) + t113; t113 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t75) >> 32) + t114 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t75) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t75) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t75) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t75) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t75) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t75) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t121 + 0x0u + cf_; t121 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t75) + t96; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t75) >> 32) + t97 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t75) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t75) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t75) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t75) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t75) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t75) >> 32) + 0x0u + cf_; t103 = (uint32_t)w_;
    w_ = (uint64_t)t113 + t95; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t132 = (uint32_t)((uint32_t)(t131 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t132) + t96 + cf_; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t132) >> 32) + t97 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t132) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t132) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t132) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t132) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t132) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t132) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t104 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t132) + t131; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t132) >> 32) + t114 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t132) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t132) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t132) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t132) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t132) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t132) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t121 + 0x0u + cf_; t121 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t78) + t96; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t78) >> 32) + t97 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t78) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t78) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t78) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t78) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t78) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t78) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t104 + 0x0u + cf_; t104 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t78) + t115; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t78) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t78) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t78) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t78) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t78) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t78) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t78) >> 32) + 0x0u + cf_; t122 = (uint32_t)w_;
    w_ = (uint64_t)t96 + t114; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t134 = (uint32_t)((uint32_t)(t133 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t134) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t134) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t134) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t134) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t134) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t134) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t134) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t134) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t123 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t134) + t133; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t134) >> 32) + t97 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t134) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t134) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t134) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t134) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t134) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t134) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t104 + 0x0u + cf_; t104 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t81) + t115; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t81) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t81) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t81) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t81) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t81) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t81) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t81) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t123 + 0x0u + cf_; t123 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t81) + t98; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t81) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t81) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t81) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t81) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t81) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t81) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t81) >> 32) + 0x0u + cf_; t105 = (uint32_t)w_;
    w_ = (uint64_t)t115 + t97; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t136 = (uint32_t)((uint32_t)(t135 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t136) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t136) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t136) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t136) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t136) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t136) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t136) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t136) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t106 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t136) + t135; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t136) >> 32) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t136) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t136) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t136) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t136) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t136) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t136) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t123 + 0x0u + cf_; t123 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t84) + t98; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t84) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t84) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t84) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t84) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t84) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t84) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t84) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t106 + 0x0u + cf_; t106 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t84) + t117; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t84) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t84) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t84) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t84) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t84) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t84) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t84) >> 32) + 0x0u + cf_; t124 = (uint32_t)w_;
    w_ = (uint64_t)t98 + t116; t137 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t138 = (uint32_t)((uint32_t)(t137 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t138) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t138) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t138) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t138) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t138) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t138) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t138) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t138) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t125 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t138) + t137; t137 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t138) >> 32) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t138) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t138) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t138) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t138) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t138) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t138) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t106 + 0x0u + cf_; t106 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t87) + t117; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t87) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t87) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t87) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t87) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t87) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t87) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t87) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t125 + 0x0u + cf_; t125 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t87) + t100; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t87) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t87) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t87) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t87) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t87) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t87) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t87) >> 32) + 0x0u + cf_; t107 = (uint32_t)w_;
    w_ = (uint64_t)t117 + t99; t139 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t140 = (uint32_t)((uint32_t)(t139 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t140) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t140) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t140) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t140) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t140) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t140) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t140) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t140) >> 32) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t108 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t140) + t139; t139 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t140) >> 32) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t140) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t140) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t140) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t140) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t140) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t140) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t125 + 0x0u + cf_; t125 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t90) + t100; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t90) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t90) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t90) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t90) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t90) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t90) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t90) >> 32) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t108 + 0x0u + cf_; t108 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t90) + t119; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t90) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t90) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t90) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t90) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t90) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t90) + t125 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t90) >> 32) + 0x0u + cf_; t126 = (uint32_t)w_;
    w_ = (uint64_t)t100 + t118; t141 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t142 = (uint32_t)((uint32_t)(t141 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t142) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t142) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t142) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t142) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t142) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t142) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t142) + t125 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t142) >> 32) + t126 + cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t127 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t142) + t141; t141 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t142) >> 32) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t142) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t142) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t142) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t142) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t142) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t142) >> 32) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t108 + 0x0u + cf_; t108 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t93) + t119; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t93) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t93) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t93) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t93) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t93) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t93) + t125 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t93) >> 32) + t126 + cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t127 + 0x0u + cf_; t127 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t93) + t102; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t93) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t93) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t93) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t93) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t93) >> 32) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t93) + t108 + cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t93) >> 32) + 0x0u + cf_; t109 = (uint32_t)w_;
    w_ = (uint64_t)t119 + t101; t143 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    t144 = (uint32_t)((uint32_t)(t143 * 0xee00bc4fu));
    w_ = (uint64_t)(uint32_t)(0xf3b9cac2u * t144) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xf3b9cac2u * t144) >> 32) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xbce6faadu * t144) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xbce6faadu * t144) >> 32) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t144) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t144) >> 32) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t144) + t108 + cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t144) >> 32) + t109 + cf_; t109 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t110 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(0xfc632551u * t144) + t143; t143 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xfc632551u * t144) >> 32) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xa7179e84u * t144) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xa7179e84u * t144) >> 32) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0xffffffffu * t144) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0xffffffffu * t144) >> 32) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(0x0u * t144) + t125 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)0x0u * t144) >> 32) + t126 + cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t127 + 0x0u + cf_; t127 = (uint32_t)w_;
    w_ = (uint64_t)t102 + t120; t145 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t103 + t121 + cf_; t146 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t104 + t122 + cf_; t147 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t105 + t123 + cf_; t148 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t106 + t124 + cf_; t149 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t107 + t125 + cf_; t150 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t108 + t126 + cf_; t151 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t109 + t127 + cf_; t152 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t110 + 0x0u + cf_; t153 = (uint32_t)w_;
    w_ = (uint64_t)t145 - 0xfc632551u; t154 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t146 - 0xf3b9cac2u - cf_; t155 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t147 - 0xa7179e84u - cf_; t156 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t148 - 0xbce6faadu - cf_; t157 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t149 - 0xffffffffu - cf_; t158 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t150 - 0xffffffffu - cf_; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t151 - 0x0u - cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t152 - 0xffffffffu - cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t153 - 0x0u - cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t163 = (uint32_t)w_;
    t164 = (uint32_t)(t154 ^ t145);
    t165 = (uint32_t)(t164 & t163);
    t166 = (uint32_t)(t165 ^ t154);
    t167 = (uint32_t)(t155 ^ t146);
    t168 = (uint32_t)(t167 & t163);
    t169 = (uint32_t)(t168 ^ t155);
    t170 = (uint32_t)(t156 ^ t147);
    t171 = (uint32_t)(t170 & t163);
    t172 = (uint32_t)(t171 ^ t156);
    t173 = (uint32_t)(t157 ^ t148);
    t174 = (uint32_t)(t173 & t163);
    t175 = (uint32_t)(t174 ^ t157);
    t176 = (uint32_t)(t158 ^ t149);
    t177 = (uint32_t)(t176 & t163);
    t178 = (uint32_t)(t177 ^ t158);
    t179 = (uint32_t)(t159 ^ t150);
    t180 = (uint32_t)(t179 & t163);
    t181 = (uint32_t)(t180 ^ t159);
    t182 = (uint32_t)(t160 ^ t151);
    t183 = (uint32_t)(t182 & t163);
    t184 = (uint32_t)(t183 ^ t160);
    t185 = (uint32_t)(t161 ^ t152);
    t186 = (uint32_t)(t185 & t163);
    t187 = (uint32_t)(t186 ^ t161);
    w_ = (uint64_t)t166 + c_0_i; t188 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t169 + c_1_i + cf_; t189 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t172 + c_2_i + cf_; t190 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t175 + c_3_i + cf_; t191 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t178 + c_4_i + cf_; t192 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t181 + c_5_i + cf_; t193 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t184 + c_6_i + cf_; t194 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t187 + c_7_i + cf_; t195 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t196 = (uint32_t)w_;
    w_ = (uint64_t)t188 - 0xfc632551u; t197 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t189 - 0xf3b9cac2u - cf_; t198 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t190 - 0xa7179e84u - cf_; t199 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t191 - 0xbce6faadu - cf_; t200 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t192 - 0xffffffffu - cf_; t201 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t193 - 0xffffffffu - cf_; t202 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t194 - 0x0u - cf_; t203 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t195 - 0xffffffffu - cf_; t204 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t196 - 0x0u - cf_; t205 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t206 = (uint32_t)w_;
    t207 = (uint32_t)(t197 ^ t188);
    t208 = (uint32_t)(t207 & t206);
    t209 = (uint32_t)(t208 ^ t197);
    t210 = (uint32_t)(t198 ^ t189);
    t211 = (uint32_t)(t210 & t206);
    t212 = (uint32_t)(t211 ^ t198);
    t213 = (uint32_t)(t199 ^ t190);
    t214 = (uint32_t)(t213 & t206);
    t215 = (uint32_t)(t214 ^ t199);
    t216 = (uint32_t)(t200 ^ t191);
    t217 = (uint32_t)(t216 & t206);
    t218 = (uint32_t)(t217 ^ t200);
    t219 = (uint32_t)(t201 ^ t192);
    t220 = (uint32_t)(t219 & t206);
    t221 = (uint32_t)(t220 ^ t201);
    t222 = (uint32_t)(t202 ^ t193);
    t223 = (uint32_t)(t222 & t206);
    t224 = (uint32_t)(t223 ^ t202);
    t225 = (uint32_t)(t203 ^ t194);
    t226 = (uint32_t)(t225 & t206);
    t227 = (uint32_t)(t226 ^ t203);
    t228 = (uint32_t)(t204 ^ t195);
    t229 = (uint32_t)(t228 & t206);
    t230 = (uint32_t)(t229 ^ t204);
    r[0] = t209;
    r[1] = t212;
    r[2] = t215;
    r[3] = t218;
    r[4] = t221;
    r[5] = t224;
    r[6] = t227;
    r[7] = t230;
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
        "addc.u32 t8, 0x0, 0x0;\n\t"
        "sub.cc.u32 t9, t0, 0xfc632551;\n\t"
        "subc.cc.u32 t10, t1, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t11, t2, 0xa7179e84;\n\t"
        "subc.cc.u32 t12, t3, 0xbce6faad;\n\t"
        "subc.cc.u32 t13, t4, 0xffffffff;\n\t"
        "subc.cc.u32 t14, t5, 0xffffffff;\n\t"
        "subc.cc.u32 t15, t6, 0x0;\n\t"
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
    w_ = (uint64_t)t0 - 0xfc632551u; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t1 - 0xf3b9cac2u - cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t2 - 0xa7179e84u - cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t3 - 0xbce6faadu - cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t4 - 0xffffffffu - cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t5 - 0xffffffffu - cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t6 - 0x0u - cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
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
        ".reg .u32 t<21>;\n\t"
        "sub.cc.u32 t0, %8, %16;\n\t"
        "subc.cc.u32 t1, %9, %17;\n\t"
        "subc.cc.u32 t2, %10, %18;\n\t"
        "subc.cc.u32 t3, %11, %19;\n\t"
        "subc.cc.u32 t4, %12, %20;\n\t"
        "subc.cc.u32 t5, %13, %21;\n\t"
        "subc.cc.u32 t6, %14, %22;\n\t"
        "subc.cc.u32 t7, %15, %23;\n\t"
        "subc.u32 t8, 0x0, 0x0;\n\t"
        "and.b32 t9, t8, 0xfc632551;\n\t"
        "and.b32 t10, t8, 0xf3b9cac2;\n\t"
        "and.b32 t11, t8, 0xa7179e84;\n\t"
        "and.b32 t12, t8, 0xbce6faad;\n\t"
        "add.cc.u32 t13, t0, t9;\n\t"
        "addc.cc.u32 t14, t1, t10;\n\t"
        "addc.cc.u32 t15, t2, t11;\n\t"
        "addc.cc.u32 t16, t3, t12;\n\t"
        "addc.cc.u32 t17, t4, t8;\n\t"
        "addc.cc.u32 t18, t5, t8;\n\t"
        "addc.cc.u32 t19, t6, 0x0;\n\t"
        "addc.u32 t20, t7, t8;\n\t"
        "mov.u32 %0, t13;\n\t"
        "mov.u32 %1, t14;\n\t"
        "mov.u32 %2, t15;\n\t"
        "mov.u32 %3, t16;\n\t"
        "mov.u32 %4, t17;\n\t"
        "mov.u32 %5, t18;\n\t"
        "mov.u32 %6, t19;\n\t"
        "mov.u32 %7, t20;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20;
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
    t9 = (uint32_t)(t8 & 0xfc632551u);
    t10 = (uint32_t)(t8 & 0xf3b9cac2u);
    t11 = (uint32_t)(t8 & 0xa7179e84u);
    t12 = (uint32_t)(t8 & 0xbce6faadu);
    w_ = (uint64_t)t0 + t9; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t1 + t10 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t11 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t12 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t8 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t8 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + 0x0u + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t8 + cf_; t20 = (uint32_t)w_;
    r[0] = t13;
    r[1] = t14;
    r[2] = t15;
    r[3] = t16;
    r[4] = t17;
    r[5] = t18;
    r[6] = t19;
    r[7] = t20;
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
        ".reg .u32 t<21>;\n\t"
        "sub.cc.u32 t0, 0x0, %8;\n\t"
        "subc.cc.u32 t1, 0x0, %9;\n\t"
        "subc.cc.u32 t2, 0x0, %10;\n\t"
        "subc.cc.u32 t3, 0x0, %11;\n\t"
        "subc.cc.u32 t4, 0x0, %12;\n\t"
        "subc.cc.u32 t5, 0x0, %13;\n\t"
        "subc.cc.u32 t6, 0x0, %14;\n\t"
        "subc.cc.u32 t7, 0x0, %15;\n\t"
        "subc.u32 t8, 0x0, 0x0;\n\t"
        "and.b32 t9, t8, 0xfc632551;\n\t"
        "and.b32 t10, t8, 0xf3b9cac2;\n\t"
        "and.b32 t11, t8, 0xa7179e84;\n\t"
        "and.b32 t12, t8, 0xbce6faad;\n\t"
        "add.cc.u32 t13, t0, t9;\n\t"
        "addc.cc.u32 t14, t1, t10;\n\t"
        "addc.cc.u32 t15, t2, t11;\n\t"
        "addc.cc.u32 t16, t3, t12;\n\t"
        "addc.cc.u32 t17, t4, t8;\n\t"
        "addc.cc.u32 t18, t5, t8;\n\t"
        "addc.cc.u32 t19, t6, 0x0;\n\t"
        "addc.u32 t20, t7, t8;\n\t"
        "mov.u32 %0, t13;\n\t"
        "mov.u32 %1, t14;\n\t"
        "mov.u32 %2, t15;\n\t"
        "mov.u32 %3, t16;\n\t"
        "mov.u32 %4, t17;\n\t"
        "mov.u32 %5, t18;\n\t"
        "mov.u32 %6, t19;\n\t"
        "mov.u32 %7, t20;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20;
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
    t9 = (uint32_t)(t8 & 0xfc632551u);
    t10 = (uint32_t)(t8 & 0xf3b9cac2u);
    t11 = (uint32_t)(t8 & 0xa7179e84u);
    t12 = (uint32_t)(t8 & 0xbce6faadu);
    w_ = (uint64_t)t0 + t9; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t1 + t10 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t11 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t12 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t8 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t8 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + 0x0u + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t8 + cf_; t20 = (uint32_t)w_;
    r[0] = t13;
    r[1] = t14;
    r[2] = t15;
    r[3] = t16;
    r[4] = t17;
    r[5] = t18;
    r[6] = t19;
    r[7] = t20;
#endif
  }

  // canonical residue of a stored value; returns 1 iff it was already < p
  // (flatten/modfsb, pseudo.py:255-283)
  static MAB_DEV uint32_t canon(uint32_t (&r)[8], const uint32_t (&a)[8]) {
    uint32_t lt;
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<34>;\n\t"
        "sub.cc.u32 t1, %9, 0xfc632551;\n\t"
        "subc.cc.u32 t2, %10, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t3, %11, 0xa7179e84;\n\t"
        "subc.cc.u32 t4, %12, 0xbce6faad;\n\t"
        "subc.cc.u32 t5, %13, 0xffffffff;\n\t"
        "subc.cc.u32 t6, %14, 0xffffffff;\n\t"
        "subc.cc.u32 t7, %15, 0x0;\n\t"
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
    w_ = (uint64_t)a_0_i - 0xfc632551u; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_1_i - 0xf3b9cac2u - cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_2_i - 0xa7179e84u - cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_3_i - 0xbce6faadu - cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_4_i - 0xffffffffu - cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_5_i - 0xffffffffu - cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_6_i - 0x0u - cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
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

  static MAB_DEV void set_p(uint32_t (&r)[8]) { r[0] = 0xfc632551u; r[1] = 0xf3b9cac2u; r[2] = 0xa7179e84u; r[3] = 0xbce6faadu; r[4] = 0xffffffffu; r[5] = 0xffffffffu; r[6] = 0x00000000u; r[7] = 0xffffffffu; }
  static MAB_DEV void set_one(uint32_t (&r)[8]) { r[0] = 0x039cdaafu; r[1] = 0x0c46353du; r[2] = 0x58e8617bu; r[3] = 0x43190552u; r[4] = 0x00000000u; r[5] = 0x00000000u; r[6] = 0xffffffffu; r[7] = 0x00000000u; }
  static MAB_DEV void set_roi(uint32_t (&r)[8]) { r[0] = 0x7e368fe1u; r[1] = 0x1015708fu; r[2] = 0x6ecc4511u; r[3] = 0x31c6c545u; r[4] = 0x98a19ea1u; r[5] = 0x5281fe89u; r[6] = 0x10c63fe8u; r[7] = 0x0279089eu; }
  static MAB_DEV void set_r2(uint32_t (&r)[8]) { r[0] = 0xbe79eea2u; r[1] = 0x83244c95u; r[2] = 0x49bd6fa6u; r[3] = 0x4699799cu; r[4] = 0x2b6bec59u; r[5] = 0x2845b239u; r[6] = 0xf3d95620u; r[7] = 0x66e12d94u; }
  static constexpr bool HAS_WEIERSTRASS = false;

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
    sqr_w(t2, t3);
    MAB_NOUNROLL
    for (int i = 1; i < 16; i++) sqr_w(t2, t2);
    mul_w(t2, t2, t3);
    sqr_w(z, t2);
    MAB_NOUNROLL
    for (int i = 1; i < 64; i++) sqr_w(z, z);
    mul_w(z, z, t2);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 32; i++) sqr_w(z, z);
    mul_w(z, z, t2);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr_w(z, z);
    mul_w(z, z, t1);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr_w(z, z);
    mul_w(z, z, t1);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr_w(z, z);
    mul_w(z, z, t1);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 6; i++) sqr_w(z, z);
    mul_w(z, z, t1);
    sqr_w(z, z);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 6; i++) sqr_w(z, z);
    mul_w(z, z, t1);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr_w(z, z);
    mul_w(z, z, t1);
    sqr_w(z, z);
    sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    sqr_w(z, z);
    mul_w(z, z, x);
    sqr_w(z, z);
    if (WEAK) (void)canon(z, z);
  }
};
