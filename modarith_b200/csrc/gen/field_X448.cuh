// Automatically generated field arithmetic for sm_100a -- do not edit.
// Command line : python -m modarith_b200.gen.monty_sm100 X448
// modulus X448 = 0xfffffffffffffffffffffffffffffffffffffffffffffffffffffffeffffffffffffffffffffffffffffffffffffffffffffffffffffffff
// plan GenMersenne: 14 saturated 32-bit limbs; stored values < 2^448; R = 2^0
//   mul   : 196 IMAD.WIDE   0 IMAD  ~101 ALU-pipe ops
//   sqr   : 105 IMAD.WIDE   0 IMAD  ~125 ALU-pipe ops
//   mli   :  14 IMAD.WIDE   0 IMAD  ~ 38 ALU-pipe ops
//   mla   :  14 IMAD.WIDE   0 IMAD  ~ 39 ALU-pipe ops
//   add   :   0 IMAD.WIDE   0 IMAD  ~ 38 ALU-pipe ops
//   sub   :   0 IMAD.WIDE   0 IMAD  ~ 46 ALU-pipe ops
//   canon :   0 IMAD.WIDE   0 IMAD  ~ 58 ALU-pipe ops
//   modpro: 445 squarings + 14 multiplies (exponent (p-1-2^k)/2^(k+1), k=1)
#pragma once
#include "mab_common.cuh"

struct F_X448 {
  static constexpr int L = 14;
  static constexpr int NBITS = 448;
  static constexpr int NBYTES = 56;
  static constexpr int PM1D2 = 1;
  static constexpr bool MONTGOMERY = false;
  static constexpr int PRO_SQR = 445, PRO_MUL = 14;
  static constexpr int LADDER_MINBLOCKS = 2;   // resident 128-thread CTAs per SM for k_rfc7748
  static constexpr bool LADDER_STASH = true;   // scalar and x1 in shared memory (see rfc7748_sm100.cuh)
  static constexpr bool HAS_CURVE = true;
  static constexpr uint32_t A24 = 39081;
  static constexpr int COF = 2;
  static constexpr uint32_t GENERATOR = 5;
  static const char* name() { return "X448"; }

  // c = a*b (pseudo.py:616-659 / monty.py:663-872)
  static MAB_DEV void mul(uint32_t (&r)[14], const uint32_t (&a)[14], const uint32_t (&b)[14]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<144>;\n\t"
        "mul.lo.u32 t0, %14, %28;\n\t"
        "mul.hi.u32 t1, %14, %28;\n\t"
        "mul.lo.u32 t2, %16, %28;\n\t"
        "mul.hi.u32 t3, %16, %28;\n\t"
        "mul.lo.u32 t4, %18, %28;\n\t"
        "mul.hi.u32 t5, %18, %28;\n\t"
        "mul.lo.u32 t6, %20, %28;\n\t"
        "mul.hi.u32 t7, %20, %28;\n\t"
        "mul.lo.u32 t8, %22, %28;\n\t"
        "mul.hi.u32 t9, %22, %28;\n\t"
        "mul.lo.u32 t10, %24, %28;\n\t"
        "mul.hi.u32 t11, %24, %28;\n\t"
        "mul.lo.u32 t12, %26, %28;\n\t"
        "mul.hi.u32 t13, %26, %28;\n\t"
        "mul.lo.u32 t29, %15, %28;\n\t"
        "mul.hi.u32 t30, %15, %28;\n\t"
        "mul.lo.u32 t31, %17, %28;\n\t"
        "mul.hi.u32 t32, %17, %28;\n\t"
        "mul.lo.u32 t33, %19, %28;\n\t"
        "mul.hi.u32 t34, %19, %28;\n\t"
        "mul.lo.u32 t35, %21, %28;\n\t"
        "mul.hi.u32 t36, %21, %28;\n\t"
        "mul.lo.u32 t37, %23, %28;\n\t"
        "mul.hi.u32 t38, %23, %28;\n\t"
        "mul.lo.u32 t39, %25, %28;\n\t"
        "mul.hi.u32 t40, %25, %28;\n\t"
        "mul.lo.u32 t41, %27, %28;\n\t"
        "mul.hi.u32 t42, %27, %28;\n\t"
        "mad.lo.cc.u32 t2, %15, %29, t2;\n\t"
        "madc.hi.cc.u32 t3, %15, %29, t3;\n\t"
        "madc.lo.cc.u32 t4, %17, %29, t4;\n\t"
        "madc.hi.cc.u32 t5, %17, %29, t5;\n\t"
        "madc.lo.cc.u32 t6, %19, %29, t6;\n\t"
        "madc.hi.cc.u32 t7, %19, %29, t7;\n\t"
        "madc.lo.cc.u32 t8, %21, %29, t8;\n\t"
        "madc.hi.cc.u32 t9, %21, %29, t9;\n\t"
        "madc.lo.cc.u32 t10, %23, %29, t10;\n\t"
        "madc.hi.cc.u32 t11, %23, %29, t11;\n\t"
        "madc.lo.cc.u32 t12, %25, %29, t12;\n\t"
        "madc.hi.cc.u32 t13, %25, %29, t13;\n\t"
        "madc.lo.cc.u32 t14, %27, %29, 0x0;\n\t"
        "madc.hi.u32 t15, %27, %29, 0x0;\n\t"
        "mad.lo.cc.u32 t29, %14, %29, t29;\n\t"
        "madc.hi.cc.u32 t30, %14, %29, t30;\n\t"
        "madc.lo.cc.u32 t31, %16, %29, t31;\n\t"
        "madc.hi.cc.u32 t32, %16, %29, t32;\n\t"
        "madc.lo.cc.u32 t33, %18, %29, t33;\n\t"
        "madc.hi.cc.u32 t34, %18, %29, t34;\n\t"
        "madc.lo.cc.u32 t35, %20, %29, t35;\n\t"
        "madc.hi.cc.u32 t36, %20, %29, t36;\n\t"
        "madc.lo.cc.u32 t37, %22, %29, t37;\n\t"
        "madc.hi.cc.u32 t38, %22, %29, t38;\n\t"
        "madc.lo.cc.u32 t39, %24, %29, t39;\n\t"
        "madc.hi.cc.u32 t40, %24, %29, t40;\n\t"
        "madc.lo.cc.u32 t41, %26, %29, t41;\n\t"
        "madc.hi.cc.u32 t42, %26, %29, t42;\n\t"
        "addc.u32 t43, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t2, %14, %30, t2;\n\t"
        "madc.hi.cc.u32 t3, %14, %30, t3;\n\t"
        "madc.lo.cc.u32 t4, %16, %30, t4;\n\t"
        "madc.hi.cc.u32 t5, %16, %30, t5;\n\t"
        "madc.lo.cc.u32 t6, %18, %30, t6;\n\t"
        "madc.hi.cc.u32 t7, %18, %30, t7;\n\t"
        "madc.lo.cc.u32 t8, %20, %30, t8;\n\t"
        "madc.hi.cc.u32 t9, %20, %30, t9;\n\t"
        "madc.lo.cc.u32 t10, %22, %30, t10;\n\t"
        "madc.hi.cc.u32 t11, %22, %30, t11;\n\t"
        "madc.lo.cc.u32 t12, %24, %30, t12;\n\t"
        "madc.hi.cc.u32 t13, %24, %30, t13;\n\t"
        "madc.lo.cc.u32 t14, %26, %30, t14;\n\t"
        "madc.hi.cc.u32 t15, %26, %30, t15;\n\t"
        "addc.u32 t16, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t31, %15, %30, t31;\n\t"
        "madc.hi.cc.u32 t32, %15, %30, t32;\n\t"
        "madc.lo.cc.u32 t33, %17, %30, t33;\n\t"
        "madc.hi.cc.u32 t34, %17, %30, t34;\n\t"
        "madc.lo.cc.u32 t35, %19, %30, t35;\n\t"
        "madc.hi.cc.u32 t36, %19, %30, t36;\n\t"
        "madc.lo.cc.u32 t37, %21, %30, t37;\n\t"
        "madc.hi.cc.u32 t38, %21, %30, t38;\n\t"
        "madc.lo.cc.u32 t39, %23, %30, t39;\n\t"
        "madc.hi.cc.u32 t40, %23, %30, t40;\n\t"
        "madc.lo.cc.u32 t41, %25, %30, t41;\n\t"
        "madc.hi.cc.u32 t42, %25, %30, t42;\n\t"
        "madc.lo.cc.u32 t43, %27, %30, t43;\n\t"
        "madc.hi.u32 t44, %27, %30, 0x0;\n\t"
        "mad.lo.cc.u32 t4, %15, %31, t4;\n\t"
        "madc.hi.cc.u32 t5, %15, %31, t5;\n\t"
        "madc.lo.cc.u32 t6, %17, %31, t6;\n\t"
        "madc.hi.cc.u32 t7, %17, %31, t7;\n\t"
        "madc.lo.cc.u32 t8, %19, %31, t8;\n\t"
        "madc.hi.cc.u32 t9, %19, %31, t9;\n\t"
        "madc.lo.cc.u32 t10, %21, %31, t10;\n\t"
        "madc.hi.cc.u32 t11, %21, %31, t11;\n\t"
        "madc.lo.cc.u32 t12, %23, %31, t12;\n\t"
        "madc.hi.cc.u32 t13, %23, %31, t13;\n\t"
        "madc.lo.cc.u32 t14, %25, %31, t14;\n\t"
        "madc.hi.cc.u32 t15, %25, %31, t15;\n\t"
        "madc.lo.cc.u32 t16, %27, %31, t16;\n\t"
        "madc.hi.u32 t17, %27, %31, 0x0;\n\t"
        "mad.lo.cc.u32 t31, %14, %31, t31;\n\t"
        "madc.hi.cc.u32 t32, %14, %31, t32;\n\t"
        "madc.lo.cc.u32 t33, %16, %31, t33;\n\t"
        "madc.hi.cc.u32 t34, %16, %31, t34;\n\t"
        "madc.lo.cc.u32 t35, %18, %31, t35;\n\t"
        "madc.hi.cc.u32 t36, %18, %31, t36;\n\t"
        "madc.lo.cc.u32 t37, %20, %31, t37;\n\t"
        "madc.hi.cc.u32 t38, %20, %31, t38;\n\t"
        "madc.lo.cc.u32 t39, %22, %31, t39;\n\t"
        "madc.hi.cc.u32 t40, %22, %31, t40;\n\t"
        "madc.lo.cc.u32 t41, %24, %31, t41;\n\t"
        "madc.hi.cc.u32 t42, %24, %31, t42;\n\t"
        "madc.lo.cc.u32 t43, %26, %31, t43;\n\t"
        "madc.hi.cc.u32 t44, %26, %31, t44;\n\t"
        "addc.u32 t45, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t4, %14, %32, t4;\n\t"
        "madc.hi.cc.u32 t5, %14, %32, t5;\n\t"
        "madc.lo.cc.u32 t6, %16, %32, t6;\n\t"
        "madc.hi.cc.u32 t7, %16, %32, t7;\n\t"
        "madc.lo.cc.u32 t8, %18, %32, t8;\n\t"
        "madc.hi.cc.u32 t9, %18, %32, t9;\n\t"
        "madc.lo.cc.u32 t10, %20, %32, t10;\n\t"
        "madc.hi.cc.u32 t11, %20, %32, t11;\n\t"
        "madc.lo.cc.u32 t12, %22, %32, t12;\n\t"
        "madc.hi.cc.u32 t13, %22, %32, t13;\n\t"
        "madc.lo.cc.u32 t14, %24, %32, t14;\n\t"
        "madc.hi.cc.u32 t15, %24, %32, t15;\n\t"
        "madc.lo.cc.u32 t16, %26, %32, t16;\n\t"
        "madc.hi.cc.u32 t17, %26, %32, t17;\n\t"
        "addc.u32 t18, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t33, %15, %32, t33;\n\t"
        "madc.hi.cc.u32 t34, %15, %32, t34;\n\t"
        "madc.lo.cc.u32 t35, %17, %32, t35;\n\t"
        "madc.hi.cc.u32 t36, %17, %32, t36;\n\t"
        "madc.lo.cc.u32 t37, %19, %32, t37;\n\t"
        "madc.hi.cc.u32 t38, %19, %32, t38;\n\t"
        "madc.lo.cc.u32 t39, %21, %32, t39;\n\t"
        "madc.hi.cc.u32 t40, %21, %32, t40;\n\t"
        "madc.lo.cc.u32 t41, %23, %32, t41;\n\t"
        "madc.hi.cc.u32 t42, %23, %32, t42;\n\t"
        "madc.lo.cc.u32 t43, %25, %32, t43;\n\t"
        "madc.hi.cc.u32 t44, %25, %32, t44;\n\t"
        "madc.lo.cc.u32 t45, %27, %32, t45;\n\t"
        "madc.hi.u32 t46, %27, %32, 0x0;\n\t"
        "mad.lo.cc.u32 t6, %15, %33, t6;\n\t"
        "madc.hi.cc.u32 t7, %15, %33, t7;\n\t"
        "madc.lo.cc.u32 t8, %17, %33, t8;\n\t"
        "madc.hi.cc.u32 t9, %17, %33, t9;\n\t"
        "madc.lo.cc.u32 t10, %19, %33, t10;\n\t"
        "madc.hi.cc.u32 t11, %19, %33, t11;\n\t"
        "madc.lo.cc.u32 t12, %21, %33, t12;\n\t"
        "madc.hi.cc.u32 t13, %21, %33, t13;\n\t"
        "madc.lo.cc.u32 t14, %23, %33, t14;\n\t"
        "madc.hi.cc.u32 t15, %23, %33, t15;\n\t"
        "madc.lo.cc.u32 t16, %25, %33, t16;\n\t"
        "madc.hi.cc.u32 t17, %25, %33, t17;\n\t"
        "madc.lo.cc.u32 t18, %27, %33, t18;\n\t"
        "madc.hi.u32 t19, %27, %33, 0x0;\n\t"
        "mad.lo.cc.u32 t33, %14, %33, t33;\n\t"
        "madc.hi.cc.u32 t34, %14, %33, t34;\n\t"
        "madc.lo.cc.u32 t35, %16, %33, t35;\n\t"
        "madc.hi.cc.u32 t36, %16, %33, t36;\n\t"
        "madc.lo.cc.u32 t37, %18, %33, t37;\n\t"
        "madc.hi.cc.u32 t38, %18, %33, t38;\n\t"
        "madc.lo.cc.u32 t39, %20, %33, t39;\n\t"
        "madc.hi.cc.u32 t40, %20, %33, t40;\n\t"
        "madc.lo.cc.u32 t41, %22, %33, t41;\n\t"
        "madc.hi.cc.u32 t42, %22, %33, t42;\n\t"
        "madc.lo.cc.u32 t43, %24, %33, t43;\n\t"
        "madc.hi.cc.u32 t44, %24, %33, t44;\n\t"
        "madc.lo.cc.u32 t45, %26, %33, t45;\n\t"
        "madc.hi.cc.u32 t46, %26, %33, t46;\n\t"
        "addc.u32 t47, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t6, %14, %34, t6;\n\t"
        "madc.hi.cc.u32 t7, %14, %34, t7;\n\t"
        "madc.lo.cc.u32 t8, %16, %34, t8;\n\t"
        "madc.hi.cc.u32 t9, %16, %34, t9;\n\t"
        "madc.lo.cc.u32 t10, %18, %34, t10;\n\t"
        "madc.hi.cc.u32 t11, %18, %34, t11;\n\t"
        "madc.lo.cc.u32 t12, %20, %34, t12;\n\t"
        "madc.hi.cc.u32 t13, %20, %34, t13;\n\t"
        "madc.lo.cc.u32 t14, %22, %34, t14;\n\t"
        "madc.hi.cc.u32 t15, %22, %34, t15;\n\t"
        "madc.lo.cc.u32 t16, %24, %34, t16;\n\t"
        "madc.hi.cc.u32 t17, %24, %34, t17;\n\t"
        "madc.lo.cc.u32 t18, %26, %34, t18;\n\t"
        "madc.hi.cc.u32 t19, %26, %34, t19;\n\t"
        "addc.u32 t20, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t35, %15, %34, t35;\n\t"
        "madc.hi.cc.u32 t36, %15, %34, t36;\n\t"
        "madc.lo.cc.u32 t37, %17, %34, t37;\n\t"
        "madc.hi.cc.u32 t38, %17, %34, t38;\n\t"
        "madc.lo.cc.u32 t39, %19, %34, t39;\n\t"
        "madc.hi.cc.u32 t40, %19, %34, t40;\n\t"
        "madc.lo.cc.u32 t41, %21, %34, t41;\n\t"
        "madc.hi.cc.u32 t42, %21, %34, t42;\n\t"
        "madc.lo.cc.u32 t43, %23, %34, t43;\n\t"
        "madc.hi.cc.u32 t44, %23, %34, t44;\n\t"
        "madc.lo.cc.u32 t45, %25, %34, t45;\n\t"
        "madc.hi.cc.u32 t46, %25, %34, t46;\n\t"
        "madc.lo.cc.u32 t47, %27, %34, t47;\n\t"
        "madc.hi.u32 t48, %27, %34, 0x0;\n\t"
        "mad.lo.cc.u32 t8, %15, %35, t8;\n\t"
        "madc.hi.cc.u32 t9, %15, %35, t9;\n\t"
        "madc.lo.cc.u32 t10, %17, %35, t10;\n\t"
        "madc.hi.cc.u32 t11, %17, %35, t11;\n\t"
        "madc.lo.cc.u32 t12, %19, %35, t12;\n\t"
        "madc.hi.cc.u32 t13, %19, %35, t13;\n\t"
        "madc.lo.cc.u32 t14, %21, %35, t14;\n\t"
        "madc.hi.cc.u32 t15, %21, %35, t15;\n\t"
        "madc.lo.cc.u32 t16, %23, %35, t16;\n\t"
        "madc.hi.cc.u32 t17, %23, %35, t17;\n\t"
        "madc.lo.cc.u32 t18, %25, %35, t18;\n\t"
        "madc.hi.cc.u32 t19, %25, %35, t19;\n\t"
        "madc.lo.cc.u32 t20, %27, %35, t20;\n\t"
        "madc.hi.u32 t21, %27, %35, 0x0;\n\t"
        "mad.lo.cc.u32 t35, %14, %35, t35;\n\t"
        "madc.hi.cc.u32 t36, %14, %35, t36;\n\t"
        "madc.lo.cc.u32 t37, %16, %35, t37;\n\t"
        "madc.hi.cc.u32 t38, %16, %35, t38;\n\t"
        "madc.lo.cc.u32 t39, %18, %35, t39;\n\t"
        "madc.hi.cc.u32 t40, %18, %35, t40;\n\t"
        "madc.lo.cc.u32 t41, %20, %35, t41;\n\t"
        "madc.hi.cc.u32 t42, %20, %35, t42;\n\t"
        "madc.lo.cc.u32 t43, %22, %35, t43;\n\t"
        "madc.hi.cc.u32 t44, %22, %35, t44;\n\t"
        "madc.lo.cc.u32 t45, %24, %35, t45;\n\t"
        "madc.hi.cc.u32 t46, %24, %35, t46;\n\t"
        "madc.lo.cc.u32 t47, %26, %35, t47;\n\t"
        "madc.hi.cc.u32 t48, %26, %35, t48;\n\t"
        "addc.u32 t49, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t8, %14, %36, t8;\n\t"
        "madc.hi.cc.u32 t9, %14, %36, t9;\n\t"
        "madc.lo.cc.u32 t10, %16, %36, t10;\n\t"
        "madc.hi.cc.u32 t11, %16, %36, t11;\n\t"
        "madc.lo.cc.u32 t12, %18, %36, t12;\n\t"
        "madc.hi.cc.u32 t13, %18, %36, t13;\n\t"
        "madc.lo.cc.u32 t14, %20, %36, t14;\n\t"
        "madc.hi.cc.u32 t15, %20, %36, t15;\n\t"
        "madc.lo.cc.u32 t16, %22, %36, t16;\n\t"
        "madc.hi.cc.u32 t17, %22, %36, t17;\n\t"
        "madc.lo.cc.u32 t18, %24, %36, t18;\n\t"
        "madc.hi.cc.u32 t19, %24, %36, t19;\n\t"
        "madc.lo.cc.u32 t20, %26, %36, t20;\n\t"
        "madc.hi.cc.u32 t21, %26, %36, t21;\n\t"
        "addc.u32 t22, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t37, %15, %36, t37;\n\t"
        "madc.hi.cc.u32 t38, %15, %36, t38;\n\t"
        "madc.lo.cc.u32 t39, %17, %36, t39;\n\t"
        "madc.hi.cc.u32 t40, %17, %36, t40;\n\t"
        "madc.lo.cc.u32 t41, %19, %36, t41;\n\t"
        "madc.hi.cc.u32 t42, %19, %36, t42;\n\t"
        "madc.lo.cc.u32 t43, %21, %36, t43;\n\t"
        "madc.hi.cc.u32 t44, %21, %36, t44;\n\t"
        "madc.lo.cc.u32 t45, %23, %36, t45;\n\t"
        "madc.hi.cc.u32 t46, %23, %36, t46;\n\t"
        "madc.lo.cc.u32 t47, %25, %36, t47;\n\t"
        "madc.hi.cc.u32 t48, %25, %36, t48;\n\t"
        "madc.lo.cc.u32 t49, %27, %36, t49;\n\t"
        "madc.hi.u32 t50, %27, %36, 0x0;\n\t"
        "mad.lo.cc.u32 t10, %15, %37, t10;\n\t"
        "madc.hi.cc.u32 t11, %15, %37, t11;\n\t"
        "madc.lo.cc.u32 t12, %17, %37, t12;\n\t"
        "madc.hi.cc.u32 t13, %17, %37, t13;\n\t"
        "madc.lo.cc.u32 t14, %19, %37, t14;\n\t"
        "madc.hi.cc.u32 t15, %19, %37, t15;\n\t"
        "madc.lo.cc.u32 t16, %21, %37, t16;\n\t"
        "madc.hi.cc.u32 t17, %21, %37, t17;\n\t"
        "madc.lo.cc.u32 t18, %23, %37, t18;\n\t"
        "madc.hi.cc.u32 t19, %23, %37, t19;\n\t"
        "madc.lo.cc.u32 t20, %25, %37, t20;\n\t"
        "madc.hi.cc.u32 t21, %25, %37, t21;\n\t"
        "madc.lo.cc.u32 t22, %27, %37, t22;\n\t"
        "madc.hi.u32 t23, %27, %37, 0x0;\n\t"
        "mad.lo.cc.u32 t37, %14, %37, t37;\n\t"
        "madc.hi.cc.u32 t38, %14, %37, t38;\n\t"
        "madc.lo.cc.u32 t39, %16, %37, t39;\n\t"
        "madc.hi.cc.u32 t40, %16, %37, t40;\n\t"
        "madc.lo.cc.u32 t41, %18, %37, t41;\n\t"
        "madc.hi.cc.u32 t42, %18, %37, t42;\n\t"
        "madc.lo.cc.u32 t43, %20, %37, t43;\n\t"
        "madc.hi.cc.u32 t44, %20, %37, t44;\n\t"
        "madc.lo.cc.u32 t45, %22, %37, t45;\n\t"
        "madc.hi.cc.u32 t46, %22, %37, t46;\n\t"
        "madc.lo.cc.u32 t47, %24, %37, t47;\n\t"
        "madc.hi.cc.u32 t48, %24, %37, t48;\n\t"
        "madc.lo.cc.u32 t49, %26, %37, t49;\n\t"
        "madc.hi.cc.u32 t50, %26, %37, t50;\n\t"
        "addc.u32 t51, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t10, %14, %38, t10;\n\t"
        "madc.hi.cc.u32 t11, %14, %38, t11;\n\t"
        "madc.lo.cc.u32 t12, %16, %38, t12;\n\t"
        "madc.hi.cc.u32 t13, %16, %38, t13;\n\t"
        "madc.lo.cc.u32 t14, %18, %38, t14;\n\t"
        "madc.hi.cc.u32 t15, %18, %38, t15;\n\t"
        "madc.lo.cc.u32 t16, %20, %38, t16;\n\t"
        "madc.hi.cc.u32 t17, %20, %38, t17;\n\t"
        "madc.lo.cc.u32 t18, %22, %38, t18;\n\t"
        "madc.hi.cc.u32 t19, %22, %38, t19;\n\t"
        "madc.lo.cc.u32 t20, %24, %38, t20;\n\t"
        "madc.hi.cc.u32 t21, %24, %38, t21;\n\t"
        "madc.lo.cc.u32 t22, %26, %38, t22;\n\t"
        "madc.hi.cc.u32 t23, %26, %38, t23;\n\t"
        "addc.u32 t24, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t39, %15, %38, t39;\n\t"
        "madc.hi.cc.u32 t40, %15, %38, t40;\n\t"
        "madc.lo.cc.u32 t41, %17, %38, t41;\n\t"
        "madc.hi.cc.u32 t42, %17, %38, t42;\n\t"
        "madc.lo.cc.u32 t43, %19, %38, t43;\n\t"
        "madc.hi.cc.u32 t44, %19, %38, t44;\n\t"
        "madc.lo.cc.u32 t45, %21, %38, t45;\n\t"
        "madc.hi.cc.u32 t46, %21, %38, t46;\n\t"
        "madc.lo.cc.u32 t47, %23, %38, t47;\n\t"
        "madc.hi.cc.u32 t48, %23, %38, t48;\n\t"
        "madc.lo.cc.u32 t49, %25, %38, t49;\n\t"
        "madc.hi.cc.u32 t50, %25, %38, t50;\n\t"
        "madc.lo.cc.u32 t51, %27, %38, t51;\n\t"
        "madc.hi.u32 t52, %27, %38, 0x0;\n\t"
        "mad.lo.cc.u32 t12, %15, %39, t12;\n\t"
        "madc.hi.cc.u32 t13, %15, %39, t13;\n\t"
        "madc.lo.cc.u32 t14, %17, %39, t14;\n\t"
        "madc.hi.cc.u32 t15, %17, %39, t15;\n\t"
        "madc.lo.cc.u32 t16, %19, %39, t16;\n\t"
        "madc.hi.cc.u32 t17, %19, %39, t17;\n\t"
        "madc.lo.cc.u32 t18, %21, %39, t18;\n\t"
        "madc.hi.cc.u32 t19, %21, %39, t19;\n\t"
        "madc.lo.cc.u32 t20, %23, %39, t20;\n\t"
        "madc.hi.cc.u32 t21, %23, %39, t21;\n\t"
        "madc.lo.cc.u32 t22, %25, %39, t22;\n\t"
        "madc.hi.cc.u32 t23, %25, %39, t23;\n\t"
        "madc.lo.cc.u32 t24, %27, %39, t24;\n\t"
        "madc.hi.u32 t25, %27, %39, 0x0;\n\t"
        "mad.lo.cc.u32 t39, %14, %39, t39;\n\t"
        "madc.hi.cc.u32 t40, %14, %39, t40;\n\t"
        "madc.lo.cc.u32 t41, %16, %39, t41;\n\t"
        "madc.hi.cc.u32 t42, %16, %39, t42;\n\t"
        "madc.lo.cc.u32 t43, %18, %39, t43;\n\t"
        "madc.hi.cc.u32 t44, %18, %39, t44;\n\t"
        "madc.lo.cc.u32 t45, %20, %39, t45;\n\t"
        "madc.hi.cc.u32 t46, %20, %39, t46;\n\t"
        "madc.lo.cc.u32 t47, %22, %39, t47;\n\t"
        "madc.hi.cc.u32 t48, %22, %39, t48;\n\t"
        "madc.lo.cc.u32 t49, %24, %39, t49;\n\t"
        "madc.hi.cc.u32 t50, %24, %39, t50;\n\t"
        "madc.lo.cc.u32 t51, %26, %39, t51;\n\t"
        "madc.hi.cc.u32 t52, %26, %39, t52;\n\t"
        "addc.u32 t53, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t12, %14, %40, t12;\n\t"
        "madc.hi.cc.u32 t13, %14, %40, t13;\n\t"
        "madc.lo.cc.u32 t14, %16, %40, t14;\n\t"
        "madc.hi.cc.u32 t15, %16, %40, t15;\n\t"
        "madc.lo.cc.u32 t16, %18, %40, t16;\n\t"
        "madc.hi.cc.u32 t17, %18, %40, t17;\n\t"
        "madc.lo.cc.u32 t18, %20, %40, t18;\n\t"
        "madc.hi.cc.u32 t19, %20, %40, t19;\n\t"
        "madc.lo.cc.u32 t20, %22, %40, t20;\n\t"
        "madc.hi.cc.u32 t21, %22, %40, t21;\n\t"
        "madc.lo.cc.u32 t22, %24, %40, t22;\n\t"
        "madc.hi.cc.u32 t23, %24, %40, t23;\n\t"
        "madc.lo.cc.u32 t24, %26, %40, t24;\n\t"
        "madc.hi.cc.u32 t25, %26, %40, t25;\n\t"
        "addc.u32 t26, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t41, %15, %40, t41;\n\t"
        "madc.hi.cc.u32 t42, %15, %40, t42;\n\t"
        "madc.lo.cc.u32 t43, %17, %40, t43;\n\t"
        "madc.hi.cc.u32 t44, %17, %40, t44;\n\t"
        "madc.lo.cc.u32 t45, %19, %40, t45;\n\t"
        "madc.hi.cc.u32 t46, %19, %40, t46;\n\t"
        "madc.lo.cc.u32 t47, %21, %40, t47;\n\t"
        "madc.hi.cc.u32 t48, %21, %40, t48;\n\t"
        "madc.lo.cc.u32 t49, %23, %40, t49;\n\t"
        "madc.hi.cc.u32 t50, %23, %40, t50;\n\t"
        "madc.lo.cc.u32 t51, %25, %40, t51;\n\t"
        "madc.hi.cc.u32 t52, %25, %40, t52;\n\t"
        "madc.lo.cc.u32 t53, %27, %40, t53;\n\t"
        "madc.hi.u32 t54, %27, %40, 0x0;\n\t"
        "mad.lo.cc.u32 t14, %15, %41, t14;\n\t"
        "madc.hi.cc.u32 t15, %15, %41, t15;\n\t"
        "madc.lo.cc.u32 t16, %17, %41, t16;\n\t"
        "madc.hi.cc.u32 t17, %17, %41, t17;\n\t"
        "madc.lo.cc.u32 t18, %19, %41, t18;\n\t"
        "madc.hi.cc.u32 t19, %19, %41, t19;\n\t"
        "madc.lo.cc.u32 t20, %21, %41, t20;\n\t"
        "madc.hi.cc.u32 t21, %21, %41, t21;\n\t"
        "madc.lo.cc.u32 t22, %23, %41, t22;\n\t"
        "madc.hi.cc.u32 t23, %23, %41, t23;\n\t"
        "madc.lo.cc.u32 t24, %25, %41, t24;\n\t"
        "madc.hi.cc.u32 t25, %25, %41, t25;\n\t"
        "madc.lo.cc.u32 t26, %27, %41, t26;\n\t"
        "madc.hi.u32 t27, %27, %41, 0x0;\n\t"
        "mad.lo.cc.u32 t41, %14, %41, t41;\n\t"
        "madc.hi.cc.u32 t42, %14, %41, t42;\n\t"
        "madc.lo.cc.u32 t43, %16, %41, t43;\n\t"
        "madc.hi.cc.u32 t44, %16, %41, t44;\n\t"
        "madc.lo.cc.u32 t45, %18, %41, t45;\n\t"
        "madc.hi.cc.u32 t46, %18, %41, t46;\n\t"
        "madc.lo.cc.u32 t47, %20, %41, t47;\n\t"
        "madc.hi.cc.u32 t48, %20, %41, t48;\n\t"
        "madc.lo.cc.u32 t49, %22, %41, t49;\n\t"
        "madc.hi.cc.u32 t50, %22, %41, t50;\n\t"
        "madc.lo.cc.u32 t51, %24, %41, t51;\n\t"
        "madc.hi.cc.u32 t52, %24, %41, t52;\n\t"
        "madc.lo.cc.u32 t53, %26, %41, t53;\n\t"
        "madc.hi.cc.u32 t54, %26, %41, t54;\n\t"
        "addc.u32 t55, 0x0, 0x0;\n\t"
        "add.cc.u32 t56, t1, t29;\n\t"
        "addc.cc.u32 t57, t2, t30;\n\t"
        "addc.cc.u32 t58, t3, t31;\n\t"
        "addc.cc.u32 t59, t4, t32;\n\t"
        "addc.cc.u32 t60, t5, t33;\n\t"
        "addc.cc.u32 t61, t6, t34;\n\t"
        "addc.cc.u32 t62, t7, t35;\n\t"
        "addc.cc.u32 t63, t8, t36;\n\t"
        "addc.cc.u32 t64, t9, t37;\n\t"
        "addc.cc.u32 t65, t10, t38;\n\t"
        "addc.cc.u32 t66, t11, t39;\n\t"
        "addc.cc.u32 t67, t12, t40;\n\t"
        "addc.cc.u32 t68, t13, t41;\n\t"
        "addc.cc.u32 t69, t14, t42;\n\t"
        "addc.cc.u32 t70, t15, t43;\n\t"
        "addc.cc.u32 t71, t16, t44;\n\t"
        "addc.cc.u32 t72, t17, t45;\n\t"
        "addc.cc.u32 t73, t18, t46;\n\t"
        "addc.cc.u32 t74, t19, t47;\n\t"
        "addc.cc.u32 t75, t20, t48;\n\t"
        "addc.cc.u32 t76, t21, t49;\n\t"
        "addc.cc.u32 t77, t22, t50;\n\t"
        "addc.cc.u32 t78, t23, t51;\n\t"
        "addc.cc.u32 t79, t24, t52;\n\t"
        "addc.cc.u32 t80, t25, t53;\n\t"
        "addc.cc.u32 t81, t26, t54;\n\t"
        "addc.u32 t82, t27, t55;\n\t"
        "add.cc.u32 t84, t0, t69;\n\t"
        "addc.cc.u32 t85, t56, t70;\n\t"
        "addc.cc.u32 t86, t57, t71;\n\t"
        "addc.cc.u32 t87, t58, t72;\n\t"
        "addc.cc.u32 t88, t59, t73;\n\t"
        "addc.cc.u32 t89, t60, t74;\n\t"
        "addc.cc.u32 t90, t61, t75;\n\t"
        "addc.cc.u32 t91, t62, t69;\n\t"
        "addc.cc.u32 t92, t63, t70;\n\t"
        "addc.cc.u32 t93, t64, t71;\n\t"
        "addc.cc.u32 t94, t65, t72;\n\t"
        "addc.cc.u32 t95, t66, t73;\n\t"
        "addc.cc.u32 t96, t67, t74;\n\t"
        "addc.cc.u32 t97, t68, t75;\n\t"
        "addc.u32 t83, 0x0, 0x0;\n\t"
        "add.cc.u32 t98, t84, t76;\n\t"
        "addc.cc.u32 t99, t85, t77;\n\t"
        "addc.cc.u32 t100, t86, t78;\n\t"
        "addc.cc.u32 t101, t87, t79;\n\t"
        "addc.cc.u32 t102, t88, t80;\n\t"
        "addc.cc.u32 t103, t89, t81;\n\t"
        "addc.cc.u32 t104, t90, t82;\n\t"
        "addc.cc.u32 t105, t91, t76;\n\t"
        "addc.cc.u32 t106, t92, t77;\n\t"
        "addc.cc.u32 t107, t93, t78;\n\t"
        "addc.cc.u32 t108, t94, t79;\n\t"
        "addc.cc.u32 t109, t95, t80;\n\t"
        "addc.cc.u32 t110, t96, t81;\n\t"
        "addc.cc.u32 t111, t97, t82;\n\t"
        "addc.u32 t112, t83, 0x0;\n\t"
        "add.cc.u32 t113, t105, t76;\n\t"
        "addc.cc.u32 t114, t106, t77;\n\t"
        "addc.cc.u32 t115, t107, t78;\n\t"
        "addc.cc.u32 t116, t108, t79;\n\t"
        "addc.cc.u32 t117, t109, t80;\n\t"
        "addc.cc.u32 t118, t110, t81;\n\t"
        "addc.cc.u32 t119, t111, t82;\n\t"
        "addc.u32 t120, t112, 0x0;\n\t"
        "add.cc.u32 t121, t98, t120;\n\t"
        "addc.cc.u32 t122, t99, 0x0;\n\t"
        "addc.cc.u32 t123, t100, 0x0;\n\t"
        "addc.cc.u32 t124, t101, 0x0;\n\t"
        "addc.cc.u32 t125, t102, 0x0;\n\t"
        "addc.cc.u32 t126, t103, 0x0;\n\t"
        "addc.cc.u32 t127, t104, 0x0;\n\t"
        "addc.cc.u32 t128, t113, t120;\n\t"
        "addc.cc.u32 t129, t114, 0x0;\n\t"
        "addc.cc.u32 t130, t115, 0x0;\n\t"
        "addc.cc.u32 t131, t116, 0x0;\n\t"
        "addc.cc.u32 t132, t117, 0x0;\n\t"
        "addc.cc.u32 t133, t118, 0x0;\n\t"
        "addc.cc.u32 t134, t119, 0x0;\n\t"
        "addc.u32 t135, 0x0, 0x0;\n\t"
        "add.cc.u32 t136, t121, t135;\n\t"
        "addc.cc.u32 t137, t122, 0x0;\n\t"
        "addc.cc.u32 t138, t123, 0x0;\n\t"
        "addc.cc.u32 t139, t124, 0x0;\n\t"
        "addc.cc.u32 t140, t125, 0x0;\n\t"
        "addc.cc.u32 t141, t126, 0x0;\n\t"
        "addc.cc.u32 t142, t127, 0x0;\n\t"
        "addc.u32 t143, t128, t135;\n\t"
        "mov.u32 %0, t136;\n\t"
        "mov.u32 %1, t137;\n\t"
        "mov.u32 %2, t138;\n\t"
        "mov.u32 %3, t139;\n\t"
        "mov.u32 %4, t140;\n\t"
        "mov.u32 %5, t141;\n\t"
        "mov.u32 %6, t142;\n\t"
        "mov.u32 %7, t143;\n\t"
        "mov.u32 %8, t129;\n\t"
        "mov.u32 %9, t130;\n\t"
        "mov.u32 %10, t131;\n\t"
        "mov.u32 %11, t132;\n\t"
        "mov.u32 %12, t133;\n\t"
        "mov.u32 %13, t134;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(b[12]), "r"(b[13]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t a_8_i = a[8];
    const uint32_t a_9_i = a[9];
    const uint32_t a_10_i = a[10];
    const uint32_t a_11_i = a[11];
    const uint32_t a_12_i = a[12];
    const uint32_t a_13_i = a[13];
    const uint32_t b_0_i = b[0];
    const uint32_t b_1_i = b[1];
    const uint32_t b_2_i = b[2];
    const uint32_t b_3_i = b[3];
    const uint32_t b_4_i = b[4];
    const uint32_t b_5_i = b[5];
    const uint32_t b_6_i = b[6];
    const uint32_t b_7_i = b[7];
    const uint32_t b_8_i = b[8];
    const uint32_t b_9_i = b[9];
    const uint32_t b_10_i = b[10];
    const uint32_t b_11_i = b[11];
    const uint32_t b_12_i = b[12];
    const uint32_t b_13_i = b[13];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127, t128, t129, t130, t131, t132, t133, t134, t135, t136, t137, t138, t139, t140, t141, t142, t143;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t0 = (uint32_t)((uint32_t)(a_0_i * b_0_i));
    t1 = (uint32_t)(((uint64_t)a_0_i * b_0_i) >> 32);
    t2 = (uint32_t)((uint32_t)(a_2_i * b_0_i));
    t3 = (uint32_t)(((uint64_t)a_2_i * b_0_i) >> 32);
    t4 = (uint32_t)((uint32_t)(a_4_i * b_0_i));
    t5 = (uint32_t)(((uint64_t)a_4_i * b_0_i) >> 32);
    t6 = (uint32_t)((uint32_t)(a_6_i * b_0_i));
    t7 = (uint32_t)(((uint64_t)a_6_i * b_0_i) >> 32);
    t8 = (uint32_t)((uint32_t)(a_8_i * b_0_i));
    t9 = (uint32_t)(((uint64_t)a_8_i * b_0_i) >> 32);
    t10 = (uint32_t)((uint32_t)(a_10_i * b_0_i));
    t11 = (uint32_t)(((uint64_t)a_10_i * b_0_i) >> 32);
    t12 = (uint32_t)((uint32_t)(a_12_i * b_0_i));
    t13 = (uint32_t)(((uint64_t)a_12_i * b_0_i) >> 32);
    t29 = (uint32_t)((uint32_t)(a_1_i * b_0_i));
    t30 = (uint32_t)(((uint64_t)a_1_i * b_0_i) >> 32);
    t31 = (uint32_t)((uint32_t)(a_3_i * b_0_i));
    t32 = (uint32_t)(((uint64_t)a_3_i * b_0_i) >> 32);
    t33 = (uint32_t)((uint32_t)(a_5_i * b_0_i));
    t34 = (uint32_t)(((uint64_t)a_5_i * b_0_i) >> 32);
    t35 = (uint32_t)((uint32_t)(a_7_i * b_0_i));
    t36 = (uint32_t)(((uint64_t)a_7_i * b_0_i) >> 32);
    t37 = (uint32_t)((uint32_t)(a_9_i * b_0_i));
    t38 = (uint32_t)(((uint64_t)a_9_i * b_0_i) >> 32);
    t39 = (uint32_t)((uint32_t)(a_11_i * b_0_i));
    t40 = (uint32_t)(((uint64_t)a_11_i * b_0_i) >> 32);
    t41 = (uint32_t)((uint32_t)(a_13_i * b_0_i));
    t42 = (uint32_t)(((uint64_t)a_13_i * b_0_i) >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * b_1_i) + t2; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_1_i) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_1_i) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_1_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_1_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_1_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_1_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_1_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_1_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_1_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_1_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_1_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_1_i) + 0x0u + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_1_i) >> 32) + 0x0u + cf_; t15 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_1_i) + t29; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_1_i) >> 32) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_1_i) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_1_i) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_1_i) + t33 + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_1_i) >> 32) + t34 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_1_i) + t35 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_1_i) >> 32) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_1_i) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_1_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_1_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_1_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_1_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_1_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t43 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_2_i) + t2; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_2_i) >> 32) + t3 + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_2_i) + t4 + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_2_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_2_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_2_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_2_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_2_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_2_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_2_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_2_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_2_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_2_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_2_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t16 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_2_i) + t31; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_2_i) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_2_i) + t33 + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_2_i) >> 32) + t34 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_2_i) + t35 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_2_i) >> 32) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_2_i) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_2_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_2_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_2_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_2_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_2_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_2_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_2_i) >> 32) + 0x0u + cf_; t44 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_3_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_3_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_3_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_3_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_3_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_3_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_3_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_3_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_3_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_3_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_3_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_3_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_3_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_3_i) >> 32) + 0x0u + cf_; t17 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_3_i) + t31; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_3_i) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_3_i) + t33 + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_3_i) >> 32) + t34 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_3_i) + t35 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_3_i) >> 32) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_3_i) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_3_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_3_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_3_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_3_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_3_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_3_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_3_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t45 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_4_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_4_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_4_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_4_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_4_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_4_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_4_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_4_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_4_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_4_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_4_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_4_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_4_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_4_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t18 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_4_i) + t33; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_4_i) >> 32) + t34 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_4_i) + t35 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_4_i) >> 32) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_4_i) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_4_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_4_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_4_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_4_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_4_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_4_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_4_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_4_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_4_i) >> 32) + 0x0u + cf_; t46 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_5_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_5_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_5_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_5_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_5_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_5_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_5_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_5_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_5_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_5_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_5_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_5_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_5_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_5_i) >> 32) + 0x0u + cf_; t19 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_5_i) + t33; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_5_i) >> 32) + t34 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_5_i) + t35 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_5_i) >> 32) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_5_i) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_5_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_5_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_5_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_5_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_5_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_5_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_5_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_5_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_5_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t47 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_6_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_6_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_6_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_6_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_6_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_6_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_6_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_6_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_6_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_6_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_6_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_6_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_6_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_6_i) >> 32) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t20 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_6_i) + t35; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_6_i) >> 32) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_6_i) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_6_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_6_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_6_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_6_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_6_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_6_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_6_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_6_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_6_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_6_i) + t47 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_6_i) >> 32) + 0x0u + cf_; t48 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_7_i) + t8; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_7_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_7_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_7_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_7_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_7_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_7_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_7_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_7_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_7_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_7_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_7_i) >> 32) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_7_i) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_7_i) >> 32) + 0x0u + cf_; t21 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_7_i) + t35; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_7_i) >> 32) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_7_i) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_7_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_7_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_7_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_7_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_7_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_7_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_7_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_7_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_7_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_7_i) + t47 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_7_i) >> 32) + t48 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t49 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_8_i) + t8; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_8_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_8_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_8_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_8_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_8_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_8_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_8_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_8_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_8_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_8_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_8_i) >> 32) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_8_i) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_8_i) >> 32) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t22 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_8_i) + t37; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_8_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_8_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_8_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_8_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_8_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_8_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_8_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_8_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_8_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_8_i) + t47 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_8_i) >> 32) + t48 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_8_i) + t49 + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_8_i) >> 32) + 0x0u + cf_; t50 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_9_i) + t10; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_9_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_9_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_9_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_9_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_9_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_9_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_9_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_9_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_9_i) >> 32) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_9_i) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_9_i) >> 32) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_9_i) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_9_i) >> 32) + 0x0u + cf_; t23 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_9_i) + t37; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_9_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_9_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_9_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_9_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_9_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_9_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_9_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_9_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_9_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_9_i) + t47 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_9_i) >> 32) + t48 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_9_i) + t49 + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_9_i) >> 32) + t50 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t51 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_10_i) + t10; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_10_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_10_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_10_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_10_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_10_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_10_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_10_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_10_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_10_i) >> 32) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_10_i) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_10_i) >> 32) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_10_i) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_10_i) >> 32) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t24 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_10_i) + t39; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_10_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_10_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_10_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_10_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_10_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_10_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_10_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_10_i) + t47 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_10_i) >> 32) + t48 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_10_i) + t49 + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_10_i) >> 32) + t50 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_10_i) + t51 + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_10_i) >> 32) + 0x0u + cf_; t52 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_11_i) + t12; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_11_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_11_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_11_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_11_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_11_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_11_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_11_i) >> 32) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_11_i) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_11_i) >> 32) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_11_i) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_11_i) >> 32) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_11_i) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_11_i) >> 32) + 0x0u + cf_; t25 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_11_i) + t39; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_11_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_11_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_11_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_11_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_11_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_11_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_11_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_11_i) + t47 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_11_i) >> 32) + t48 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_11_i) + t49 + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_11_i) >> 32) + t50 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_11_i) + t51 + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_11_i) >> 32) + t52 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t53 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_12_i) + t12; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_12_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_12_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_12_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_12_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_12_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_12_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_12_i) >> 32) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_12_i) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_12_i) >> 32) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_12_i) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_12_i) >> 32) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_12_i) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_12_i) >> 32) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t26 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_12_i) + t41; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_12_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_12_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_12_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_12_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_12_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_12_i) + t47 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_12_i) >> 32) + t48 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_12_i) + t49 + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_12_i) >> 32) + t50 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_12_i) + t51 + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_12_i) >> 32) + t52 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_12_i) + t53 + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_12_i) >> 32) + 0x0u + cf_; t54 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * b_13_i) + t14; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * b_13_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * b_13_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * b_13_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * b_13_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * b_13_i) >> 32) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * b_13_i) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * b_13_i) >> 32) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * b_13_i) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * b_13_i) >> 32) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * b_13_i) + t24 + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * b_13_i) >> 32) + t25 + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * b_13_i) + t26 + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * b_13_i) >> 32) + 0x0u + cf_; t27 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_13_i) + t41; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_13_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_13_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_13_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_13_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_13_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_13_i) + t47 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_13_i) >> 32) + t48 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_13_i) + t49 + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_13_i) >> 32) + t50 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_13_i) + t51 + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_13_i) >> 32) + t52 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_13_i) + t53 + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_13_i) >> 32) + t54 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t55 = (uint32_t)w_;
    w_ = (uint64_t)t1 + t29; t56 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t30 + cf_; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t31 + cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t32 + cf_; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t33 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t34 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t35 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t8 + t36 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t37 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t38 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t39 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t40 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t41 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + t42 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + t43 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t16 + t44 + cf_; t71 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t17 + t45 + cf_; t72 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t18 + t46 + cf_; t73 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t19 + t47 + cf_; t74 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t20 + t48 + cf_; t75 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t21 + t49 + cf_; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t22 + t50 + cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t23 + t51 + cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t24 + t52 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t25 + t53 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t26 + t54 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t27 + t55 + cf_; t82 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t69; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t56 + t70 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t57 + t71 + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t58 + t72 + cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t59 + t73 + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t60 + t74 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t61 + t75 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t62 + t69 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + t70 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t64 + t71 + cf_; t93 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t65 + t72 + cf_; t94 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t66 + t73 + cf_; t95 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t67 + t74 + cf_; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t68 + t75 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t83 = (uint32_t)w_;
    w_ = (uint64_t)t84 + t76; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t85 + t77 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t86 + t78 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t87 + t79 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t88 + t80 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t89 + t81 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t90 + t82 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t91 + t76 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t92 + t77 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t93 + t78 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t94 + t79 + cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t95 + t80 + cf_; t109 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t96 + t81 + cf_; t110 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t97 + t82 + cf_; t111 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t83 + 0x0u + cf_; t112 = (uint32_t)w_;
    w_ = (uint64_t)t105 + t76; t113 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t106 + t77 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t107 + t78 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t108 + t79 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t109 + t80 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t110 + t81 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t111 + t82 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t112 + 0x0u + cf_; t120 = (uint32_t)w_;
    w_ = (uint64_t)t98 + t120; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t99 + 0x0u + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t100 + 0x0u + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t101 + 0x0u + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t102 + 0x0u + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t103 + 0x0u + cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t104 + 0x0u + cf_; t127 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t113 + t120 + cf_; t128 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t114 + 0x0u + cf_; t129 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t115 + 0x0u + cf_; t130 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t116 + 0x0u + cf_; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t117 + 0x0u + cf_; t132 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t118 + 0x0u + cf_; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t119 + 0x0u + cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t135 = (uint32_t)w_;
    w_ = (uint64_t)t121 + t135; t136 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t122 + 0x0u + cf_; t137 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t123 + 0x0u + cf_; t138 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t124 + 0x0u + cf_; t139 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t125 + 0x0u + cf_; t140 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t126 + 0x0u + cf_; t141 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t127 + 0x0u + cf_; t142 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t128 + t135 + cf_; t143 = (uint32_t)w_;
    r[0] = t136;
    r[1] = t137;
    r[2] = t138;
    r[3] = t139;
    r[4] = t140;
    r[5] = t141;
    r[6] = t142;
    r[7] = t143;
    r[8] = t129;
    r[9] = t130;
    r[10] = t131;
    r[11] = t132;
    r[12] = t133;
    r[13] = t134;
#endif
  }

  // c = a*a (pseudo.py:663-702 / monty.py:982-1165)
  static MAB_DEV void sqr(uint32_t (&r)[14], const uint32_t (&a)[14]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<198>;\n\t"
        "mul.lo.u32 t2, %14, %16;\n\t"
        "mul.hi.u32 t3, %14, %16;\n\t"
        "mul.lo.u32 t4, %14, %18;\n\t"
        "mul.hi.u32 t5, %14, %18;\n\t"
        "mul.lo.u32 t6, %14, %20;\n\t"
        "mul.hi.u32 t7, %14, %20;\n\t"
        "mul.lo.u32 t8, %14, %22;\n\t"
        "mul.hi.u32 t9, %14, %22;\n\t"
        "mul.lo.u32 t10, %14, %24;\n\t"
        "mul.hi.u32 t11, %14, %24;\n\t"
        "mul.lo.u32 t12, %14, %26;\n\t"
        "mul.hi.u32 t13, %14, %26;\n\t"
        "mul.lo.u32 t29, %14, %15;\n\t"
        "mul.hi.u32 t30, %14, %15;\n\t"
        "mul.lo.u32 t31, %14, %17;\n\t"
        "mul.hi.u32 t32, %14, %17;\n\t"
        "mul.lo.u32 t33, %14, %19;\n\t"
        "mul.hi.u32 t34, %14, %19;\n\t"
        "mul.lo.u32 t35, %14, %21;\n\t"
        "mul.hi.u32 t36, %14, %21;\n\t"
        "mul.lo.u32 t37, %14, %23;\n\t"
        "mul.hi.u32 t38, %14, %23;\n\t"
        "mul.lo.u32 t39, %14, %25;\n\t"
        "mul.hi.u32 t40, %14, %25;\n\t"
        "mul.lo.u32 t41, %14, %27;\n\t"
        "mul.hi.u32 t42, %14, %27;\n\t"
        "mad.lo.cc.u32 t4, %15, %17, t4;\n\t"
        "madc.hi.cc.u32 t5, %15, %17, t5;\n\t"
        "madc.lo.cc.u32 t6, %15, %19, t6;\n\t"
        "madc.hi.cc.u32 t7, %15, %19, t7;\n\t"
        "madc.lo.cc.u32 t8, %15, %21, t8;\n\t"
        "madc.hi.cc.u32 t9, %15, %21, t9;\n\t"
        "madc.lo.cc.u32 t10, %15, %23, t10;\n\t"
        "madc.hi.cc.u32 t11, %15, %23, t11;\n\t"
        "madc.lo.cc.u32 t12, %15, %25, t12;\n\t"
        "madc.hi.cc.u32 t13, %15, %25, t13;\n\t"
        "madc.lo.cc.u32 t14, %15, %27, 0x0;\n\t"
        "madc.hi.u32 t15, %15, %27, 0x0;\n\t"
        "mad.lo.cc.u32 t31, %15, %16, t31;\n\t"
        "madc.hi.cc.u32 t32, %15, %16, t32;\n\t"
        "madc.lo.cc.u32 t33, %15, %18, t33;\n\t"
        "madc.hi.cc.u32 t34, %15, %18, t34;\n\t"
        "madc.lo.cc.u32 t35, %15, %20, t35;\n\t"
        "madc.hi.cc.u32 t36, %15, %20, t36;\n\t"
        "madc.lo.cc.u32 t37, %15, %22, t37;\n\t"
        "madc.hi.cc.u32 t38, %15, %22, t38;\n\t"
        "madc.lo.cc.u32 t39, %15, %24, t39;\n\t"
        "madc.hi.cc.u32 t40, %15, %24, t40;\n\t"
        "madc.lo.cc.u32 t41, %15, %26, t41;\n\t"
        "madc.hi.cc.u32 t42, %15, %26, t42;\n\t"
        "addc.u32 t43, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t6, %16, %18, t6;\n\t"
        "madc.hi.cc.u32 t7, %16, %18, t7;\n\t"
        "madc.lo.cc.u32 t8, %16, %20, t8;\n\t"
        "madc.hi.cc.u32 t9, %16, %20, t9;\n\t"
        "madc.lo.cc.u32 t10, %16, %22, t10;\n\t"
        "madc.hi.cc.u32 t11, %16, %22, t11;\n\t"
        "madc.lo.cc.u32 t12, %16, %24, t12;\n\t"
        "madc.hi.cc.u32 t13, %16, %24, t13;\n\t"
        "madc.lo.cc.u32 t14, %16, %26, t14;\n\t"
        "madc.hi.cc.u32 t15, %16, %26, t15;\n\t"
        "addc.u32 t16, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t33, %16, %17, t33;\n\t"
        "madc.hi.cc.u32 t34, %16, %17, t34;\n\t"
        "madc.lo.cc.u32 t35, %16, %19, t35;\n\t"
        "madc.hi.cc.u32 t36, %16, %19, t36;\n\t"
        "madc.lo.cc.u32 t37, %16, %21, t37;\n\t"
        "madc.hi.cc.u32 t38, %16, %21, t38;\n\t"
        "madc.lo.cc.u32 t39, %16, %23, t39;\n\t"
        "madc.hi.cc.u32 t40, %16, %23, t40;\n\t"
        "madc.lo.cc.u32 t41, %16, %25, t41;\n\t"
        "madc.hi.cc.u32 t42, %16, %25, t42;\n\t"
        "madc.lo.cc.u32 t43, %16, %27, t43;\n\t"
        "madc.hi.u32 t44, %16, %27, 0x0;\n\t"
        "mad.lo.cc.u32 t8, %17, %19, t8;\n\t"
        "madc.hi.cc.u32 t9, %17, %19, t9;\n\t"
        "madc.lo.cc.u32 t10, %17, %21, t10;\n\t"
        "madc.hi.cc.u32 t11, %17, %21, t11;\n\t"
        "madc.lo.cc.u32 t12, %17, %23, t12;\n\t"
        "madc.hi.cc.u32 t13, %17, %23, t13;\n\t"
        "madc.lo.cc.u32 t14, %17, %25, t14;\n\t"
        "madc.hi.cc.u32 t15, %17, %25, t15;\n\t"
        "madc.lo.cc.u32 t16, %17, %27, t16;\n\t"
        "madc.hi.u32 t17, %17, %27, 0x0;\n\t"
        "mad.lo.cc.u32 t35, %17, %18, t35;\n\t"
        "madc.hi.cc.u32 t36, %17, %18, t36;\n\t"
        "madc.lo.cc.u32 t37, %17, %20, t37;\n\t"
        "madc.hi.cc.u32 t38, %17, %20, t38;\n\t"
        "madc.lo.cc.u32 t39, %17, %22, t39;\n\t"
        "madc.hi.cc.u32 t40, %17, %22, t40;\n\t"
        "madc.lo.cc.u32 t41, %17, %24, t41;\n\t"
        "madc.hi.cc.u32 t42, %17, %24, t42;\n\t"
        "madc.lo.cc.u32 t43, %17, %26, t43;\n\t"
        "madc.hi.cc.u32 t44, %17, %26, t44;\n\t"
        "addc.u32 t45, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t10, %18, %20, t10;\n\t"
        "madc.hi.cc.u32 t11, %18, %20, t11;\n\t"
        "madc.lo.cc.u32 t12, %18, %22, t12;\n\t"
        "madc.hi.cc.u32 t13, %18, %22, t13;\n\t"
        "madc.lo.cc.u32 t14, %18, %24, t14;\n\t"
        "madc.hi.cc.u32 t15, %18, %24, t15;\n\t"
        "madc.lo.cc.u32 t16, %18, %26, t16;\n\t"
        "madc.hi.cc.u32 t17, %18, %26, t17;\n\t"
        "addc.u32 t18, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t37, %18, %19, t37;\n\t"
        "madc.hi.cc.u32 t38, %18, %19, t38;\n\t"
        "madc.lo.cc.u32 t39, %18, %21, t39;\n\t"
        "madc.hi.cc.u32 t40, %18, %21, t40;\n\t"
        "madc.lo.cc.u32 t41, %18, %23, t41;\n\t"
        "madc.hi.cc.u32 t42, %18, %23, t42;\n\t"
        "madc.lo.cc.u32 t43, %18, %25, t43;\n\t"
        "madc.hi.cc.u32 t44, %18, %25, t44;\n\t"
        "madc.lo.cc.u32 t45, %18, %27, t45;\n\t"
        "madc.hi.u32 t46, %18, %27, 0x0;\n\t"
        "mad.lo.cc.u32 t12, %19, %21, t12;\n\t"
        "madc.hi.cc.u32 t13, %19, %21, t13;\n\t"
        "madc.lo.cc.u32 t14, %19, %23, t14;\n\t"
        "madc.hi.cc.u32 t15, %19, %23, t15;\n\t"
        "madc.lo.cc.u32 t16, %19, %25, t16;\n\t"
        "madc.hi.cc.u32 t17, %19, %25, t17;\n\t"
        "madc.lo.cc.u32 t18, %19, %27, t18;\n\t"
        "madc.hi.u32 t19, %19, %27, 0x0;\n\t"
        "mad.lo.cc.u32 t39, %19, %20, t39;\n\t"
        "madc.hi.cc.u32 t40, %19, %20, t40;\n\t"
        "madc.lo.cc.u32 t41, %19, %22, t41;\n\t"
        "madc.hi.cc.u32 t42, %19, %22, t42;\n\t"
        "madc.lo.cc.u32 t43, %19, %24, t43;\n\t"
        "madc.hi.cc.u32 t44, %19, %24, t44;\n\t"
        "madc.lo.cc.u32 t45, %19, %26, t45;\n\t"
        "madc.hi.cc.u32 t46, %19, %26, t46;\n\t"
        "addc.u32 t47, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t14, %20, %22, t14;\n\t"
        "madc.hi.cc.u32 t15, %20, %22, t15;\n\t"
        "madc.lo.cc.u32 t16, %20, %24, t16;\n\t"
        "madc.hi.cc.u32 t17, %20, %24, t17;\n\t"
        "madc.lo.cc.u32 t18, %20, %26, t18;\n\t"
        "madc.hi.cc.u32 t19, %20, %26, t19;\n\t"
        "addc.u32 t20, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t41, %20, %21, t41;\n\t"
        "madc.hi.cc.u32 t42, %20, %21, t42;\n\t"
        "madc.lo.cc.u32 t43, %20, %23, t43;\n\t"
        "madc.hi.cc.u32 t44, %20, %23, t44;\n\t"
        "madc.lo.cc.u32 t45, %20, %25, t45;\n\t"
        "madc.hi.cc.u32 t46, %20, %25, t46;\n\t"
        "madc.lo.cc.u32 t47, %20, %27, t47;\n\t"
        "madc.hi.u32 t48, %20, %27, 0x0;\n\t"
        "mad.lo.cc.u32 t16, %21, %23, t16;\n\t"
        "madc.hi.cc.u32 t17, %21, %23, t17;\n\t"
        "madc.lo.cc.u32 t18, %21, %25, t18;\n\t"
        "madc.hi.cc.u32 t19, %21, %25, t19;\n\t"
        "madc.lo.cc.u32 t20, %21, %27, t20;\n\t"
        "madc.hi.u32 t21, %21, %27, 0x0;\n\t"
        "mad.lo.cc.u32 t43, %21, %22, t43;\n\t"
        "madc.hi.cc.u32 t44, %21, %22, t44;\n\t"
        "madc.lo.cc.u32 t45, %21, %24, t45;\n\t"
        "madc.hi.cc.u32 t46, %21, %24, t46;\n\t"
        "madc.lo.cc.u32 t47, %21, %26, t47;\n\t"
        "madc.hi.cc.u32 t48, %21, %26, t48;\n\t"
        "addc.u32 t49, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t18, %22, %24, t18;\n\t"
        "madc.hi.cc.u32 t19, %22, %24, t19;\n\t"
        "madc.lo.cc.u32 t20, %22, %26, t20;\n\t"
        "madc.hi.cc.u32 t21, %22, %26, t21;\n\t"
        "addc.u32 t22, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t45, %22, %23, t45;\n\t"
        "madc.hi.cc.u32 t46, %22, %23, t46;\n\t"
        "madc.lo.cc.u32 t47, %22, %25, t47;\n\t"
        "madc.hi.cc.u32 t48, %22, %25, t48;\n\t"
        "madc.lo.cc.u32 t49, %22, %27, t49;\n\t"
        "madc.hi.u32 t50, %22, %27, 0x0;\n\t"
        "mad.lo.cc.u32 t20, %23, %25, t20;\n\t"
        "madc.hi.cc.u32 t21, %23, %25, t21;\n\t"
        "madc.lo.cc.u32 t22, %23, %27, t22;\n\t"
        "madc.hi.u32 t23, %23, %27, 0x0;\n\t"
        "mad.lo.cc.u32 t47, %23, %24, t47;\n\t"
        "madc.hi.cc.u32 t48, %23, %24, t48;\n\t"
        "madc.lo.cc.u32 t49, %23, %26, t49;\n\t"
        "madc.hi.cc.u32 t50, %23, %26, t50;\n\t"
        "addc.u32 t51, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t22, %24, %26, t22;\n\t"
        "madc.hi.cc.u32 t23, %24, %26, t23;\n\t"
        "addc.u32 t24, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t49, %24, %25, t49;\n\t"
        "madc.hi.cc.u32 t50, %24, %25, t50;\n\t"
        "madc.lo.cc.u32 t51, %24, %27, t51;\n\t"
        "madc.hi.u32 t52, %24, %27, 0x0;\n\t"
        "mad.lo.cc.u32 t24, %25, %27, t24;\n\t"
        "madc.hi.u32 t25, %25, %27, 0x0;\n\t"
        "mad.lo.cc.u32 t51, %25, %26, t51;\n\t"
        "madc.hi.cc.u32 t52, %25, %26, t52;\n\t"
        "addc.u32 t53, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t53, %26, %27, t53;\n\t"
        "madc.hi.u32 t54, %26, %27, 0x0;\n\t"
        "add.cc.u32 t56, t2, t30;\n\t"
        "addc.cc.u32 t57, t3, t31;\n\t"
        "addc.cc.u32 t58, t4, t32;\n\t"
        "addc.cc.u32 t59, t5, t33;\n\t"
        "addc.cc.u32 t60, t6, t34;\n\t"
        "addc.cc.u32 t61, t7, t35;\n\t"
        "addc.cc.u32 t62, t8, t36;\n\t"
        "addc.cc.u32 t63, t9, t37;\n\t"
        "addc.cc.u32 t64, t10, t38;\n\t"
        "addc.cc.u32 t65, t11, t39;\n\t"
        "addc.cc.u32 t66, t12, t40;\n\t"
        "addc.cc.u32 t67, t13, t41;\n\t"
        "addc.cc.u32 t68, t14, t42;\n\t"
        "addc.cc.u32 t69, t15, t43;\n\t"
        "addc.cc.u32 t70, t16, t44;\n\t"
        "addc.cc.u32 t71, t17, t45;\n\t"
        "addc.cc.u32 t72, t18, t46;\n\t"
        "addc.cc.u32 t73, t19, t47;\n\t"
        "addc.cc.u32 t74, t20, t48;\n\t"
        "addc.cc.u32 t75, t21, t49;\n\t"
        "addc.cc.u32 t76, t22, t50;\n\t"
        "addc.cc.u32 t77, t23, t51;\n\t"
        "addc.cc.u32 t78, t24, t52;\n\t"
        "addc.cc.u32 t79, t25, t53;\n\t"
        "addc.cc.u32 t80, 0x0, t54;\n\t"
        "addc.u32 t81, 0x0, 0x0;\n\t"
        "shl.b32 t82, t29, 1;\n\t"
        "shf.l.wrap.b32 t83, t29, t56, 1;\n\t"
        "shf.l.wrap.b32 t84, t56, t57, 1;\n\t"
        "shf.l.wrap.b32 t85, t57, t58, 1;\n\t"
        "shf.l.wrap.b32 t86, t58, t59, 1;\n\t"
        "shf.l.wrap.b32 t87, t59, t60, 1;\n\t"
        "shf.l.wrap.b32 t88, t60, t61, 1;\n\t"
        "shf.l.wrap.b32 t89, t61, t62, 1;\n\t"
        "shf.l.wrap.b32 t90, t62, t63, 1;\n\t"
        "shf.l.wrap.b32 t91, t63, t64, 1;\n\t"
        "shf.l.wrap.b32 t92, t64, t65, 1;\n\t"
        "shf.l.wrap.b32 t93, t65, t66, 1;\n\t"
        "shf.l.wrap.b32 t94, t66, t67, 1;\n\t"
        "shf.l.wrap.b32 t95, t67, t68, 1;\n\t"
        "shf.l.wrap.b32 t96, t68, t69, 1;\n\t"
        "shf.l.wrap.b32 t97, t69, t70, 1;\n\t"
        "shf.l.wrap.b32 t98, t70, t71, 1;\n\t"
        "shf.l.wrap.b32 t99, t71, t72, 1;\n\t"
        "shf.l.wrap.b32 t100, t72, t73, 1;\n\t"
        "shf.l.wrap.b32 t101, t73, t74, 1;\n\t"
        "shf.l.wrap.b32 t102, t74, t75, 1;\n\t"
        "shf.l.wrap.b32 t103, t75, t76, 1;\n\t"
        "shf.l.wrap.b32 t104, t76, t77, 1;\n\t"
        "shf.l.wrap.b32 t105, t77, t78, 1;\n\t"
        "shf.l.wrap.b32 t106, t78, t79, 1;\n\t"
        "shf.l.wrap.b32 t107, t79, t80, 1;\n\t"
        "shf.l.wrap.b32 t108, t80, t81, 1;\n\t"
        "mad.lo.cc.u32 t109, %14, %14, 0x0;\n\t"
        "madc.hi.cc.u32 t110, %14, %14, t82;\n\t"
        "madc.lo.cc.u32 t111, %15, %15, t83;\n\t"
        "madc.hi.cc.u32 t112, %15, %15, t84;\n\t"
        "madc.lo.cc.u32 t113, %16, %16, t85;\n\t"
        "madc.hi.cc.u32 t114, %16, %16, t86;\n\t"
        "madc.lo.cc.u32 t115, %17, %17, t87;\n\t"
        "madc.hi.cc.u32 t116, %17, %17, t88;\n\t"
        "madc.lo.cc.u32 t117, %18, %18, t89;\n\t"
        "madc.hi.cc.u32 t118, %18, %18, t90;\n\t"
        "madc.lo.cc.u32 t119, %19, %19, t91;\n\t"
        "madc.hi.cc.u32 t120, %19, %19, t92;\n\t"
        "madc.lo.cc.u32 t121, %20, %20, t93;\n\t"
        "madc.hi.cc.u32 t122, %20, %20, t94;\n\t"
        "madc.lo.cc.u32 t123, %21, %21, t95;\n\t"
        "madc.hi.cc.u32 t124, %21, %21, t96;\n\t"
        "madc.lo.cc.u32 t125, %22, %22, t97;\n\t"
        "madc.hi.cc.u32 t126, %22, %22, t98;\n\t"
        "madc.lo.cc.u32 t127, %23, %23, t99;\n\t"
        "madc.hi.cc.u32 t128, %23, %23, t100;\n\t"
        "madc.lo.cc.u32 t129, %24, %24, t101;\n\t"
        "madc.hi.cc.u32 t130, %24, %24, t102;\n\t"
        "madc.lo.cc.u32 t131, %25, %25, t103;\n\t"
        "madc.hi.cc.u32 t132, %25, %25, t104;\n\t"
        "madc.lo.cc.u32 t133, %26, %26, t105;\n\t"
        "madc.hi.cc.u32 t134, %26, %26, t106;\n\t"
        "madc.lo.cc.u32 t135, %27, %27, t107;\n\t"
        "madc.hi.u32 t136, %27, %27, t108;\n\t"
        "add.cc.u32 t138, t109, t123;\n\t"
        "addc.cc.u32 t139, t110, t124;\n\t"
        "addc.cc.u32 t140, t111, t125;\n\t"
        "addc.cc.u32 t141, t112, t126;\n\t"
        "addc.cc.u32 t142, t113, t127;\n\t"
        "addc.cc.u32 t143, t114, t128;\n\t"
        "addc.cc.u32 t144, t115, t129;\n\t"
        "addc.cc.u32 t145, t116, t123;\n\t"
        "addc.cc.u32 t146, t117, t124;\n\t"
        "addc.cc.u32 t147, t118, t125;\n\t"
        "addc.cc.u32 t148, t119, t126;\n\t"
        "addc.cc.u32 t149, t120, t127;\n\t"
        "addc.cc.u32 t150, t121, t128;\n\t"
        "addc.cc.u32 t151, t122, t129;\n\t"
        "addc.u32 t137, 0x0, 0x0;\n\t"
        "add.cc.u32 t152, t138, t130;\n\t"
        "addc.cc.u32 t153, t139, t131;\n\t"
        "addc.cc.u32 t154, t140, t132;\n\t"
        "addc.cc.u32 t155, t141, t133;\n\t"
        "addc.cc.u32 t156, t142, t134;\n\t"
        "addc.cc.u32 t157, t143, t135;\n\t"
        "addc.cc.u32 t158, t144, t136;\n\t"
        "addc.cc.u32 t159, t145, t130;\n\t"
        "addc.cc.u32 t160, t146, t131;\n\t"
        "addc.cc.u32 t161, t147, t132;\n\t"
        "addc.cc.u32 t162, t148, t133;\n\t"
        "addc.cc.u32 t163, t149, t134;\n\t"
        "addc.cc.u32 t164, t150, t135;\n\t"
        "addc.cc.u32 t165, t151, t136;\n\t"
        "addc.u32 t166, t137, 0x0;\n\t"
        "add.cc.u32 t167, t159, t130;\n\t"
        "addc.cc.u32 t168, t160, t131;\n\t"
        "addc.cc.u32 t169, t161, t132;\n\t"
        "addc.cc.u32 t170, t162, t133;\n\t"
        "addc.cc.u32 t171, t163, t134;\n\t"
        "addc.cc.u32 t172, t164, t135;\n\t"
        "addc.cc.u32 t173, t165, t136;\n\t"
        "addc.u32 t174, t166, 0x0;\n\t"
        "add.cc.u32 t175, t152, t174;\n\t"
        "addc.cc.u32 t176, t153, 0x0;\n\t"
        "addc.cc.u32 t177, t154, 0x0;\n\t"
        "addc.cc.u32 t178, t155, 0x0;\n\t"
        "addc.cc.u32 t179, t156, 0x0;\n\t"
        "addc.cc.u32 t180, t157, 0x0;\n\t"
        "addc.cc.u32 t181, t158, 0x0;\n\t"
        "addc.cc.u32 t182, t167, t174;\n\t"
        "addc.cc.u32 t183, t168, 0x0;\n\t"
        "addc.cc.u32 t184, t169, 0x0;\n\t"
        "addc.cc.u32 t185, t170, 0x0;\n\t"
        "addc.cc.u32 t186, t171, 0x0;\n\t"
        "addc.cc.u32 t187, t172, 0x0;\n\t"
        "addc.cc.u32 t188, t173, 0x0;\n\t"
        "addc.u32 t189, 0x0, 0x0;\n\t"
        "add.cc.u32 t190, t175, t189;\n\t"
        "addc.cc.u32 t191, t176, 0x0;\n\t"
        "addc.cc.u32 t192, t177, 0x0;\n\t"
        "addc.cc.u32 t193, t178, 0x0;\n\t"
        "addc.cc.u32 t194, t179, 0x0;\n\t"
        "addc.cc.u32 t195, t180, 0x0;\n\t"
        "addc.cc.u32 t196, t181, 0x0;\n\t"
        "addc.u32 t197, t182, t189;\n\t"
        "mov.u32 %0, t190;\n\t"
        "mov.u32 %1, t191;\n\t"
        "mov.u32 %2, t192;\n\t"
        "mov.u32 %3, t193;\n\t"
        "mov.u32 %4, t194;\n\t"
        "mov.u32 %5, t195;\n\t"
        "mov.u32 %6, t196;\n\t"
        "mov.u32 %7, t197;\n\t"
        "mov.u32 %8, t183;\n\t"
        "mov.u32 %9, t184;\n\t"
        "mov.u32 %10, t185;\n\t"
        "mov.u32 %11, t186;\n\t"
        "mov.u32 %12, t187;\n\t"
        "mov.u32 %13, t188;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t a_8_i = a[8];
    const uint32_t a_9_i = a[9];
    const uint32_t a_10_i = a[10];
    const uint32_t a_11_i = a[11];
    const uint32_t a_12_i = a[12];
    const uint32_t a_13_i = a[13];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127, t128, t129, t130, t131, t132, t133, t134, t135, t136, t137, t138, t139, t140, t141, t142, t143, t144, t145, t146, t147, t148, t149, t150, t151, t152, t153, t154, t155, t156, t157, t158, t159, t160, t161, t162, t163, t164, t165, t166, t167, t168, t169, t170, t171, t172, t173, t174, t175, t176, t177, t178, t179, t180, t181, t182, t183, t184, t185, t186, t187, t188, t189, t190, t191, t192, t193, t194, t195, t196, t197;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t2 = (uint32_t)((uint32_t)(a_0_i * a_2_i));
    t3 = (uint32_t)(((uint64_t)a_0_i * a_2_i) >> 32);
    t4 = (uint32_t)((uint32_t)(a_0_i * a_4_i));
    t5 = (uint32_t)(((uint64_t)a_0_i * a_4_i) >> 32);
    t6 = (uint32_t)((uint32_t)(a_0_i * a_6_i));
    t7 = (uint32_t)(((uint64_t)a_0_i * a_6_i) >> 32);
    t8 = (uint32_t)((uint32_t)(a_0_i * a_8_i));
    t9 = (uint32_t)(((uint64_t)a_0_i * a_8_i) >> 32);
    t10 = (uint32_t)((uint32_t)(a_0_i * a_10_i));
    t11 = (uint32_t)(((uint64_t)a_0_i * a_10_i) >> 32);
    t12 = (uint32_t)((uint32_t)(a_0_i * a_12_i));
    t13 = (uint32_t)(((uint64_t)a_0_i * a_12_i) >> 32);
    t29 = (uint32_t)((uint32_t)(a_0_i * a_1_i));
    t30 = (uint32_t)(((uint64_t)a_0_i * a_1_i) >> 32);
    t31 = (uint32_t)((uint32_t)(a_0_i * a_3_i));
    t32 = (uint32_t)(((uint64_t)a_0_i * a_3_i) >> 32);
    t33 = (uint32_t)((uint32_t)(a_0_i * a_5_i));
    t34 = (uint32_t)(((uint64_t)a_0_i * a_5_i) >> 32);
    t35 = (uint32_t)((uint32_t)(a_0_i * a_7_i));
    t36 = (uint32_t)(((uint64_t)a_0_i * a_7_i) >> 32);
    t37 = (uint32_t)((uint32_t)(a_0_i * a_9_i));
    t38 = (uint32_t)(((uint64_t)a_0_i * a_9_i) >> 32);
    t39 = (uint32_t)((uint32_t)(a_0_i * a_11_i));
    t40 = (uint32_t)(((uint64_t)a_0_i * a_11_i) >> 32);
    t41 = (uint32_t)((uint32_t)(a_0_i * a_13_i));
    t42 = (uint32_t)(((uint64_t)a_0_i * a_13_i) >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_3_i) + t4; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_3_i) >> 32) + t5 + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_5_i) + t6 + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_5_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_7_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_7_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_9_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_9_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_11_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_11_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_13_i) + 0x0u + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_13_i) >> 32) + 0x0u + cf_; t15 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * a_2_i) + t31; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_2_i) >> 32) + t32 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_4_i) + t33 + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_4_i) >> 32) + t34 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_6_i) + t35 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_6_i) >> 32) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_8_i) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_8_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_10_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_10_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_12_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_12_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t43 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_2_i * a_4_i) + t6; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_4_i) >> 32) + t7 + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_6_i) + t8 + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_6_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_8_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_8_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_10_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_10_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_12_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_12_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t16 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_2_i * a_3_i) + t33; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_3_i) >> 32) + t34 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_5_i) + t35 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_5_i) >> 32) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_7_i) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_7_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_9_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_9_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_11_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_11_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_13_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_13_i) >> 32) + 0x0u + cf_; t44 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_3_i * a_5_i) + t8; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_5_i) >> 32) + t9 + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_7_i) + t10 + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_7_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_9_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_9_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_11_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_11_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_13_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_13_i) >> 32) + 0x0u + cf_; t17 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_3_i * a_4_i) + t35; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_4_i) >> 32) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_6_i) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_6_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_8_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_8_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_10_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_10_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_12_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_12_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t45 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_4_i * a_6_i) + t10; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_6_i) >> 32) + t11 + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_8_i) + t12 + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_8_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_10_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_10_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_12_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_12_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t18 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_4_i * a_5_i) + t37; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_5_i) >> 32) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_7_i) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_7_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_9_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_9_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_11_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_11_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_13_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_13_i) >> 32) + 0x0u + cf_; t46 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_5_i * a_7_i) + t12; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_7_i) >> 32) + t13 + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_9_i) + t14 + cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_9_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_11_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_11_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_13_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_13_i) >> 32) + 0x0u + cf_; t19 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_5_i * a_6_i) + t39; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_6_i) >> 32) + t40 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_8_i) + t41 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_8_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_10_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_10_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_12_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_12_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t47 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_6_i * a_8_i) + t14; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_8_i) >> 32) + t15 + cf_; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_10_i) + t16 + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_10_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_12_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_12_i) >> 32) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t20 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_6_i * a_7_i) + t41; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_7_i) >> 32) + t42 + cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_9_i) + t43 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_9_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_11_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_11_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_13_i) + t47 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_13_i) >> 32) + 0x0u + cf_; t48 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_7_i * a_9_i) + t16; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_9_i) >> 32) + t17 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_11_i) + t18 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_11_i) >> 32) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_13_i) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_13_i) >> 32) + 0x0u + cf_; t21 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_7_i * a_8_i) + t43; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_8_i) >> 32) + t44 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_10_i) + t45 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_10_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_12_i) + t47 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_12_i) >> 32) + t48 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t49 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_8_i * a_10_i) + t18; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * a_10_i) >> 32) + t19 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * a_12_i) + t20 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * a_12_i) >> 32) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t22 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_8_i * a_9_i) + t45; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * a_9_i) >> 32) + t46 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * a_11_i) + t47 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * a_11_i) >> 32) + t48 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * a_13_i) + t49 + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * a_13_i) >> 32) + 0x0u + cf_; t50 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_9_i * a_11_i) + t20; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * a_11_i) >> 32) + t21 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * a_13_i) + t22 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * a_13_i) >> 32) + 0x0u + cf_; t23 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_9_i * a_10_i) + t47; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * a_10_i) >> 32) + t48 + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * a_12_i) + t49 + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * a_12_i) >> 32) + t50 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t51 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_10_i * a_12_i) + t22; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * a_12_i) >> 32) + t23 + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t24 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_10_i * a_11_i) + t49; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * a_11_i) >> 32) + t50 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * a_13_i) + t51 + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * a_13_i) >> 32) + 0x0u + cf_; t52 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_11_i * a_13_i) + t24; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * a_13_i) >> 32) + 0x0u + cf_; t25 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_11_i * a_12_i) + t51; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * a_12_i) >> 32) + t52 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t53 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_12_i * a_13_i) + t53; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * a_13_i) >> 32) + 0x0u + cf_; t54 = (uint32_t)w_;
    w_ = (uint64_t)t2 + t30; t56 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t31 + cf_; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t32 + cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t33 + cf_; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t34 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t35 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t8 + t36 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t37 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t38 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t39 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t40 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t41 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + t42 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + t43 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t16 + t44 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t17 + t45 + cf_; t71 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t18 + t46 + cf_; t72 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t19 + t47 + cf_; t73 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t20 + t48 + cf_; t74 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t21 + t49 + cf_; t75 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t22 + t50 + cf_; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t23 + t51 + cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t24 + t52 + cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t25 + t53 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t54 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t81 = (uint32_t)w_;
    t82 = (uint32_t)(t29 << 1);
    t83 = (uint32_t)(((((uint64_t)t56 << 32) | t29) << 1) >> 32);
    t84 = (uint32_t)(((((uint64_t)t57 << 32) | t56) << 1) >> 32);
    t85 = (uint32_t)(((((uint64_t)t58 << 32) | t57) << 1) >> 32);
    t86 = (uint32_t)(((((uint64_t)t59 << 32) | t58) << 1) >> 32);
    t87 = (uint32_t)(((((uint64_t)t60 << 32) | t59) << 1) >> 32);
    t88 = (uint32_t)(((((uint64_t)t61 << 32) | t60) << 1) >> 32);
    t89 = (uint32_t)(((((uint64_t)t62 << 32) | t61) << 1) >> 32);
    t90 = (uint32_t)(((((uint64_t)t63 << 32) | t62) << 1) >> 32);
    t91 = (uint32_t)(((((uint64_t)t64 << 32) | t63) << 1) >> 32);
    t92 = (uint32_t)(((((uint64_t)t65 << 32) | t64) << 1) >> 32);
    t93 = (uint32_t)(((((uint64_t)t66 << 32) | t65) << 1) >> 32);
    t94 = (uint32_t)(((((uint64_t)t67 << 32) | t66) << 1) >> 32);
    t95 = (uint32_t)(((((uint64_t)t68 << 32) | t67) << 1) >> 32);
    t96 = (uint32_t)(((((uint64_t)t69 << 32) | t68) << 1) >> 32);
    t97 = (uint32_t)(((((uint64_t)t70 << 32) | t69) << 1) >> 32);
    t98 = (uint32_t)(((((uint64_t)t71 << 32) | t70) << 1) >> 32);
    t99 = (uint32_t)(((((uint64_t)t72 << 32) | t71) << 1) >> 32);
    t100 = (uint32_t)(((((uint64_t)t73 << 32) | t72) << 1) >> 32);
    t101 = (uint32_t)(((((uint64_t)t74 << 32) | t73) << 1) >> 32);
    t102 = (uint32_t)(((((uint64_t)t75 << 32) | t74) << 1) >> 32);
    t103 = (uint32_t)(((((uint64_t)t76 << 32) | t75) << 1) >> 32);
    t104 = (uint32_t)(((((uint64_t)t77 << 32) | t76) << 1) >> 32);
    t105 = (uint32_t)(((((uint64_t)t78 << 32) | t77) << 1) >> 32);
    t106 = (uint32_t)(((((uint64_t)t79 << 32) | t78) << 1) >> 32);
    t107 = (uint32_t)(((((uint64_t)t80 << 32) | t79) << 1) >> 32);
    t108 = (uint32_t)(((((uint64_t)t81 << 32) | t80) << 1) >> 32);
    w_ = (uint64_t)(uint32_t)(a_0_i * a_0_i) + 0x0u; t109 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * a_0_i) >> 32) + t82 + cf_; t110 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * a_1_i) + t83 + cf_; t111 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * a_1_i) >> 32) + t84 + cf_; t112 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * a_2_i) + t85 + cf_; t113 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * a_2_i) >> 32) + t86 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * a_3_i) + t87 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * a_3_i) >> 32) + t88 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * a_4_i) + t89 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * a_4_i) >> 32) + t90 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * a_5_i) + t91 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * a_5_i) >> 32) + t92 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * a_6_i) + t93 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * a_6_i) >> 32) + t94 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * a_7_i) + t95 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * a_7_i) >> 32) + t96 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * a_8_i) + t97 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * a_8_i) >> 32) + t98 + cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_9_i * a_9_i) + t99 + cf_; t127 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_9_i * a_9_i) >> 32) + t100 + cf_; t128 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * a_10_i) + t101 + cf_; t129 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * a_10_i) >> 32) + t102 + cf_; t130 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_11_i * a_11_i) + t103 + cf_; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_11_i * a_11_i) >> 32) + t104 + cf_; t132 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * a_12_i) + t105 + cf_; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * a_12_i) >> 32) + t106 + cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_13_i * a_13_i) + t107 + cf_; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_13_i * a_13_i) >> 32) + t108 + cf_; t136 = (uint32_t)w_;
    w_ = (uint64_t)t109 + t123; t138 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t110 + t124 + cf_; t139 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t111 + t125 + cf_; t140 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t112 + t126 + cf_; t141 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t113 + t127 + cf_; t142 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t114 + t128 + cf_; t143 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t115 + t129 + cf_; t144 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t116 + t123 + cf_; t145 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t117 + t124 + cf_; t146 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t118 + t125 + cf_; t147 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t119 + t126 + cf_; t148 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t120 + t127 + cf_; t149 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t121 + t128 + cf_; t150 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t122 + t129 + cf_; t151 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t137 = (uint32_t)w_;
    w_ = (uint64_t)t138 + t130; t152 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t139 + t131 + cf_; t153 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t140 + t132 + cf_; t154 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t141 + t133 + cf_; t155 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t142 + t134 + cf_; t156 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t143 + t135 + cf_; t157 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t144 + t136 + cf_; t158 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t145 + t130 + cf_; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t146 + t131 + cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t147 + t132 + cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t148 + t133 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t149 + t134 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t150 + t135 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t151 + t136 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t137 + 0x0u + cf_; t166 = (uint32_t)w_;
    w_ = (uint64_t)t159 + t130; t167 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t160 + t131 + cf_; t168 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t161 + t132 + cf_; t169 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t162 + t133 + cf_; t170 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t163 + t134 + cf_; t171 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t164 + t135 + cf_; t172 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t165 + t136 + cf_; t173 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t166 + 0x0u + cf_; t174 = (uint32_t)w_;
    w_ = (uint64_t)t152 + t174; t175 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t153 + 0x0u + cf_; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t154 + 0x0u + cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t155 + 0x0u + cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t156 + 0x0u + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t157 + 0x0u + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t158 + 0x0u + cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t167 + t174 + cf_; t182 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t168 + 0x0u + cf_; t183 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t169 + 0x0u + cf_; t184 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t170 + 0x0u + cf_; t185 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t171 + 0x0u + cf_; t186 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t172 + 0x0u + cf_; t187 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t173 + 0x0u + cf_; t188 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t189 = (uint32_t)w_;
    w_ = (uint64_t)t175 + t189; t190 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t176 + 0x0u + cf_; t191 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t177 + 0x0u + cf_; t192 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t178 + 0x0u + cf_; t193 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t179 + 0x0u + cf_; t194 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t180 + 0x0u + cf_; t195 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t181 + 0x0u + cf_; t196 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t182 + t189 + cf_; t197 = (uint32_t)w_;
    r[0] = t190;
    r[1] = t191;
    r[2] = t192;
    r[3] = t193;
    r[4] = t194;
    r[5] = t195;
    r[6] = t196;
    r[7] = t197;
    r[8] = t183;
    r[9] = t184;
    r[10] = t185;
    r[11] = t186;
    r[12] = t187;
    r[13] = t188;
#endif
  }

  // c = a*b for a small integer b (pseudo.py:705-728 / monty.py:876-978)
  static MAB_DEV void mli(uint32_t (&r)[14], const uint32_t (&a)[14], uint32_t b) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<66>;\n\t"
        "mul.lo.u32 t0, %14, %28;\n\t"
        "mul.hi.u32 t14, %14, %28;\n\t"
        "mul.lo.u32 t1, %15, %28;\n\t"
        "mul.hi.u32 t15, %15, %28;\n\t"
        "mul.lo.u32 t2, %16, %28;\n\t"
        "mul.hi.u32 t16, %16, %28;\n\t"
        "mul.lo.u32 t3, %17, %28;\n\t"
        "mul.hi.u32 t17, %17, %28;\n\t"
        "mul.lo.u32 t4, %18, %28;\n\t"
        "mul.hi.u32 t18, %18, %28;\n\t"
        "mul.lo.u32 t5, %19, %28;\n\t"
        "mul.hi.u32 t19, %19, %28;\n\t"
        "mul.lo.u32 t6, %20, %28;\n\t"
        "mul.hi.u32 t20, %20, %28;\n\t"
        "mul.lo.u32 t7, %21, %28;\n\t"
        "mul.hi.u32 t21, %21, %28;\n\t"
        "mul.lo.u32 t8, %22, %28;\n\t"
        "mul.hi.u32 t22, %22, %28;\n\t"
        "mul.lo.u32 t9, %23, %28;\n\t"
        "mul.hi.u32 t23, %23, %28;\n\t"
        "mul.lo.u32 t10, %24, %28;\n\t"
        "mul.hi.u32 t24, %24, %28;\n\t"
        "mul.lo.u32 t11, %25, %28;\n\t"
        "mul.hi.u32 t25, %25, %28;\n\t"
        "mul.lo.u32 t12, %26, %28;\n\t"
        "mul.hi.u32 t26, %26, %28;\n\t"
        "mul.lo.u32 t13, %27, %28;\n\t"
        "mul.hi.u32 t27, %27, %28;\n\t"
        "add.cc.u32 t28, t1, t14;\n\t"
        "addc.cc.u32 t29, t2, t15;\n\t"
        "addc.cc.u32 t30, t3, t16;\n\t"
        "addc.cc.u32 t31, t4, t17;\n\t"
        "addc.cc.u32 t32, t5, t18;\n\t"
        "addc.cc.u32 t33, t6, t19;\n\t"
        "addc.cc.u32 t34, t7, t20;\n\t"
        "addc.cc.u32 t35, t8, t21;\n\t"
        "addc.cc.u32 t36, t9, t22;\n\t"
        "addc.cc.u32 t37, t10, t23;\n\t"
        "addc.cc.u32 t38, t11, t24;\n\t"
        "addc.cc.u32 t39, t12, t25;\n\t"
        "addc.cc.u32 t40, t13, t26;\n\t"
        "addc.u32 t41, t27, 0x0;\n\t"
        "add.cc.u32 t42, t0, t41;\n\t"
        "addc.cc.u32 t43, t28, 0x0;\n\t"
        "addc.cc.u32 t44, t29, 0x0;\n\t"
        "addc.cc.u32 t45, t30, 0x0;\n\t"
        "addc.cc.u32 t46, t31, 0x0;\n\t"
        "addc.cc.u32 t47, t32, 0x0;\n\t"
        "addc.cc.u32 t48, t33, 0x0;\n\t"
        "addc.cc.u32 t49, t34, t41;\n\t"
        "addc.cc.u32 t50, t35, 0x0;\n\t"
        "addc.cc.u32 t51, t36, 0x0;\n\t"
        "addc.cc.u32 t52, t37, 0x0;\n\t"
        "addc.cc.u32 t53, t38, 0x0;\n\t"
        "addc.cc.u32 t54, t39, 0x0;\n\t"
        "addc.cc.u32 t55, t40, 0x0;\n\t"
        "addc.u32 t56, 0x0, 0x0;\n\t"
        "add.cc.u32 t57, t42, t56;\n\t"
        "addc.cc.u32 t58, t43, 0x0;\n\t"
        "addc.cc.u32 t59, t44, 0x0;\n\t"
        "addc.cc.u32 t60, t45, 0x0;\n\t"
        "addc.cc.u32 t61, t46, 0x0;\n\t"
        "addc.cc.u32 t62, t47, 0x0;\n\t"
        "addc.cc.u32 t63, t48, 0x0;\n\t"
        "addc.cc.u32 t64, t49, t56;\n\t"
        "addc.u32 t65, t50, 0x0;\n\t"
        "mov.u32 %0, t57;\n\t"
        "mov.u32 %1, t58;\n\t"
        "mov.u32 %2, t59;\n\t"
        "mov.u32 %3, t60;\n\t"
        "mov.u32 %4, t61;\n\t"
        "mov.u32 %5, t62;\n\t"
        "mov.u32 %6, t63;\n\t"
        "mov.u32 %7, t64;\n\t"
        "mov.u32 %8, t65;\n\t"
        "mov.u32 %9, t51;\n\t"
        "mov.u32 %10, t52;\n\t"
        "mov.u32 %11, t53;\n\t"
        "mov.u32 %12, t54;\n\t"
        "mov.u32 %13, t55;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(b));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t a_8_i = a[8];
    const uint32_t a_9_i = a[9];
    const uint32_t a_10_i = a[10];
    const uint32_t a_11_i = a[11];
    const uint32_t a_12_i = a[12];
    const uint32_t a_13_i = a[13];
    const uint32_t b_i = b;
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t0 = (uint32_t)((uint32_t)(a_0_i * b_i));
    t14 = (uint32_t)(((uint64_t)a_0_i * b_i) >> 32);
    t1 = (uint32_t)((uint32_t)(a_1_i * b_i));
    t15 = (uint32_t)(((uint64_t)a_1_i * b_i) >> 32);
    t2 = (uint32_t)((uint32_t)(a_2_i * b_i));
    t16 = (uint32_t)(((uint64_t)a_2_i * b_i) >> 32);
    t3 = (uint32_t)((uint32_t)(a_3_i * b_i));
    t17 = (uint32_t)(((uint64_t)a_3_i * b_i) >> 32);
    t4 = (uint32_t)((uint32_t)(a_4_i * b_i));
    t18 = (uint32_t)(((uint64_t)a_4_i * b_i) >> 32);
    t5 = (uint32_t)((uint32_t)(a_5_i * b_i));
    t19 = (uint32_t)(((uint64_t)a_5_i * b_i) >> 32);
    t6 = (uint32_t)((uint32_t)(a_6_i * b_i));
    t20 = (uint32_t)(((uint64_t)a_6_i * b_i) >> 32);
    t7 = (uint32_t)((uint32_t)(a_7_i * b_i));
    t21 = (uint32_t)(((uint64_t)a_7_i * b_i) >> 32);
    t8 = (uint32_t)((uint32_t)(a_8_i * b_i));
    t22 = (uint32_t)(((uint64_t)a_8_i * b_i) >> 32);
    t9 = (uint32_t)((uint32_t)(a_9_i * b_i));
    t23 = (uint32_t)(((uint64_t)a_9_i * b_i) >> 32);
    t10 = (uint32_t)((uint32_t)(a_10_i * b_i));
    t24 = (uint32_t)(((uint64_t)a_10_i * b_i) >> 32);
    t11 = (uint32_t)((uint32_t)(a_11_i * b_i));
    t25 = (uint32_t)(((uint64_t)a_11_i * b_i) >> 32);
    t12 = (uint32_t)((uint32_t)(a_12_i * b_i));
    t26 = (uint32_t)(((uint64_t)a_12_i * b_i) >> 32);
    t13 = (uint32_t)((uint32_t)(a_13_i * b_i));
    t27 = (uint32_t)(((uint64_t)a_13_i * b_i) >> 32);
    w_ = (uint64_t)t1 + t14; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t15 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t16 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t17 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t18 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t19 + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t20 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t8 + t21 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t22 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t23 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t24 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t25 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t26 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t27 + 0x0u + cf_; t41 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t41; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t28 + 0x0u + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t29 + 0x0u + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t30 + 0x0u + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t31 + 0x0u + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t33 + 0x0u + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t34 + t41 + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t35 + 0x0u + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t36 + 0x0u + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t37 + 0x0u + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t38 + 0x0u + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t39 + 0x0u + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + 0x0u + cf_; t55 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t56 = (uint32_t)w_;
    w_ = (uint64_t)t42 + t56; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t43 + 0x0u + cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t44 + 0x0u + cf_; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t45 + 0x0u + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t46 + 0x0u + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t47 + 0x0u + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t48 + 0x0u + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t49 + t56 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t50 + 0x0u + cf_; t65 = (uint32_t)w_;
    r[0] = t57;
    r[1] = t58;
    r[2] = t59;
    r[3] = t60;
    r[4] = t61;
    r[5] = t62;
    r[6] = t63;
    r[7] = t64;
    r[8] = t65;
    r[9] = t51;
    r[10] = t52;
    r[11] = t53;
    r[12] = t54;
    r[13] = t55;
#endif
  }

  // r = a*b + c, small integer b: modmli + modadd fused (rfc7748.c:209,212)
  static MAB_DEV void mla(uint32_t (&r)[14], const uint32_t (&a)[14], uint32_t b, const uint32_t (&c)[14]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<67>;\n\t"
        "mad.lo.cc.u32 t0, %14, %42, %28;\n\t"
        "madc.hi.cc.u32 t1, %14, %42, %29;\n\t"
        "madc.lo.cc.u32 t2, %16, %42, %30;\n\t"
        "madc.hi.cc.u32 t3, %16, %42, %31;\n\t"
        "madc.lo.cc.u32 t4, %18, %42, %32;\n\t"
        "madc.hi.cc.u32 t5, %18, %42, %33;\n\t"
        "madc.lo.cc.u32 t6, %20, %42, %34;\n\t"
        "madc.hi.cc.u32 t7, %20, %42, %35;\n\t"
        "madc.lo.cc.u32 t8, %22, %42, %36;\n\t"
        "madc.hi.cc.u32 t9, %22, %42, %37;\n\t"
        "madc.lo.cc.u32 t10, %24, %42, %38;\n\t"
        "madc.hi.cc.u32 t11, %24, %42, %39;\n\t"
        "madc.lo.cc.u32 t12, %26, %42, %40;\n\t"
        "madc.hi.cc.u32 t13, %26, %42, %41;\n\t"
        "addc.u32 t14, 0x0, 0x0;\n\t"
        "mul.lo.u32 t15, %15, %42;\n\t"
        "mul.hi.u32 t16, %15, %42;\n\t"
        "mul.lo.u32 t17, %17, %42;\n\t"
        "mul.hi.u32 t18, %17, %42;\n\t"
        "mul.lo.u32 t19, %19, %42;\n\t"
        "mul.hi.u32 t20, %19, %42;\n\t"
        "mul.lo.u32 t21, %21, %42;\n\t"
        "mul.hi.u32 t22, %21, %42;\n\t"
        "mul.lo.u32 t23, %23, %42;\n\t"
        "mul.hi.u32 t24, %23, %42;\n\t"
        "mul.lo.u32 t25, %25, %42;\n\t"
        "mul.hi.u32 t26, %25, %42;\n\t"
        "mul.lo.u32 t27, %27, %42;\n\t"
        "mul.hi.u32 t28, %27, %42;\n\t"
        "add.cc.u32 t29, t1, t15;\n\t"
        "addc.cc.u32 t30, t2, t16;\n\t"
        "addc.cc.u32 t31, t3, t17;\n\t"
        "addc.cc.u32 t32, t4, t18;\n\t"
        "addc.cc.u32 t33, t5, t19;\n\t"
        "addc.cc.u32 t34, t6, t20;\n\t"
        "addc.cc.u32 t35, t7, t21;\n\t"
        "addc.cc.u32 t36, t8, t22;\n\t"
        "addc.cc.u32 t37, t9, t23;\n\t"
        "addc.cc.u32 t38, t10, t24;\n\t"
        "addc.cc.u32 t39, t11, t25;\n\t"
        "addc.cc.u32 t40, t12, t26;\n\t"
        "addc.cc.u32 t41, t13, t27;\n\t"
        "addc.u32 t42, t14, t28;\n\t"
        "add.cc.u32 t43, t0, t42;\n\t"
        "addc.cc.u32 t44, t29, 0x0;\n\t"
        "addc.cc.u32 t45, t30, 0x0;\n\t"
        "addc.cc.u32 t46, t31, 0x0;\n\t"
        "addc.cc.u32 t47, t32, 0x0;\n\t"
        "addc.cc.u32 t48, t33, 0x0;\n\t"
        "addc.cc.u32 t49, t34, 0x0;\n\t"
        "addc.cc.u32 t50, t35, t42;\n\t"
        "addc.cc.u32 t51, t36, 0x0;\n\t"
        "addc.cc.u32 t52, t37, 0x0;\n\t"
        "addc.cc.u32 t53, t38, 0x0;\n\t"
        "addc.cc.u32 t54, t39, 0x0;\n\t"
        "addc.cc.u32 t55, t40, 0x0;\n\t"
        "addc.cc.u32 t56, t41, 0x0;\n\t"
        "addc.u32 t57, 0x0, 0x0;\n\t"
        "add.cc.u32 t58, t43, t57;\n\t"
        "addc.cc.u32 t59, t44, 0x0;\n\t"
        "addc.cc.u32 t60, t45, 0x0;\n\t"
        "addc.cc.u32 t61, t46, 0x0;\n\t"
        "addc.cc.u32 t62, t47, 0x0;\n\t"
        "addc.cc.u32 t63, t48, 0x0;\n\t"
        "addc.cc.u32 t64, t49, 0x0;\n\t"
        "addc.cc.u32 t65, t50, t57;\n\t"
        "addc.u32 t66, t51, 0x0;\n\t"
        "mov.u32 %0, t58;\n\t"
        "mov.u32 %1, t59;\n\t"
        "mov.u32 %2, t60;\n\t"
        "mov.u32 %3, t61;\n\t"
        "mov.u32 %4, t62;\n\t"
        "mov.u32 %5, t63;\n\t"
        "mov.u32 %6, t64;\n\t"
        "mov.u32 %7, t65;\n\t"
        "mov.u32 %8, t66;\n\t"
        "mov.u32 %9, t52;\n\t"
        "mov.u32 %10, t53;\n\t"
        "mov.u32 %11, t54;\n\t"
        "mov.u32 %12, t55;\n\t"
        "mov.u32 %13, t56;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]), "r"(c[8]), "r"(c[9]), "r"(c[10]), "r"(c[11]), "r"(c[12]), "r"(c[13]), "r"(b));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t a_8_i = a[8];
    const uint32_t a_9_i = a[9];
    const uint32_t a_10_i = a[10];
    const uint32_t a_11_i = a[11];
    const uint32_t a_12_i = a[12];
    const uint32_t a_13_i = a[13];
    const uint32_t c_0_i = c[0];
    const uint32_t c_1_i = c[1];
    const uint32_t c_2_i = c[2];
    const uint32_t c_3_i = c[3];
    const uint32_t c_4_i = c[4];
    const uint32_t c_5_i = c[5];
    const uint32_t c_6_i = c[6];
    const uint32_t c_7_i = c[7];
    const uint32_t c_8_i = c[8];
    const uint32_t c_9_i = c[9];
    const uint32_t c_10_i = c[10];
    const uint32_t c_11_i = c[11];
    const uint32_t c_12_i = c[12];
    const uint32_t c_13_i = c[13];
    const uint32_t b_i = b;
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * b_i) + c_0_i; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * b_i) >> 32) + c_1_i + cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * b_i) + c_2_i + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * b_i) >> 32) + c_3_i + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * b_i) + c_4_i + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * b_i) >> 32) + c_5_i + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * b_i) + c_6_i + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * b_i) >> 32) + c_7_i + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_8_i * b_i) + c_8_i + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_8_i * b_i) >> 32) + c_9_i + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_10_i * b_i) + c_10_i + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_10_i * b_i) >> 32) + c_11_i + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_12_i * b_i) + c_12_i + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_12_i * b_i) >> 32) + c_13_i + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t14 = (uint32_t)w_;
    t15 = (uint32_t)((uint32_t)(a_1_i * b_i));
    t16 = (uint32_t)(((uint64_t)a_1_i * b_i) >> 32);
    t17 = (uint32_t)((uint32_t)(a_3_i * b_i));
    t18 = (uint32_t)(((uint64_t)a_3_i * b_i) >> 32);
    t19 = (uint32_t)((uint32_t)(a_5_i * b_i));
    t20 = (uint32_t)(((uint64_t)a_5_i * b_i) >> 32);
    t21 = (uint32_t)((uint32_t)(a_7_i * b_i));
    t22 = (uint32_t)(((uint64_t)a_7_i * b_i) >> 32);
    t23 = (uint32_t)((uint32_t)(a_9_i * b_i));
    t24 = (uint32_t)(((uint64_t)a_9_i * b_i) >> 32);
    t25 = (uint32_t)((uint32_t)(a_11_i * b_i));
    t26 = (uint32_t)(((uint64_t)a_11_i * b_i) >> 32);
    t27 = (uint32_t)((uint32_t)(a_13_i * b_i));
    t28 = (uint32_t)(((uint64_t)a_13_i * b_i) >> 32);
    w_ = (uint64_t)t1 + t15; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t16 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t17 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t18 + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t19 + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t20 + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t21 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t8 + t22 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + t23 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + t24 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + t25 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + t26 + cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + t27 + cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t14 + t28 + cf_; t42 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t42; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t29 + 0x0u + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t30 + 0x0u + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t31 + 0x0u + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t33 + 0x0u + cf_; t48 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t34 + 0x0u + cf_; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t35 + t42 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t36 + 0x0u + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t37 + 0x0u + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t38 + 0x0u + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t39 + 0x0u + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + 0x0u + cf_; t55 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t41 + 0x0u + cf_; t56 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t57 = (uint32_t)w_;
    w_ = (uint64_t)t43 + t57; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t44 + 0x0u + cf_; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t45 + 0x0u + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t46 + 0x0u + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t47 + 0x0u + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t48 + 0x0u + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t49 + 0x0u + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t50 + t57 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t51 + 0x0u + cf_; t66 = (uint32_t)w_;
    r[0] = t58;
    r[1] = t59;
    r[2] = t60;
    r[3] = t61;
    r[4] = t62;
    r[5] = t63;
    r[6] = t64;
    r[7] = t65;
    r[8] = t66;
    r[9] = t52;
    r[10] = t53;
    r[11] = t54;
    r[12] = t55;
    r[13] = t56;
#endif
  }

  // n = a+b (pseudo.py:286-304)
  static MAB_DEV void add(uint32_t (&r)[14], const uint32_t (&a)[14], const uint32_t (&b)[14]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<38>;\n\t"
        "add.cc.u32 t0, %14, %28;\n\t"
        "addc.cc.u32 t1, %15, %29;\n\t"
        "addc.cc.u32 t2, %16, %30;\n\t"
        "addc.cc.u32 t3, %17, %31;\n\t"
        "addc.cc.u32 t4, %18, %32;\n\t"
        "addc.cc.u32 t5, %19, %33;\n\t"
        "addc.cc.u32 t6, %20, %34;\n\t"
        "addc.cc.u32 t7, %21, %35;\n\t"
        "addc.cc.u32 t8, %22, %36;\n\t"
        "addc.cc.u32 t9, %23, %37;\n\t"
        "addc.cc.u32 t10, %24, %38;\n\t"
        "addc.cc.u32 t11, %25, %39;\n\t"
        "addc.cc.u32 t12, %26, %40;\n\t"
        "addc.cc.u32 t13, %27, %41;\n\t"
        "addc.u32 t14, 0x0, 0x0;\n\t"
        "add.cc.u32 t15, t0, t14;\n\t"
        "addc.cc.u32 t16, t1, 0x0;\n\t"
        "addc.cc.u32 t17, t2, 0x0;\n\t"
        "addc.cc.u32 t18, t3, 0x0;\n\t"
        "addc.cc.u32 t19, t4, 0x0;\n\t"
        "addc.cc.u32 t20, t5, 0x0;\n\t"
        "addc.cc.u32 t21, t6, 0x0;\n\t"
        "addc.cc.u32 t22, t7, t14;\n\t"
        "addc.cc.u32 t23, t8, 0x0;\n\t"
        "addc.cc.u32 t24, t9, 0x0;\n\t"
        "addc.cc.u32 t25, t10, 0x0;\n\t"
        "addc.cc.u32 t26, t11, 0x0;\n\t"
        "addc.cc.u32 t27, t12, 0x0;\n\t"
        "addc.cc.u32 t28, t13, 0x0;\n\t"
        "addc.u32 t29, 0x0, 0x0;\n\t"
        "add.cc.u32 t30, t15, t29;\n\t"
        "addc.cc.u32 t31, t16, 0x0;\n\t"
        "addc.cc.u32 t32, t17, 0x0;\n\t"
        "addc.cc.u32 t33, t18, 0x0;\n\t"
        "addc.cc.u32 t34, t19, 0x0;\n\t"
        "addc.cc.u32 t35, t20, 0x0;\n\t"
        "addc.cc.u32 t36, t21, 0x0;\n\t"
        "addc.u32 t37, t22, t29;\n\t"
        "mov.u32 %0, t30;\n\t"
        "mov.u32 %1, t31;\n\t"
        "mov.u32 %2, t32;\n\t"
        "mov.u32 %3, t33;\n\t"
        "mov.u32 %4, t34;\n\t"
        "mov.u32 %5, t35;\n\t"
        "mov.u32 %6, t36;\n\t"
        "mov.u32 %7, t37;\n\t"
        "mov.u32 %8, t23;\n\t"
        "mov.u32 %9, t24;\n\t"
        "mov.u32 %10, t25;\n\t"
        "mov.u32 %11, t26;\n\t"
        "mov.u32 %12, t27;\n\t"
        "mov.u32 %13, t28;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(b[12]), "r"(b[13]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t a_8_i = a[8];
    const uint32_t a_9_i = a[9];
    const uint32_t a_10_i = a[10];
    const uint32_t a_11_i = a[11];
    const uint32_t a_12_i = a[12];
    const uint32_t a_13_i = a[13];
    const uint32_t b_0_i = b[0];
    const uint32_t b_1_i = b[1];
    const uint32_t b_2_i = b[2];
    const uint32_t b_3_i = b[3];
    const uint32_t b_4_i = b[4];
    const uint32_t b_5_i = b[5];
    const uint32_t b_6_i = b[6];
    const uint32_t b_7_i = b[7];
    const uint32_t b_8_i = b[8];
    const uint32_t b_9_i = b[9];
    const uint32_t b_10_i = b[10];
    const uint32_t b_11_i = b[11];
    const uint32_t b_12_i = b[12];
    const uint32_t b_13_i = b[13];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    w_ = (uint64_t)a_0_i + b_0_i; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_1_i + b_1_i + cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_2_i + b_2_i + cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_3_i + b_3_i + cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_4_i + b_4_i + cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_5_i + b_5_i + cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_6_i + b_6_i + cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_7_i + b_7_i + cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_8_i + b_8_i + cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_9_i + b_9_i + cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_10_i + b_10_i + cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_11_i + b_11_i + cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_12_i + b_12_i + cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)a_13_i + b_13_i + cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t14 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t14; t15 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t1 + 0x0u + cf_; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + 0x0u + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + 0x0u + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + 0x0u + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + 0x0u + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + 0x0u + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t14 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t8 + 0x0u + cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t9 + 0x0u + cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t10 + 0x0u + cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t11 + 0x0u + cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t12 + 0x0u + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t13 + 0x0u + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t29 = (uint32_t)w_;
    w_ = (uint64_t)t15 + t29; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t16 + 0x0u + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t17 + 0x0u + cf_; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t18 + 0x0u + cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t19 + 0x0u + cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t20 + 0x0u + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t21 + 0x0u + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t22 + t29 + cf_; t37 = (uint32_t)w_;
    r[0] = t30;
    r[1] = t31;
    r[2] = t32;
    r[3] = t33;
    r[4] = t34;
    r[5] = t35;
    r[6] = t36;
    r[7] = t37;
    r[8] = t23;
    r[9] = t24;
    r[10] = t25;
    r[11] = t26;
    r[12] = t27;
    r[13] = t28;
#endif
  }

  // n = a-b (pseudo.py:307-326)
  static MAB_DEV void sub(uint32_t (&r)[14], const uint32_t (&a)[14], const uint32_t (&b)[14]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<46>;\n\t"
        "sub.cc.u32 t0, %14, %28;\n\t"
        "subc.cc.u32 t1, %15, %29;\n\t"
        "subc.cc.u32 t2, %16, %30;\n\t"
        "subc.cc.u32 t3, %17, %31;\n\t"
        "subc.cc.u32 t4, %18, %32;\n\t"
        "subc.cc.u32 t5, %19, %33;\n\t"
        "subc.cc.u32 t6, %20, %34;\n\t"
        "subc.cc.u32 t7, %21, %35;\n\t"
        "subc.cc.u32 t8, %22, %36;\n\t"
        "subc.cc.u32 t9, %23, %37;\n\t"
        "subc.cc.u32 t10, %24, %38;\n\t"
        "subc.cc.u32 t11, %25, %39;\n\t"
        "subc.cc.u32 t12, %26, %40;\n\t"
        "subc.cc.u32 t13, %27, %41;\n\t"
        "subc.u32 t14, 0x0, 0x0;\n\t"
        "and.b32 t15, t14, 0x1;\n\t"
        "sub.cc.u32 t16, t0, t15;\n\t"
        "subc.cc.u32 t17, t1, 0x0;\n\t"
        "subc.cc.u32 t18, t2, 0x0;\n\t"
        "subc.cc.u32 t19, t3, 0x0;\n\t"
        "subc.cc.u32 t20, t4, 0x0;\n\t"
        "subc.cc.u32 t21, t5, 0x0;\n\t"
        "subc.cc.u32 t22, t6, 0x0;\n\t"
        "subc.cc.u32 t23, t7, t15;\n\t"
        "subc.cc.u32 t24, t8, 0x0;\n\t"
        "subc.cc.u32 t25, t9, 0x0;\n\t"
        "subc.cc.u32 t26, t10, 0x0;\n\t"
        "subc.cc.u32 t27, t11, 0x0;\n\t"
        "subc.cc.u32 t28, t12, 0x0;\n\t"
        "subc.cc.u32 t29, t13, 0x0;\n\t"
        "subc.u32 t30, 0x0, 0x0;\n\t"
        "and.b32 t31, t30, 0x1;\n\t"
        "sub.cc.u32 t32, t16, t31;\n\t"
        "subc.cc.u32 t33, t17, 0x0;\n\t"
        "subc.cc.u32 t34, t18, 0x0;\n\t"
        "subc.cc.u32 t35, t19, 0x0;\n\t"
        "subc.cc.u32 t36, t20, 0x0;\n\t"
        "subc.cc.u32 t37, t21, 0x0;\n\t"
        "subc.cc.u32 t38, t22, 0x0;\n\t"
        "subc.cc.u32 t39, t23, t31;\n\t"
        "subc.cc.u32 t40, t24, 0x0;\n\t"
        "subc.cc.u32 t41, t25, 0x0;\n\t"
        "subc.cc.u32 t42, t26, 0x0;\n\t"
        "subc.cc.u32 t43, t27, 0x0;\n\t"
        "subc.cc.u32 t44, t28, 0x0;\n\t"
        "subc.u32 t45, t29, 0x0;\n\t"
        "mov.u32 %0, t32;\n\t"
        "mov.u32 %1, t33;\n\t"
        "mov.u32 %2, t34;\n\t"
        "mov.u32 %3, t35;\n\t"
        "mov.u32 %4, t36;\n\t"
        "mov.u32 %5, t37;\n\t"
        "mov.u32 %6, t38;\n\t"
        "mov.u32 %7, t39;\n\t"
        "mov.u32 %8, t40;\n\t"
        "mov.u32 %9, t41;\n\t"
        "mov.u32 %10, t42;\n\t"
        "mov.u32 %11, t43;\n\t"
        "mov.u32 %12, t44;\n\t"
        "mov.u32 %13, t45;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(b[12]), "r"(b[13]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t a_8_i = a[8];
    const uint32_t a_9_i = a[9];
    const uint32_t a_10_i = a[10];
    const uint32_t a_11_i = a[11];
    const uint32_t a_12_i = a[12];
    const uint32_t a_13_i = a[13];
    const uint32_t b_0_i = b[0];
    const uint32_t b_1_i = b[1];
    const uint32_t b_2_i = b[2];
    const uint32_t b_3_i = b[3];
    const uint32_t b_4_i = b[4];
    const uint32_t b_5_i = b[5];
    const uint32_t b_6_i = b[6];
    const uint32_t b_7_i = b[7];
    const uint32_t b_8_i = b[8];
    const uint32_t b_9_i = b[9];
    const uint32_t b_10_i = b[10];
    const uint32_t b_11_i = b[11];
    const uint32_t b_12_i = b[12];
    const uint32_t b_13_i = b[13];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    w_ = (uint64_t)a_0_i - b_0_i; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_1_i - b_1_i - cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_2_i - b_2_i - cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_3_i - b_3_i - cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_4_i - b_4_i - cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_5_i - b_5_i - cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_6_i - b_6_i - cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_7_i - b_7_i - cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_8_i - b_8_i - cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_9_i - b_9_i - cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_10_i - b_10_i - cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_11_i - b_11_i - cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_12_i - b_12_i - cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_13_i - b_13_i - cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t14 = (uint32_t)w_;
    t15 = (uint32_t)(t14 & 0x1u);
    w_ = (uint64_t)t0 - t15; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t1 - 0x0u - cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t2 - 0x0u - cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t3 - 0x0u - cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t4 - 0x0u - cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t5 - 0x0u - cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t6 - 0x0u - cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t7 - t15 - cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t8 - 0x0u - cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t9 - 0x0u - cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t10 - 0x0u - cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t11 - 0x0u - cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t12 - 0x0u - cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t13 - 0x0u - cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t30 = (uint32_t)w_;
    t31 = (uint32_t)(t30 & 0x1u);
    w_ = (uint64_t)t16 - t31; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t17 - 0x0u - cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t18 - 0x0u - cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t19 - 0x0u - cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t20 - 0x0u - cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t21 - 0x0u - cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t22 - 0x0u - cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t23 - t31 - cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t24 - 0x0u - cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t25 - 0x0u - cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t26 - 0x0u - cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t27 - 0x0u - cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t28 - 0x0u - cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t29 - 0x0u - cf_; t45 = (uint32_t)w_;
    r[0] = t32;
    r[1] = t33;
    r[2] = t34;
    r[3] = t35;
    r[4] = t36;
    r[5] = t37;
    r[6] = t38;
    r[7] = t39;
    r[8] = t40;
    r[9] = t41;
    r[10] = t42;
    r[11] = t43;
    r[12] = t44;
    r[13] = t45;
#endif
  }

  // no spare bit above Nbits in this plan: the product-operand forms are the general ones
  static constexpr bool TIGHT = false;
  static MAB_DEV void add_tt(uint32_t (&r)[14], const uint32_t (&a)[14], const uint32_t (&b)[14]) { add(r, a, b); }
  static MAB_DEV void sub_tt(uint32_t (&r)[14], const uint32_t (&a)[14], const uint32_t (&b)[14]) { sub(r, a, b); }

  // no separate weakly-reduced products in this plan: chains use the ordinary ones
  static constexpr bool WEAK = false;
  static MAB_DEV void mul_w(uint32_t (&r)[14], const uint32_t (&a)[14], const uint32_t (&b)[14]) { mul(r, a, b); }
  static MAB_DEV void sqr_w(uint32_t (&r)[14], const uint32_t (&a)[14]) { sqr(r, a); }

  // n = -b (pseudo.py:329-348)
  static MAB_DEV void neg(uint32_t (&r)[14], const uint32_t (&b)[14]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<46>;\n\t"
        "sub.cc.u32 t0, 0x0, %14;\n\t"
        "subc.cc.u32 t1, 0x0, %15;\n\t"
        "subc.cc.u32 t2, 0x0, %16;\n\t"
        "subc.cc.u32 t3, 0x0, %17;\n\t"
        "subc.cc.u32 t4, 0x0, %18;\n\t"
        "subc.cc.u32 t5, 0x0, %19;\n\t"
        "subc.cc.u32 t6, 0x0, %20;\n\t"
        "subc.cc.u32 t7, 0x0, %21;\n\t"
        "subc.cc.u32 t8, 0x0, %22;\n\t"
        "subc.cc.u32 t9, 0x0, %23;\n\t"
        "subc.cc.u32 t10, 0x0, %24;\n\t"
        "subc.cc.u32 t11, 0x0, %25;\n\t"
        "subc.cc.u32 t12, 0x0, %26;\n\t"
        "subc.cc.u32 t13, 0x0, %27;\n\t"
        "subc.u32 t14, 0x0, 0x0;\n\t"
        "and.b32 t15, t14, 0x1;\n\t"
        "sub.cc.u32 t16, t0, t15;\n\t"
        "subc.cc.u32 t17, t1, 0x0;\n\t"
        "subc.cc.u32 t18, t2, 0x0;\n\t"
        "subc.cc.u32 t19, t3, 0x0;\n\t"
        "subc.cc.u32 t20, t4, 0x0;\n\t"
        "subc.cc.u32 t21, t5, 0x0;\n\t"
        "subc.cc.u32 t22, t6, 0x0;\n\t"
        "subc.cc.u32 t23, t7, t15;\n\t"
        "subc.cc.u32 t24, t8, 0x0;\n\t"
        "subc.cc.u32 t25, t9, 0x0;\n\t"
        "subc.cc.u32 t26, t10, 0x0;\n\t"
        "subc.cc.u32 t27, t11, 0x0;\n\t"
        "subc.cc.u32 t28, t12, 0x0;\n\t"
        "subc.cc.u32 t29, t13, 0x0;\n\t"
        "subc.u32 t30, 0x0, 0x0;\n\t"
        "and.b32 t31, t30, 0x1;\n\t"
        "sub.cc.u32 t32, t16, t31;\n\t"
        "subc.cc.u32 t33, t17, 0x0;\n\t"
        "subc.cc.u32 t34, t18, 0x0;\n\t"
        "subc.cc.u32 t35, t19, 0x0;\n\t"
        "subc.cc.u32 t36, t20, 0x0;\n\t"
        "subc.cc.u32 t37, t21, 0x0;\n\t"
        "subc.cc.u32 t38, t22, 0x0;\n\t"
        "subc.cc.u32 t39, t23, t31;\n\t"
        "subc.cc.u32 t40, t24, 0x0;\n\t"
        "subc.cc.u32 t41, t25, 0x0;\n\t"
        "subc.cc.u32 t42, t26, 0x0;\n\t"
        "subc.cc.u32 t43, t27, 0x0;\n\t"
        "subc.cc.u32 t44, t28, 0x0;\n\t"
        "subc.u32 t45, t29, 0x0;\n\t"
        "mov.u32 %0, t32;\n\t"
        "mov.u32 %1, t33;\n\t"
        "mov.u32 %2, t34;\n\t"
        "mov.u32 %3, t35;\n\t"
        "mov.u32 %4, t36;\n\t"
        "mov.u32 %5, t37;\n\t"
        "mov.u32 %6, t38;\n\t"
        "mov.u32 %7, t39;\n\t"
        "mov.u32 %8, t40;\n\t"
        "mov.u32 %9, t41;\n\t"
        "mov.u32 %10, t42;\n\t"
        "mov.u32 %11, t43;\n\t"
        "mov.u32 %12, t44;\n\t"
        "mov.u32 %13, t45;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(b[12]), "r"(b[13]));
#else
    const uint32_t b_0_i = b[0];
    const uint32_t b_1_i = b[1];
    const uint32_t b_2_i = b[2];
    const uint32_t b_3_i = b[3];
    const uint32_t b_4_i = b[4];
    const uint32_t b_5_i = b[5];
    const uint32_t b_6_i = b[6];
    const uint32_t b_7_i = b[7];
    const uint32_t b_8_i = b[8];
    const uint32_t b_9_i = b[9];
    const uint32_t b_10_i = b[10];
    const uint32_t b_11_i = b[11];
    const uint32_t b_12_i = b[12];
    const uint32_t b_13_i = b[13];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    w_ = (uint64_t)0x0u - b_0_i; t0 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_1_i - cf_; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_2_i - cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_3_i - cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_4_i - cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_5_i - cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_6_i - cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_7_i - cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_8_i - cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_9_i - cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_10_i - cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_11_i - cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_12_i - cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - b_13_i - cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t14 = (uint32_t)w_;
    t15 = (uint32_t)(t14 & 0x1u);
    w_ = (uint64_t)t0 - t15; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t1 - 0x0u - cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t2 - 0x0u - cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t3 - 0x0u - cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t4 - 0x0u - cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t5 - 0x0u - cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t6 - 0x0u - cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t7 - t15 - cf_; t23 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t8 - 0x0u - cf_; t24 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t9 - 0x0u - cf_; t25 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t10 - 0x0u - cf_; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t11 - 0x0u - cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t12 - 0x0u - cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t13 - 0x0u - cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t30 = (uint32_t)w_;
    t31 = (uint32_t)(t30 & 0x1u);
    w_ = (uint64_t)t16 - t31; t32 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t17 - 0x0u - cf_; t33 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t18 - 0x0u - cf_; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t19 - 0x0u - cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t20 - 0x0u - cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t21 - 0x0u - cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t22 - 0x0u - cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t23 - t31 - cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t24 - 0x0u - cf_; t40 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t25 - 0x0u - cf_; t41 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t26 - 0x0u - cf_; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t27 - 0x0u - cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t28 - 0x0u - cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t29 - 0x0u - cf_; t45 = (uint32_t)w_;
    r[0] = t32;
    r[1] = t33;
    r[2] = t34;
    r[3] = t35;
    r[4] = t36;
    r[5] = t37;
    r[6] = t38;
    r[7] = t39;
    r[8] = t40;
    r[9] = t41;
    r[10] = t42;
    r[11] = t43;
    r[12] = t44;
    r[13] = t45;
#endif
  }

  // canonical residue of a stored value; returns 1 iff it was already < p
  // (flatten/modfsb, pseudo.py:255-283)
  static MAB_DEV uint32_t canon(uint32_t (&r)[14], const uint32_t (&a)[14]) {
    uint32_t lt;
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<58>;\n\t"
        "sub.cc.u32 t1, %15, 0xffffffff;\n\t"
        "subc.cc.u32 t2, %16, 0xffffffff;\n\t"
        "subc.cc.u32 t3, %17, 0xffffffff;\n\t"
        "subc.cc.u32 t4, %18, 0xffffffff;\n\t"
        "subc.cc.u32 t5, %19, 0xffffffff;\n\t"
        "subc.cc.u32 t6, %20, 0xffffffff;\n\t"
        "subc.cc.u32 t7, %21, 0xffffffff;\n\t"
        "subc.cc.u32 t8, %22, 0xfffffffe;\n\t"
        "subc.cc.u32 t9, %23, 0xffffffff;\n\t"
        "subc.cc.u32 t10, %24, 0xffffffff;\n\t"
        "subc.cc.u32 t11, %25, 0xffffffff;\n\t"
        "subc.cc.u32 t12, %26, 0xffffffff;\n\t"
        "subc.cc.u32 t13, %27, 0xffffffff;\n\t"
        "subc.cc.u32 t14, %28, 0xffffffff;\n\t"
        "subc.u32 t15, 0x0, 0x0;\n\t"
        "xor.b32 t16, t1, %15;\n\t"
        "and.b32 t17, t16, t15;\n\t"
        "xor.b32 t18, t17, t1;\n\t"
        "xor.b32 t19, t2, %16;\n\t"
        "and.b32 t20, t19, t15;\n\t"
        "xor.b32 t21, t20, t2;\n\t"
        "xor.b32 t22, t3, %17;\n\t"
        "and.b32 t23, t22, t15;\n\t"
        "xor.b32 t24, t23, t3;\n\t"
        "xor.b32 t25, t4, %18;\n\t"
        "and.b32 t26, t25, t15;\n\t"
        "xor.b32 t27, t26, t4;\n\t"
        "xor.b32 t28, t5, %19;\n\t"
        "and.b32 t29, t28, t15;\n\t"
        "xor.b32 t30, t29, t5;\n\t"
        "xor.b32 t31, t6, %20;\n\t"
        "and.b32 t32, t31, t15;\n\t"
        "xor.b32 t33, t32, t6;\n\t"
        "xor.b32 t34, t7, %21;\n\t"
        "and.b32 t35, t34, t15;\n\t"
        "xor.b32 t36, t35, t7;\n\t"
        "xor.b32 t37, t8, %22;\n\t"
        "and.b32 t38, t37, t15;\n\t"
        "xor.b32 t39, t38, t8;\n\t"
        "xor.b32 t40, t9, %23;\n\t"
        "and.b32 t41, t40, t15;\n\t"
        "xor.b32 t42, t41, t9;\n\t"
        "xor.b32 t43, t10, %24;\n\t"
        "and.b32 t44, t43, t15;\n\t"
        "xor.b32 t45, t44, t10;\n\t"
        "xor.b32 t46, t11, %25;\n\t"
        "and.b32 t47, t46, t15;\n\t"
        "xor.b32 t48, t47, t11;\n\t"
        "xor.b32 t49, t12, %26;\n\t"
        "and.b32 t50, t49, t15;\n\t"
        "xor.b32 t51, t50, t12;\n\t"
        "xor.b32 t52, t13, %27;\n\t"
        "and.b32 t53, t52, t15;\n\t"
        "xor.b32 t54, t53, t13;\n\t"
        "xor.b32 t55, t14, %28;\n\t"
        "and.b32 t56, t55, t15;\n\t"
        "xor.b32 t57, t56, t14;\n\t"
        "and.b32 t0, t15, 0x1;\n\t"
        "mov.u32 %0, t18;\n\t"
        "mov.u32 %1, t21;\n\t"
        "mov.u32 %2, t24;\n\t"
        "mov.u32 %3, t27;\n\t"
        "mov.u32 %4, t30;\n\t"
        "mov.u32 %5, t33;\n\t"
        "mov.u32 %6, t36;\n\t"
        "mov.u32 %7, t39;\n\t"
        "mov.u32 %8, t42;\n\t"
        "mov.u32 %9, t45;\n\t"
        "mov.u32 %10, t48;\n\t"
        "mov.u32 %11, t51;\n\t"
        "mov.u32 %12, t54;\n\t"
        "mov.u32 %13, t57;\n\t"
        "mov.u32 %14, t0;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(lt)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]));
#else
    const uint32_t a_0_i = a[0];
    const uint32_t a_1_i = a[1];
    const uint32_t a_2_i = a[2];
    const uint32_t a_3_i = a[3];
    const uint32_t a_4_i = a[4];
    const uint32_t a_5_i = a[5];
    const uint32_t a_6_i = a[6];
    const uint32_t a_7_i = a[7];
    const uint32_t a_8_i = a[8];
    const uint32_t a_9_i = a[9];
    const uint32_t a_10_i = a[10];
    const uint32_t a_11_i = a[11];
    const uint32_t a_12_i = a[12];
    const uint32_t a_13_i = a[13];
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    w_ = (uint64_t)a_0_i - 0xffffffffu; t1 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_1_i - 0xffffffffu - cf_; t2 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_2_i - 0xffffffffu - cf_; t3 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_3_i - 0xffffffffu - cf_; t4 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_4_i - 0xffffffffu - cf_; t5 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_5_i - 0xffffffffu - cf_; t6 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_6_i - 0xffffffffu - cf_; t7 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_7_i - 0xfffffffeu - cf_; t8 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_8_i - 0xffffffffu - cf_; t9 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_9_i - 0xffffffffu - cf_; t10 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_10_i - 0xffffffffu - cf_; t11 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_11_i - 0xffffffffu - cf_; t12 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_12_i - 0xffffffffu - cf_; t13 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)a_13_i - 0xffffffffu - cf_; t14 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t15 = (uint32_t)w_;
    t16 = (uint32_t)(t1 ^ a_0_i);
    t17 = (uint32_t)(t16 & t15);
    t18 = (uint32_t)(t17 ^ t1);
    t19 = (uint32_t)(t2 ^ a_1_i);
    t20 = (uint32_t)(t19 & t15);
    t21 = (uint32_t)(t20 ^ t2);
    t22 = (uint32_t)(t3 ^ a_2_i);
    t23 = (uint32_t)(t22 & t15);
    t24 = (uint32_t)(t23 ^ t3);
    t25 = (uint32_t)(t4 ^ a_3_i);
    t26 = (uint32_t)(t25 & t15);
    t27 = (uint32_t)(t26 ^ t4);
    t28 = (uint32_t)(t5 ^ a_4_i);
    t29 = (uint32_t)(t28 & t15);
    t30 = (uint32_t)(t29 ^ t5);
    t31 = (uint32_t)(t6 ^ a_5_i);
    t32 = (uint32_t)(t31 & t15);
    t33 = (uint32_t)(t32 ^ t6);
    t34 = (uint32_t)(t7 ^ a_6_i);
    t35 = (uint32_t)(t34 & t15);
    t36 = (uint32_t)(t35 ^ t7);
    t37 = (uint32_t)(t8 ^ a_7_i);
    t38 = (uint32_t)(t37 & t15);
    t39 = (uint32_t)(t38 ^ t8);
    t40 = (uint32_t)(t9 ^ a_8_i);
    t41 = (uint32_t)(t40 & t15);
    t42 = (uint32_t)(t41 ^ t9);
    t43 = (uint32_t)(t10 ^ a_9_i);
    t44 = (uint32_t)(t43 & t15);
    t45 = (uint32_t)(t44 ^ t10);
    t46 = (uint32_t)(t11 ^ a_10_i);
    t47 = (uint32_t)(t46 & t15);
    t48 = (uint32_t)(t47 ^ t11);
    t49 = (uint32_t)(t12 ^ a_11_i);
    t50 = (uint32_t)(t49 & t15);
    t51 = (uint32_t)(t50 ^ t12);
    t52 = (uint32_t)(t13 ^ a_12_i);
    t53 = (uint32_t)(t52 & t15);
    t54 = (uint32_t)(t53 ^ t13);
    t55 = (uint32_t)(t14 ^ a_13_i);
    t56 = (uint32_t)(t55 & t15);
    t57 = (uint32_t)(t56 ^ t14);
    t0 = (uint32_t)(t15 & 0x1u);
    r[0] = t18;
    r[1] = t21;
    r[2] = t24;
    r[3] = t27;
    r[4] = t30;
    r[5] = t33;
    r[6] = t36;
    r[7] = t39;
    r[8] = t42;
    r[9] = t45;
    r[10] = t48;
    r[11] = t51;
    r[12] = t54;
    r[13] = t57;
    lt = t0;
#endif
    return lt;
  }

  static MAB_DEV void set_p(uint32_t (&r)[14]) { r[0] = 0xffffffffu; r[1] = 0xffffffffu; r[2] = 0xffffffffu; r[3] = 0xffffffffu; r[4] = 0xffffffffu; r[5] = 0xffffffffu; r[6] = 0xffffffffu; r[7] = 0xfffffffeu; r[8] = 0xffffffffu; r[9] = 0xffffffffu; r[10] = 0xffffffffu; r[11] = 0xffffffffu; r[12] = 0xffffffffu; r[13] = 0xffffffffu; }
  static MAB_DEV void set_one(uint32_t (&r)[14]) { r[0] = 0x00000001u; r[1] = 0x00000000u; r[2] = 0x00000000u; r[3] = 0x00000000u; r[4] = 0x00000000u; r[5] = 0x00000000u; r[6] = 0x00000000u; r[7] = 0x00000000u; r[8] = 0x00000000u; r[9] = 0x00000000u; r[10] = 0x00000000u; r[11] = 0x00000000u; r[12] = 0x00000000u; r[13] = 0x00000000u; }
  static MAB_DEV void set_roi(uint32_t (&r)[14]) { r[0] = 0xfffffffeu; r[1] = 0xffffffffu; r[2] = 0xffffffffu; r[3] = 0xffffffffu; r[4] = 0xffffffffu; r[5] = 0xffffffffu; r[6] = 0xffffffffu; r[7] = 0xfffffffeu; r[8] = 0xffffffffu; r[9] = 0xffffffffu; r[10] = 0xffffffffu; r[11] = 0xffffffffu; r[12] = 0xffffffffu; r[13] = 0xffffffffu; }
  static MAB_DEV void set_r2(uint32_t (&r)[14]) { r[0] = 0x00000001u; r[1] = 0x00000000u; r[2] = 0x00000000u; r[3] = 0x00000000u; r[4] = 0x00000000u; r[5] = 0x00000000u; r[6] = 0x00000000u; r[7] = 0x00000000u; r[8] = 0x00000000u; r[9] = 0x00000000u; r[10] = 0x00000000u; r[11] = 0x00000000u; r[12] = 0x00000000u; r[13] = 0x00000000u; }
  static constexpr bool HAS_WEIERSTRASS = false;

  // nres: copy (pseudo.py:952-962); redc: copy + final subtract (pseudo.py:965-976)
  static MAB_DEV void nres(uint32_t (&r)[14], const uint32_t (&a)[14]) { for (int i = 0; i < L; i++) r[i] = a[i]; }
  static MAB_DEV void redc(uint32_t (&r)[14], const uint32_t (&a)[14]) { (void)canon(r, a); }

  // z = w^PE, straight-line addition chain (pseudo.py:758-785; our own chain finder)
  static MAB_DEV void pro(uint32_t (&z)[14], const uint32_t (&w)[14]) {
    uint32_t x[L];
    for (int i = 0; i < L; i++) x[i] = w[i];
    uint32_t t0[L];
    uint32_t t1[L];
    sqr_w(t0, x);
    mul_w(t0, t0, x);
    sqr_w(t0, t0);
    mul_w(t0, t0, x);
    sqr_w(t1, t0);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr_w(t1, t1);
    mul_w(t1, t1, t0);
    sqr_w(t0, t1);
    MAB_NOUNROLL
    for (int i = 1; i < 6; i++) sqr_w(t0, t0);
    mul_w(t0, t0, t1);
    sqr_w(t0, t0);
    mul_w(t0, t0, x);
    sqr_w(t1, t0);
    MAB_NOUNROLL
    for (int i = 1; i < 13; i++) sqr_w(t1, t1);
    mul_w(t1, t1, t0);
    sqr_w(t1, t1);
    mul_w(t1, t1, x);
    sqr_w(t0, t1);
    MAB_NOUNROLL
    for (int i = 1; i < 27; i++) sqr_w(t0, t0);
    mul_w(t0, t0, t1);
    sqr_w(t0, t0);
    mul_w(t0, t0, x);
    sqr_w(t1, t0);
    MAB_NOUNROLL
    for (int i = 1; i < 55; i++) sqr_w(t1, t1);
    mul_w(t1, t1, t0);
    sqr_w(t1, t1);
    mul_w(t1, t1, x);
    sqr_w(t0, t1);
    MAB_NOUNROLL
    for (int i = 1; i < 111; i++) sqr_w(t0, t0);
    mul_w(t0, t0, t1);
    sqr_w(t1, t0);
    mul_w(t1, t1, x);
    sqr_w(z, t1);
    MAB_NOUNROLL
    for (int i = 1; i < 223; i++) sqr_w(z, z);
    mul_w(z, z, t0);
    if (WEAK) (void)canon(z, z);
  }
};
