// Automatically generated field arithmetic for sm_100a -- do not edit.
// Command line : python -m modarith_b200.gen.monty_sm100 NIST256ORDER
// modulus NIST256ORDER = 0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551
// plan MontgomeryFull: 8 saturated 32-bit limbs; stored values < p; R = 2^256
//   mul   : 156 IMAD.WIDE   8 IMAD  ~113 ALU-pipe ops
//   sqr   : 128 IMAD.WIDE   8 IMAD  ~125 ALU-pipe ops
//   mli   : 256 IMAD.WIDE  16 IMAD  ~212 ALU-pipe ops
//   mla   : 256 IMAD.WIDE  16 IMAD  ~255 ALU-pipe ops
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
  static constexpr int LADDER_MINBLOCKS = 4;   // resident 128-thread CTAs per SM for k_rfc7748
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
        ".reg .u32 t<178>;\n\t"
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
        "mul.lo.u32 t47, t0, 0xee00bc4f;\n\t"
        "mul.hi.u32 t48, t0, 0xee00bc4f;\n\t"
        "mul.lo.u32 t49, t33, 0xee00bc4f;\n\t"
        "mul.hi.u32 t50, t33, 0xee00bc4f;\n\t"
        "mul.lo.u32 t51, t35, 0xee00bc4f;\n\t"
        "mul.hi.u32 t52, t35, 0xee00bc4f;\n\t"
        "mul.lo.u32 t53, t37, 0xee00bc4f;\n\t"
        "mul.hi.u32 t54, t37, 0xee00bc4f;\n\t"
        "mul.lo.u32 t57, t32, 0xee00bc4f;\n\t"
        "mul.hi.u32 t58, t32, 0xee00bc4f;\n\t"
        "mul.lo.u32 t59, t34, 0xee00bc4f;\n\t"
        "mul.hi.u32 t60, t34, 0xee00bc4f;\n\t"
        "mul.lo.u32 t61, t36, 0xee00bc4f;\n\t"
        "mul.hi.u32 t62, t36, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t49, t32, 0xccd1c8aa, t49;\n\t"
        "madc.hi.cc.u32 t50, t32, 0xccd1c8aa, t50;\n\t"
        "madc.lo.cc.u32 t51, t34, 0xccd1c8aa, t51;\n\t"
        "madc.hi.cc.u32 t52, t34, 0xccd1c8aa, t52;\n\t"
        "madc.lo.cc.u32 t53, t36, 0xccd1c8aa, t53;\n\t"
        "madc.hi.cc.u32 t54, t36, 0xccd1c8aa, t54;\n\t"
        "addc.u32 t55, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t57, t0, 0xccd1c8aa, t57;\n\t"
        "madc.hi.cc.u32 t58, t0, 0xccd1c8aa, t58;\n\t"
        "madc.lo.cc.u32 t59, t33, 0xccd1c8aa, t59;\n\t"
        "madc.hi.cc.u32 t60, t33, 0xccd1c8aa, t60;\n\t"
        "madc.lo.cc.u32 t61, t35, 0xccd1c8aa, t61;\n\t"
        "madc.hi.cc.u32 t62, t35, 0xccd1c8aa, t62;\n\t"
        "addc.u32 t63, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t49, t0, 0x7d74d2e4, t49;\n\t"
        "madc.hi.cc.u32 t50, t0, 0x7d74d2e4, t50;\n\t"
        "madc.lo.cc.u32 t51, t33, 0x7d74d2e4, t51;\n\t"
        "madc.hi.cc.u32 t52, t33, 0x7d74d2e4, t52;\n\t"
        "madc.lo.cc.u32 t53, t35, 0x7d74d2e4, t53;\n\t"
        "madc.hi.cc.u32 t54, t35, 0x7d74d2e4, t54;\n\t"
        "addc.u32 t55, t55, 0x0;\n\t"
        "mad.lo.cc.u32 t59, t32, 0x7d74d2e4, t59;\n\t"
        "madc.hi.cc.u32 t60, t32, 0x7d74d2e4, t60;\n\t"
        "madc.lo.cc.u32 t61, t34, 0x7d74d2e4, t61;\n\t"
        "madc.hi.cc.u32 t62, t34, 0x7d74d2e4, t62;\n\t"
        "addc.u32 t63, t63, 0x0;\n\t"
        "mad.lo.cc.u32 t51, t32, 0x48c94408, t51;\n\t"
        "madc.hi.cc.u32 t52, t32, 0x48c94408, t52;\n\t"
        "madc.lo.cc.u32 t53, t34, 0x48c94408, t53;\n\t"
        "madc.hi.cc.u32 t54, t34, 0x48c94408, t54;\n\t"
        "addc.u32 t55, t55, 0x0;\n\t"
        "mad.lo.cc.u32 t59, t0, 0x48c94408, t59;\n\t"
        "madc.hi.cc.u32 t60, t0, 0x48c94408, t60;\n\t"
        "madc.lo.cc.u32 t61, t33, 0x48c94408, t61;\n\t"
        "madc.hi.cc.u32 t62, t33, 0x48c94408, t62;\n\t"
        "addc.u32 t63, t63, 0x0;\n\t"
        "mad.lo.cc.u32 t51, t0, 0xc588c6f6, t51;\n\t"
        "madc.hi.cc.u32 t52, t0, 0xc588c6f6, t52;\n\t"
        "madc.lo.cc.u32 t53, t33, 0xc588c6f6, t53;\n\t"
        "madc.hi.cc.u32 t54, t33, 0xc588c6f6, t54;\n\t"
        "addc.u32 t55, t55, 0x0;\n\t"
        "mad.lo.cc.u32 t61, t32, 0xc588c6f6, t61;\n\t"
        "madc.hi.cc.u32 t62, t32, 0xc588c6f6, t62;\n\t"
        "addc.u32 t63, t63, 0x0;\n\t"
        "mad.lo.cc.u32 t53, t32, 0x50fe77ec, t53;\n\t"
        "madc.hi.cc.u32 t54, t32, 0x50fe77ec, t54;\n\t"
        "addc.u32 t55, t55, 0x0;\n\t"
        "mad.lo.cc.u32 t61, t0, 0x50fe77ec, t61;\n\t"
        "madc.hi.cc.u32 t62, t0, 0x50fe77ec, t62;\n\t"
        "addc.u32 t63, t63, 0x0;\n\t"
        "mad.lo.cc.u32 t53, t0, 0xa9d6281c, t53;\n\t"
        "madc.hi.cc.u32 t54, t0, 0xa9d6281c, t54;\n\t"
        "addc.u32 t55, t55, 0x0;\n\t"
        "add.cc.u32 t65, t48, t57;\n\t"
        "addc.cc.u32 t66, t49, t58;\n\t"
        "addc.cc.u32 t67, t50, t59;\n\t"
        "addc.cc.u32 t68, t51, t60;\n\t"
        "addc.cc.u32 t69, t52, t61;\n\t"
        "addc.cc.u32 t70, t53, t62;\n\t"
        "addc.u32 t71, t54, t63;\n\t"
        "mad.lo.u32 t72, t38, 0xee00bc4f, t71;\n\t"
        "mad.lo.u32 t73, t37, 0xccd1c8aa, t72;\n\t"
        "mad.lo.u32 t74, t36, 0x7d74d2e4, t73;\n\t"
        "mad.lo.u32 t75, t35, 0x48c94408, t74;\n\t"
        "mad.lo.u32 t76, t34, 0xc588c6f6, t75;\n\t"
        "mad.lo.u32 t77, t33, 0x50fe77ec, t76;\n\t"
        "mad.lo.u32 t78, t32, 0xa9d6281c, t77;\n\t"
        "mad.lo.u32 t79, t0, 0x60d06633, t78;\n\t"
        "mul.lo.u32 t80, t47, 0xfc632551;\n\t"
        "mul.hi.u32 t81, t47, 0xfc632551;\n\t"
        "mul.lo.u32 t82, t66, 0xfc632551;\n\t"
        "mul.hi.u32 t83, t66, 0xfc632551;\n\t"
        "mul.lo.u32 t84, t68, 0xfc632551;\n\t"
        "mul.hi.u32 t85, t68, 0xfc632551;\n\t"
        "mul.lo.u32 t86, t70, 0xfc632551;\n\t"
        "mul.hi.u32 t87, t70, 0xfc632551;\n\t"
        "mul.lo.u32 t97, t65, 0xfc632551;\n\t"
        "mul.hi.u32 t98, t65, 0xfc632551;\n\t"
        "mul.lo.u32 t99, t67, 0xfc632551;\n\t"
        "mul.hi.u32 t100, t67, 0xfc632551;\n\t"
        "mul.lo.u32 t101, t69, 0xfc632551;\n\t"
        "mul.hi.u32 t102, t69, 0xfc632551;\n\t"
        "mul.lo.u32 t103, t79, 0xfc632551;\n\t"
        "mul.hi.u32 t104, t79, 0xfc632551;\n\t"
        "mad.lo.cc.u32 t82, t65, 0xf3b9cac2, t82;\n\t"
        "madc.hi.cc.u32 t83, t65, 0xf3b9cac2, t83;\n\t"
        "madc.lo.cc.u32 t84, t67, 0xf3b9cac2, t84;\n\t"
        "madc.hi.cc.u32 t85, t67, 0xf3b9cac2, t85;\n\t"
        "madc.lo.cc.u32 t86, t69, 0xf3b9cac2, t86;\n\t"
        "madc.hi.cc.u32 t87, t69, 0xf3b9cac2, t87;\n\t"
        "madc.lo.cc.u32 t88, t79, 0xf3b9cac2, 0x0;\n\t"
        "madc.hi.u32 t89, t79, 0xf3b9cac2, 0x0;\n\t"
        "mad.lo.cc.u32 t97, t47, 0xf3b9cac2, t97;\n\t"
        "madc.hi.cc.u32 t98, t47, 0xf3b9cac2, t98;\n\t"
        "madc.lo.cc.u32 t99, t66, 0xf3b9cac2, t99;\n\t"
        "madc.hi.cc.u32 t100, t66, 0xf3b9cac2, t100;\n\t"
        "madc.lo.cc.u32 t101, t68, 0xf3b9cac2, t101;\n\t"
        "madc.hi.cc.u32 t102, t68, 0xf3b9cac2, t102;\n\t"
        "madc.lo.cc.u32 t103, t70, 0xf3b9cac2, t103;\n\t"
        "madc.hi.cc.u32 t104, t70, 0xf3b9cac2, t104;\n\t"
        "addc.u32 t105, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t82, t47, 0xa7179e84, t82;\n\t"
        "madc.hi.cc.u32 t83, t47, 0xa7179e84, t83;\n\t"
        "madc.lo.cc.u32 t84, t66, 0xa7179e84, t84;\n\t"
        "madc.hi.cc.u32 t85, t66, 0xa7179e84, t85;\n\t"
        "madc.lo.cc.u32 t86, t68, 0xa7179e84, t86;\n\t"
        "madc.hi.cc.u32 t87, t68, 0xa7179e84, t87;\n\t"
        "madc.lo.cc.u32 t88, t70, 0xa7179e84, t88;\n\t"
        "madc.hi.cc.u32 t89, t70, 0xa7179e84, t89;\n\t"
        "addc.u32 t90, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t99, t65, 0xa7179e84, t99;\n\t"
        "madc.hi.cc.u32 t100, t65, 0xa7179e84, t100;\n\t"
        "madc.lo.cc.u32 t101, t67, 0xa7179e84, t101;\n\t"
        "madc.hi.cc.u32 t102, t67, 0xa7179e84, t102;\n\t"
        "madc.lo.cc.u32 t103, t69, 0xa7179e84, t103;\n\t"
        "madc.hi.cc.u32 t104, t69, 0xa7179e84, t104;\n\t"
        "madc.lo.cc.u32 t105, t79, 0xa7179e84, t105;\n\t"
        "madc.hi.u32 t106, t79, 0xa7179e84, 0x0;\n\t"
        "mad.lo.cc.u32 t84, t65, 0xbce6faad, t84;\n\t"
        "madc.hi.cc.u32 t85, t65, 0xbce6faad, t85;\n\t"
        "madc.lo.cc.u32 t86, t67, 0xbce6faad, t86;\n\t"
        "madc.hi.cc.u32 t87, t67, 0xbce6faad, t87;\n\t"
        "madc.lo.cc.u32 t88, t69, 0xbce6faad, t88;\n\t"
        "madc.hi.cc.u32 t89, t69, 0xbce6faad, t89;\n\t"
        "madc.lo.cc.u32 t90, t79, 0xbce6faad, t90;\n\t"
        "madc.hi.u32 t91, t79, 0xbce6faad, 0x0;\n\t"
        "mad.lo.cc.u32 t99, t47, 0xbce6faad, t99;\n\t"
        "madc.hi.cc.u32 t100, t47, 0xbce6faad, t100;\n\t"
        "madc.lo.cc.u32 t101, t66, 0xbce6faad, t101;\n\t"
        "madc.hi.cc.u32 t102, t66, 0xbce6faad, t102;\n\t"
        "madc.lo.cc.u32 t103, t68, 0xbce6faad, t103;\n\t"
        "madc.hi.cc.u32 t104, t68, 0xbce6faad, t104;\n\t"
        "madc.lo.cc.u32 t105, t70, 0xbce6faad, t105;\n\t"
        "madc.hi.cc.u32 t106, t70, 0xbce6faad, t106;\n\t"
        "addc.u32 t107, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t84, t47, 0xffffffff, t84;\n\t"
        "madc.hi.cc.u32 t85, t47, 0xffffffff, t85;\n\t"
        "madc.lo.cc.u32 t86, t66, 0xffffffff, t86;\n\t"
        "madc.hi.cc.u32 t87, t66, 0xffffffff, t87;\n\t"
        "madc.lo.cc.u32 t88, t68, 0xffffffff, t88;\n\t"
        "madc.hi.cc.u32 t89, t68, 0xffffffff, t89;\n\t"
        "madc.lo.cc.u32 t90, t70, 0xffffffff, t90;\n\t"
        "madc.hi.cc.u32 t91, t70, 0xffffffff, t91;\n\t"
        "addc.u32 t92, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t101, t65, 0xffffffff, t101;\n\t"
        "madc.hi.cc.u32 t102, t65, 0xffffffff, t102;\n\t"
        "madc.lo.cc.u32 t103, t67, 0xffffffff, t103;\n\t"
        "madc.hi.cc.u32 t104, t67, 0xffffffff, t104;\n\t"
        "madc.lo.cc.u32 t105, t69, 0xffffffff, t105;\n\t"
        "madc.hi.cc.u32 t106, t69, 0xffffffff, t106;\n\t"
        "madc.lo.cc.u32 t107, t79, 0xffffffff, t107;\n\t"
        "madc.hi.u32 t108, t79, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t86, t65, 0xffffffff, t86;\n\t"
        "madc.hi.cc.u32 t87, t65, 0xffffffff, t87;\n\t"
        "madc.lo.cc.u32 t88, t67, 0xffffffff, t88;\n\t"
        "madc.hi.cc.u32 t89, t67, 0xffffffff, t89;\n\t"
        "madc.lo.cc.u32 t90, t69, 0xffffffff, t90;\n\t"
        "madc.hi.cc.u32 t91, t69, 0xffffffff, t91;\n\t"
        "madc.lo.cc.u32 t92, t79, 0xffffffff, t92;\n\t"
        "madc.hi.u32 t93, t79, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t101, t47, 0xffffffff, t101;\n\t"
        "madc.hi.cc.u32 t102, t47, 0xffffffff, t102;\n\t"
        "madc.lo.cc.u32 t103, t66, 0xffffffff, t103;\n\t"
        "madc.hi.cc.u32 t104, t66, 0xffffffff, t104;\n\t"
        "madc.lo.cc.u32 t105, t68, 0xffffffff, t105;\n\t"
        "madc.hi.cc.u32 t106, t68, 0xffffffff, t106;\n\t"
        "madc.lo.cc.u32 t107, t70, 0xffffffff, t107;\n\t"
        "madc.hi.cc.u32 t108, t70, 0xffffffff, t108;\n\t"
        "addc.u32 t109, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t86, t47, 0x0, t86;\n\t"
        "madc.hi.cc.u32 t87, t47, 0x0, t87;\n\t"
        "madc.lo.cc.u32 t88, t66, 0x0, t88;\n\t"
        "madc.hi.cc.u32 t89, t66, 0x0, t89;\n\t"
        "madc.lo.cc.u32 t90, t68, 0x0, t90;\n\t"
        "madc.hi.cc.u32 t91, t68, 0x0, t91;\n\t"
        "madc.lo.cc.u32 t92, t70, 0x0, t92;\n\t"
        "madc.hi.cc.u32 t93, t70, 0x0, t93;\n\t"
        "addc.u32 t94, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t103, t65, 0x0, t103;\n\t"
        "madc.hi.cc.u32 t104, t65, 0x0, t104;\n\t"
        "madc.lo.cc.u32 t105, t67, 0x0, t105;\n\t"
        "madc.hi.cc.u32 t106, t67, 0x0, t106;\n\t"
        "madc.lo.cc.u32 t107, t69, 0x0, t107;\n\t"
        "madc.hi.cc.u32 t108, t69, 0x0, t108;\n\t"
        "madc.lo.cc.u32 t109, t79, 0x0, t109;\n\t"
        "madc.hi.u32 t110, t79, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t88, t65, 0xffffffff, t88;\n\t"
        "madc.hi.cc.u32 t89, t65, 0xffffffff, t89;\n\t"
        "madc.lo.cc.u32 t90, t67, 0xffffffff, t90;\n\t"
        "madc.hi.cc.u32 t91, t67, 0xffffffff, t91;\n\t"
        "madc.lo.cc.u32 t92, t69, 0xffffffff, t92;\n\t"
        "madc.hi.cc.u32 t93, t69, 0xffffffff, t93;\n\t"
        "madc.lo.cc.u32 t94, t79, 0xffffffff, t94;\n\t"
        "madc.hi.u32 t95, t79, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t103, t47, 0xffffffff, t103;\n\t"
        "madc.hi.cc.u32 t104, t47, 0xffffffff, t104;\n\t"
        "madc.lo.cc.u32 t105, t66, 0xffffffff, t105;\n\t"
        "madc.hi.cc.u32 t106, t66, 0xffffffff, t106;\n\t"
        "madc.lo.cc.u32 t107, t68, 0xffffffff, t107;\n\t"
        "madc.hi.cc.u32 t108, t68, 0xffffffff, t108;\n\t"
        "madc.lo.cc.u32 t109, t70, 0xffffffff, t109;\n\t"
        "madc.hi.cc.u32 t110, t70, 0xffffffff, t110;\n\t"
        "addc.u32 t111, 0x0, 0x0;\n\t"
        "add.cc.u32 t112, t81, t97;\n\t"
        "addc.cc.u32 t113, t82, t98;\n\t"
        "addc.cc.u32 t114, t83, t99;\n\t"
        "addc.cc.u32 t115, t84, t100;\n\t"
        "addc.cc.u32 t116, t85, t101;\n\t"
        "addc.cc.u32 t117, t86, t102;\n\t"
        "addc.cc.u32 t118, t87, t103;\n\t"
        "addc.cc.u32 t119, t88, t104;\n\t"
        "addc.cc.u32 t120, t89, t105;\n\t"
        "addc.cc.u32 t121, t90, t106;\n\t"
        "addc.cc.u32 t122, t91, t107;\n\t"
        "addc.cc.u32 t123, t92, t108;\n\t"
        "addc.cc.u32 t124, t93, t109;\n\t"
        "addc.cc.u32 t125, t94, t110;\n\t"
        "addc.u32 t126, t95, t111;\n\t"
        "add.cc.u32 t127, t0, t80;\n\t"
        "addc.cc.u32 t128, t32, t112;\n\t"
        "addc.cc.u32 t129, t33, t113;\n\t"
        "addc.cc.u32 t130, t34, t114;\n\t"
        "addc.cc.u32 t131, t35, t115;\n\t"
        "addc.cc.u32 t132, t36, t116;\n\t"
        "addc.cc.u32 t133, t37, t117;\n\t"
        "addc.cc.u32 t134, t38, t118;\n\t"
        "addc.cc.u32 t135, t39, t119;\n\t"
        "addc.cc.u32 t136, t40, t120;\n\t"
        "addc.cc.u32 t137, t41, t121;\n\t"
        "addc.cc.u32 t138, t42, t122;\n\t"
        "addc.cc.u32 t139, t43, t123;\n\t"
        "addc.cc.u32 t140, t44, t124;\n\t"
        "addc.cc.u32 t141, t45, t125;\n\t"
        "addc.cc.u32 t142, t46, t126;\n\t"
        "addc.u32 t143, 0x0, 0x0;\n\t"
        "sub.cc.u32 t144, t135, 0xfc632551;\n\t"
        "subc.cc.u32 t145, t136, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t146, t137, 0xa7179e84;\n\t"
        "subc.cc.u32 t147, t138, 0xbce6faad;\n\t"
        "subc.cc.u32 t148, t139, 0xffffffff;\n\t"
        "subc.cc.u32 t149, t140, 0xffffffff;\n\t"
        "subc.cc.u32 t150, t141, 0x0;\n\t"
        "subc.cc.u32 t151, t142, 0xffffffff;\n\t"
        "subc.cc.u32 t152, t143, 0x0;\n\t"
        "subc.u32 t153, 0x0, 0x0;\n\t"
        "xor.b32 t154, t144, t135;\n\t"
        "and.b32 t155, t154, t153;\n\t"
        "xor.b32 t156, t155, t144;\n\t"
        "xor.b32 t157, t145, t136;\n\t"
        "and.b32 t158, t157, t153;\n\t"
        "xor.b32 t159, t158, t145;\n\t"
        "xor.b32 t160, t146, t137;\n\t"
        "and.b32 t161, t160, t153;\n\t"
        "xor.b32 t162, t161, t146;\n\t"
        "xor.b32 t163, t147, t138;\n\t"
        "and.b32 t164, t163, t153;\n\t"
        "xor.b32 t165, t164, t147;\n\t"
        "xor.b32 t166, t148, t139;\n\t"
        "and.b32 t167, t166, t153;\n\t"
        "xor.b32 t168, t167, t148;\n\t"
        "xor.b32 t169, t149, t140;\n\t"
        "and.b32 t170, t169, t153;\n\t"
        "xor.b32 t171, t170, t149;\n\t"
        "xor.b32 t172, t150, t141;\n\t"
        "and.b32 t173, t172, t153;\n\t"
        "xor.b32 t174, t173, t150;\n\t"
        "xor.b32 t175, t151, t142;\n\t"
        "and.b32 t176, t175, t153;\n\t"
        "xor.b32 t177, t176, t151;\n\t"
        "mov.u32 %0, t156;\n\t"
        "mov.u32 %1, t159;\n\t"
        "mov.u32 %2, t162;\n\t"
        "mov.u32 %3, t165;\n\t"
        "mov.u32 %4, t168;\n\t"
        "mov.u32 %5, t171;\n\t"
        "mov.u32 %6, t174;\n\t"
        "mov.u32 %7, t177;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127, t128, t129, t130, t131, t132, t133, t134, t135, t136, t137, t138, t139, t140, t141, t142, t143, t144, t145, t146, t147, t148, t149, t150, t151, t152, t153, t154, t155, t156, t157, t158, t159, t160, t161, t162, t163, t164, t165, t166, t167, t168, t169, t170, t171, t172, t173, t174, t175, t176, t177;
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
    t47 = (uint32_t)((uint32_t)(t0 * 0xee00bc4fu));
    t48 = (uint32_t)(((uint64_t)t0 * 0xee00bc4fu) >> 32);
    t49 = (uint32_t)((uint32_t)(t33 * 0xee00bc4fu));
    t50 = (uint32_t)(((uint64_t)t33 * 0xee00bc4fu) >> 32);
    t51 = (uint32_t)((uint32_t)(t35 * 0xee00bc4fu));
    t52 = (uint32_t)(((uint64_t)t35 * 0xee00bc4fu) >> 32);
    t53 = (uint32_t)((uint32_t)(t37 * 0xee00bc4fu));
    t54 = (uint32_t)(((uint64_t)t37 * 0xee00bc4fu) >> 32);
    t57 = (uint32_t)((uint32_t)(t32 * 0xee00bc4fu));
    t58 = (uint32_t)(((uint64_t)t32 * 0xee00bc4fu) >> 32);
    t59 = (uint32_t)((uint32_t)(t34 * 0xee00bc4fu));
    t60 = (uint32_t)(((uint64_t)t34 * 0xee00bc4fu) >> 32);
    t61 = (uint32_t)((uint32_t)(t36 * 0xee00bc4fu));
    t62 = (uint32_t)(((uint64_t)t36 * 0xee00bc4fu) >> 32);
    w_ = (uint64_t)(uint32_t)(t32 * 0xccd1c8aau) + t49; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t32 * 0xccd1c8aau) >> 32) + t50 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t34 * 0xccd1c8aau) + t51 + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t34 * 0xccd1c8aau) >> 32) + t52 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t36 * 0xccd1c8aau) + t53 + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t36 * 0xccd1c8aau) >> 32) + t54 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t55 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0xccd1c8aau) + t57; t57 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0xccd1c8aau) >> 32) + t58 + cf_; t58 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t33 * 0xccd1c8aau) + t59 + cf_; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t33 * 0xccd1c8aau) >> 32) + t60 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t35 * 0xccd1c8aau) + t61 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t35 * 0xccd1c8aau) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t63 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x7d74d2e4u) + t49; t49 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0x7d74d2e4u) >> 32) + t50 + cf_; t50 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t33 * 0x7d74d2e4u) + t51 + cf_; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t33 * 0x7d74d2e4u) >> 32) + t52 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t35 * 0x7d74d2e4u) + t53 + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t35 * 0x7d74d2e4u) >> 32) + t54 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t55 + 0x0u + cf_; t55 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t32 * 0x7d74d2e4u) + t59; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t32 * 0x7d74d2e4u) >> 32) + t60 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t34 * 0x7d74d2e4u) + t61 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t34 * 0x7d74d2e4u) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + 0x0u + cf_; t63 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t32 * 0x48c94408u) + t51; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t32 * 0x48c94408u) >> 32) + t52 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t34 * 0x48c94408u) + t53 + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t34 * 0x48c94408u) >> 32) + t54 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t55 + 0x0u + cf_; t55 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x48c94408u) + t59; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0x48c94408u) >> 32) + t60 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t33 * 0x48c94408u) + t61 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t33 * 0x48c94408u) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + 0x0u + cf_; t63 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0xc588c6f6u) + t51; t51 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0xc588c6f6u) >> 32) + t52 + cf_; t52 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t33 * 0xc588c6f6u) + t53 + cf_; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t33 * 0xc588c6f6u) >> 32) + t54 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t55 + 0x0u + cf_; t55 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t32 * 0xc588c6f6u) + t61; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t32 * 0xc588c6f6u) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + 0x0u + cf_; t63 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t32 * 0x50fe77ecu) + t53; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t32 * 0x50fe77ecu) >> 32) + t54 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t55 + 0x0u + cf_; t55 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x50fe77ecu) + t61; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0x50fe77ecu) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + 0x0u + cf_; t63 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0xa9d6281cu) + t53; t53 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0xa9d6281cu) >> 32) + t54 + cf_; t54 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t55 + 0x0u + cf_; t55 = (uint32_t)w_;
    w_ = (uint64_t)t48 + t57; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t49 + t58 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t50 + t59 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t51 + t60 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t52 + t61 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t53 + t62 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t54 + t63 + cf_; t71 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t38 * 0xee00bc4fu) + t71; t72 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t37 * 0xccd1c8aau) + t72; t73 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t36 * 0x7d74d2e4u) + t73; t74 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t35 * 0x48c94408u) + t74; t75 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t34 * 0xc588c6f6u) + t75; t76 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t33 * 0x50fe77ecu) + t76; t77 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t32 * 0xa9d6281cu) + t77; t78 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x60d06633u) + t78; t79 = (uint32_t)w_;
    t80 = (uint32_t)((uint32_t)(t47 * 0xfc632551u));
    t81 = (uint32_t)(((uint64_t)t47 * 0xfc632551u) >> 32);
    t82 = (uint32_t)((uint32_t)(t66 * 0xfc632551u));
    t83 = (uint32_t)(((uint64_t)t66 * 0xfc632551u) >> 32);
    t84 = (uint32_t)((uint32_t)(t68 * 0xfc632551u));
    t85 = (uint32_t)(((uint64_t)t68 * 0xfc632551u) >> 32);
    t86 = (uint32_t)((uint32_t)(t70 * 0xfc632551u));
    t87 = (uint32_t)(((uint64_t)t70 * 0xfc632551u) >> 32);
    t97 = (uint32_t)((uint32_t)(t65 * 0xfc632551u));
    t98 = (uint32_t)(((uint64_t)t65 * 0xfc632551u) >> 32);
    t99 = (uint32_t)((uint32_t)(t67 * 0xfc632551u));
    t100 = (uint32_t)(((uint64_t)t67 * 0xfc632551u) >> 32);
    t101 = (uint32_t)((uint32_t)(t69 * 0xfc632551u));
    t102 = (uint32_t)(((uint64_t)t69 * 0xfc632551u) >> 32);
    t103 = (uint32_t)((uint32_t)(t79 * 0xfc632551u));
    t104 = (uint32_t)(((uint64_t)t79 * 0xfc632551u) >> 32);
    w_ = (uint64_t)(uint32_t)(t65 * 0xf3b9cac2u) + t82; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t65 * 0xf3b9cac2u) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t67 * 0xf3b9cac2u) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t67 * 0xf3b9cac2u) >> 32) + t85 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t69 * 0xf3b9cac2u) + t86 + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t69 * 0xf3b9cac2u) >> 32) + t87 + cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t79 * 0xf3b9cac2u) + 0x0u + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t79 * 0xf3b9cac2u) >> 32) + 0x0u + cf_; t89 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t47 * 0xf3b9cac2u) + t97; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xf3b9cac2u) >> 32) + t98 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t66 * 0xf3b9cac2u) + t99 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t66 * 0xf3b9cac2u) >> 32) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t68 * 0xf3b9cac2u) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t68 * 0xf3b9cac2u) >> 32) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t70 * 0xf3b9cac2u) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t70 * 0xf3b9cac2u) >> 32) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t105 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t47 * 0xa7179e84u) + t82; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xa7179e84u) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t66 * 0xa7179e84u) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t66 * 0xa7179e84u) >> 32) + t85 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t68 * 0xa7179e84u) + t86 + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t68 * 0xa7179e84u) >> 32) + t87 + cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t70 * 0xa7179e84u) + t88 + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t70 * 0xa7179e84u) >> 32) + t89 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t90 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t65 * 0xa7179e84u) + t99; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t65 * 0xa7179e84u) >> 32) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t67 * 0xa7179e84u) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t67 * 0xa7179e84u) >> 32) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t69 * 0xa7179e84u) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t69 * 0xa7179e84u) >> 32) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t79 * 0xa7179e84u) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t79 * 0xa7179e84u) >> 32) + 0x0u + cf_; t106 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t65 * 0xbce6faadu) + t84; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t65 * 0xbce6faadu) >> 32) + t85 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t67 * 0xbce6faadu) + t86 + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t67 * 0xbce6faadu) >> 32) + t87 + cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t69 * 0xbce6faadu) + t88 + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t69 * 0xbce6faadu) >> 32) + t89 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t79 * 0xbce6faadu) + t90 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t79 * 0xbce6faadu) >> 32) + 0x0u + cf_; t91 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t47 * 0xbce6faadu) + t99; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xbce6faadu) >> 32) + t100 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t66 * 0xbce6faadu) + t101 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t66 * 0xbce6faadu) >> 32) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t68 * 0xbce6faadu) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t68 * 0xbce6faadu) >> 32) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t70 * 0xbce6faadu) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t70 * 0xbce6faadu) >> 32) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t107 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t47 * 0xffffffffu) + t84; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xffffffffu) >> 32) + t85 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t66 * 0xffffffffu) + t86 + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t66 * 0xffffffffu) >> 32) + t87 + cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t68 * 0xffffffffu) + t88 + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t68 * 0xffffffffu) >> 32) + t89 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t70 * 0xffffffffu) + t90 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t70 * 0xffffffffu) >> 32) + t91 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t92 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t65 * 0xffffffffu) + t101; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t65 * 0xffffffffu) >> 32) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t67 * 0xffffffffu) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t67 * 0xffffffffu) >> 32) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t69 * 0xffffffffu) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t69 * 0xffffffffu) >> 32) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t79 * 0xffffffffu) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t79 * 0xffffffffu) >> 32) + 0x0u + cf_; t108 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t65 * 0xffffffffu) + t86; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t65 * 0xffffffffu) >> 32) + t87 + cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t67 * 0xffffffffu) + t88 + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t67 * 0xffffffffu) >> 32) + t89 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t69 * 0xffffffffu) + t90 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t69 * 0xffffffffu) >> 32) + t91 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t79 * 0xffffffffu) + t92 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t79 * 0xffffffffu) >> 32) + 0x0u + cf_; t93 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t47 * 0xffffffffu) + t101; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xffffffffu) >> 32) + t102 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t66 * 0xffffffffu) + t103 + cf_; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t66 * 0xffffffffu) >> 32) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t68 * 0xffffffffu) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t68 * 0xffffffffu) >> 32) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t70 * 0xffffffffu) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t70 * 0xffffffffu) >> 32) + t108 + cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t109 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t47 * 0x0u) + t86; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0x0u) >> 32) + t87 + cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t66 * 0x0u) + t88 + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t66 * 0x0u) >> 32) + t89 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t68 * 0x0u) + t90 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t68 * 0x0u) >> 32) + t91 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t70 * 0x0u) + t92 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t70 * 0x0u) >> 32) + t93 + cf_; t93 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t94 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t65 * 0x0u) + t103; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t65 * 0x0u) >> 32) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t67 * 0x0u) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t67 * 0x0u) >> 32) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t69 * 0x0u) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t69 * 0x0u) >> 32) + t108 + cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t79 * 0x0u) + t109 + cf_; t109 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t79 * 0x0u) >> 32) + 0x0u + cf_; t110 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t65 * 0xffffffffu) + t88; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t65 * 0xffffffffu) >> 32) + t89 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t67 * 0xffffffffu) + t90 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t67 * 0xffffffffu) >> 32) + t91 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t69 * 0xffffffffu) + t92 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t69 * 0xffffffffu) >> 32) + t93 + cf_; t93 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t79 * 0xffffffffu) + t94 + cf_; t94 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t79 * 0xffffffffu) >> 32) + 0x0u + cf_; t95 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t47 * 0xffffffffu) + t103; t103 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xffffffffu) >> 32) + t104 + cf_; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t66 * 0xffffffffu) + t105 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t66 * 0xffffffffu) >> 32) + t106 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t68 * 0xffffffffu) + t107 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t68 * 0xffffffffu) >> 32) + t108 + cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t70 * 0xffffffffu) + t109 + cf_; t109 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t70 * 0xffffffffu) >> 32) + t110 + cf_; t110 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t111 = (uint32_t)w_;
    w_ = (uint64_t)t81 + t97; t112 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t82 + t98 + cf_; t113 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t83 + t99 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t84 + t100 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t85 + t101 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t86 + t102 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t87 + t103 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t88 + t104 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t89 + t105 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t90 + t106 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t91 + t107 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t92 + t108 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t93 + t109 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t94 + t110 + cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t95 + t111 + cf_; t126 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t80; t127 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + t112 + cf_; t128 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t33 + t113 + cf_; t129 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t34 + t114 + cf_; t130 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t35 + t115 + cf_; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t36 + t116 + cf_; t132 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t37 + t117 + cf_; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t38 + t118 + cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t39 + t119 + cf_; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + t120 + cf_; t136 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t41 + t121 + cf_; t137 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t42 + t122 + cf_; t138 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t43 + t123 + cf_; t139 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t44 + t124 + cf_; t140 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t45 + t125 + cf_; t141 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t46 + t126 + cf_; t142 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t143 = (uint32_t)w_;
    w_ = (uint64_t)t135 - 0xfc632551u; t144 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t136 - 0xf3b9cac2u - cf_; t145 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t137 - 0xa7179e84u - cf_; t146 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t138 - 0xbce6faadu - cf_; t147 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t139 - 0xffffffffu - cf_; t148 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t140 - 0xffffffffu - cf_; t149 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t141 - 0x0u - cf_; t150 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t142 - 0xffffffffu - cf_; t151 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t143 - 0x0u - cf_; t152 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t153 = (uint32_t)w_;
    t154 = (uint32_t)(t144 ^ t135);
    t155 = (uint32_t)(t154 & t153);
    t156 = (uint32_t)(t155 ^ t144);
    t157 = (uint32_t)(t145 ^ t136);
    t158 = (uint32_t)(t157 & t153);
    t159 = (uint32_t)(t158 ^ t145);
    t160 = (uint32_t)(t146 ^ t137);
    t161 = (uint32_t)(t160 & t153);
    t162 = (uint32_t)(t161 ^ t146);
    t163 = (uint32_t)(t147 ^ t138);
    t164 = (uint32_t)(t163 & t153);
    t165 = (uint32_t)(t164 ^ t147);
    t166 = (uint32_t)(t148 ^ t139);
    t167 = (uint32_t)(t166 & t153);
    t168 = (uint32_t)(t167 ^ t148);
    t169 = (uint32_t)(t149 ^ t140);
    t170 = (uint32_t)(t169 & t153);
    t171 = (uint32_t)(t170 ^ t149);
    t172 = (uint32_t)(t150 ^ t141);
    t173 = (uint32_t)(t172 & t153);
    t174 = (uint32_t)(t173 ^ t150);
    t175 = (uint32_t)(t151 ^ t142);
    t176 = (uint32_t)(t175 & t153);
    t177 = (uint32_t)(t176 ^ t151);
    r[0] = t156;
    r[1] = t159;
    r[2] = t162;
    r[3] = t165;
    r[4] = t168;
    r[5] = t171;
    r[6] = t174;
    r[7] = t177;
#endif
  }

  // c = a*a (pseudo.py:663-702 / monty.py:982-1165)
  static MAB_DEV void sqr(uint32_t (&r)[8], const uint32_t (&a)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<208>;\n\t"
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
        "mul.lo.u32 t77, t61, 0xee00bc4f;\n\t"
        "mul.hi.u32 t78, t61, 0xee00bc4f;\n\t"
        "mul.lo.u32 t79, t63, 0xee00bc4f;\n\t"
        "mul.hi.u32 t80, t63, 0xee00bc4f;\n\t"
        "mul.lo.u32 t81, t65, 0xee00bc4f;\n\t"
        "mul.hi.u32 t82, t65, 0xee00bc4f;\n\t"
        "mul.lo.u32 t83, t67, 0xee00bc4f;\n\t"
        "mul.hi.u32 t84, t67, 0xee00bc4f;\n\t"
        "mul.lo.u32 t87, t62, 0xee00bc4f;\n\t"
        "mul.hi.u32 t88, t62, 0xee00bc4f;\n\t"
        "mul.lo.u32 t89, t64, 0xee00bc4f;\n\t"
        "mul.hi.u32 t90, t64, 0xee00bc4f;\n\t"
        "mul.lo.u32 t91, t66, 0xee00bc4f;\n\t"
        "mul.hi.u32 t92, t66, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t79, t62, 0xccd1c8aa, t79;\n\t"
        "madc.hi.cc.u32 t80, t62, 0xccd1c8aa, t80;\n\t"
        "madc.lo.cc.u32 t81, t64, 0xccd1c8aa, t81;\n\t"
        "madc.hi.cc.u32 t82, t64, 0xccd1c8aa, t82;\n\t"
        "madc.lo.cc.u32 t83, t66, 0xccd1c8aa, t83;\n\t"
        "madc.hi.cc.u32 t84, t66, 0xccd1c8aa, t84;\n\t"
        "addc.u32 t85, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t87, t61, 0xccd1c8aa, t87;\n\t"
        "madc.hi.cc.u32 t88, t61, 0xccd1c8aa, t88;\n\t"
        "madc.lo.cc.u32 t89, t63, 0xccd1c8aa, t89;\n\t"
        "madc.hi.cc.u32 t90, t63, 0xccd1c8aa, t90;\n\t"
        "madc.lo.cc.u32 t91, t65, 0xccd1c8aa, t91;\n\t"
        "madc.hi.cc.u32 t92, t65, 0xccd1c8aa, t92;\n\t"
        "addc.u32 t93, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t79, t61, 0x7d74d2e4, t79;\n\t"
        "madc.hi.cc.u32 t80, t61, 0x7d74d2e4, t80;\n\t"
        "madc.lo.cc.u32 t81, t63, 0x7d74d2e4, t81;\n\t"
        "madc.hi.cc.u32 t82, t63, 0x7d74d2e4, t82;\n\t"
        "madc.lo.cc.u32 t83, t65, 0x7d74d2e4, t83;\n\t"
        "madc.hi.cc.u32 t84, t65, 0x7d74d2e4, t84;\n\t"
        "addc.u32 t85, t85, 0x0;\n\t"
        "mad.lo.cc.u32 t89, t62, 0x7d74d2e4, t89;\n\t"
        "madc.hi.cc.u32 t90, t62, 0x7d74d2e4, t90;\n\t"
        "madc.lo.cc.u32 t91, t64, 0x7d74d2e4, t91;\n\t"
        "madc.hi.cc.u32 t92, t64, 0x7d74d2e4, t92;\n\t"
        "addc.u32 t93, t93, 0x0;\n\t"
        "mad.lo.cc.u32 t81, t62, 0x48c94408, t81;\n\t"
        "madc.hi.cc.u32 t82, t62, 0x48c94408, t82;\n\t"
        "madc.lo.cc.u32 t83, t64, 0x48c94408, t83;\n\t"
        "madc.hi.cc.u32 t84, t64, 0x48c94408, t84;\n\t"
        "addc.u32 t85, t85, 0x0;\n\t"
        "mad.lo.cc.u32 t89, t61, 0x48c94408, t89;\n\t"
        "madc.hi.cc.u32 t90, t61, 0x48c94408, t90;\n\t"
        "madc.lo.cc.u32 t91, t63, 0x48c94408, t91;\n\t"
        "madc.hi.cc.u32 t92, t63, 0x48c94408, t92;\n\t"
        "addc.u32 t93, t93, 0x0;\n\t"
        "mad.lo.cc.u32 t81, t61, 0xc588c6f6, t81;\n\t"
        "madc.hi.cc.u32 t82, t61, 0xc588c6f6, t82;\n\t"
        "madc.lo.cc.u32 t83, t63, 0xc588c6f6, t83;\n\t"
        "madc.hi.cc.u32 t84, t63, 0xc588c6f6, t84;\n\t"
        "addc.u32 t85, t85, 0x0;\n\t"
        "mad.lo.cc.u32 t91, t62, 0xc588c6f6, t91;\n\t"
        "madc.hi.cc.u32 t92, t62, 0xc588c6f6, t92;\n\t"
        "addc.u32 t93, t93, 0x0;\n\t"
        "mad.lo.cc.u32 t83, t62, 0x50fe77ec, t83;\n\t"
        "madc.hi.cc.u32 t84, t62, 0x50fe77ec, t84;\n\t"
        "addc.u32 t85, t85, 0x0;\n\t"
        "mad.lo.cc.u32 t91, t61, 0x50fe77ec, t91;\n\t"
        "madc.hi.cc.u32 t92, t61, 0x50fe77ec, t92;\n\t"
        "addc.u32 t93, t93, 0x0;\n\t"
        "mad.lo.cc.u32 t83, t61, 0xa9d6281c, t83;\n\t"
        "madc.hi.cc.u32 t84, t61, 0xa9d6281c, t84;\n\t"
        "addc.u32 t85, t85, 0x0;\n\t"
        "add.cc.u32 t95, t78, t87;\n\t"
        "addc.cc.u32 t96, t79, t88;\n\t"
        "addc.cc.u32 t97, t80, t89;\n\t"
        "addc.cc.u32 t98, t81, t90;\n\t"
        "addc.cc.u32 t99, t82, t91;\n\t"
        "addc.cc.u32 t100, t83, t92;\n\t"
        "addc.u32 t101, t84, t93;\n\t"
        "mad.lo.u32 t102, t68, 0xee00bc4f, t101;\n\t"
        "mad.lo.u32 t103, t67, 0xccd1c8aa, t102;\n\t"
        "mad.lo.u32 t104, t66, 0x7d74d2e4, t103;\n\t"
        "mad.lo.u32 t105, t65, 0x48c94408, t104;\n\t"
        "mad.lo.u32 t106, t64, 0xc588c6f6, t105;\n\t"
        "mad.lo.u32 t107, t63, 0x50fe77ec, t106;\n\t"
        "mad.lo.u32 t108, t62, 0xa9d6281c, t107;\n\t"
        "mad.lo.u32 t109, t61, 0x60d06633, t108;\n\t"
        "mul.lo.u32 t110, t77, 0xfc632551;\n\t"
        "mul.hi.u32 t111, t77, 0xfc632551;\n\t"
        "mul.lo.u32 t112, t96, 0xfc632551;\n\t"
        "mul.hi.u32 t113, t96, 0xfc632551;\n\t"
        "mul.lo.u32 t114, t98, 0xfc632551;\n\t"
        "mul.hi.u32 t115, t98, 0xfc632551;\n\t"
        "mul.lo.u32 t116, t100, 0xfc632551;\n\t"
        "mul.hi.u32 t117, t100, 0xfc632551;\n\t"
        "mul.lo.u32 t127, t95, 0xfc632551;\n\t"
        "mul.hi.u32 t128, t95, 0xfc632551;\n\t"
        "mul.lo.u32 t129, t97, 0xfc632551;\n\t"
        "mul.hi.u32 t130, t97, 0xfc632551;\n\t"
        "mul.lo.u32 t131, t99, 0xfc632551;\n\t"
        "mul.hi.u32 t132, t99, 0xfc632551;\n\t"
        "mul.lo.u32 t133, t109, 0xfc632551;\n\t"
        "mul.hi.u32 t134, t109, 0xfc632551;\n\t"
        "mad.lo.cc.u32 t112, t95, 0xf3b9cac2, t112;\n\t"
        "madc.hi.cc.u32 t113, t95, 0xf3b9cac2, t113;\n\t"
        "madc.lo.cc.u32 t114, t97, 0xf3b9cac2, t114;\n\t"
        "madc.hi.cc.u32 t115, t97, 0xf3b9cac2, t115;\n\t"
        "madc.lo.cc.u32 t116, t99, 0xf3b9cac2, t116;\n\t"
        "madc.hi.cc.u32 t117, t99, 0xf3b9cac2, t117;\n\t"
        "madc.lo.cc.u32 t118, t109, 0xf3b9cac2, 0x0;\n\t"
        "madc.hi.u32 t119, t109, 0xf3b9cac2, 0x0;\n\t"
        "mad.lo.cc.u32 t127, t77, 0xf3b9cac2, t127;\n\t"
        "madc.hi.cc.u32 t128, t77, 0xf3b9cac2, t128;\n\t"
        "madc.lo.cc.u32 t129, t96, 0xf3b9cac2, t129;\n\t"
        "madc.hi.cc.u32 t130, t96, 0xf3b9cac2, t130;\n\t"
        "madc.lo.cc.u32 t131, t98, 0xf3b9cac2, t131;\n\t"
        "madc.hi.cc.u32 t132, t98, 0xf3b9cac2, t132;\n\t"
        "madc.lo.cc.u32 t133, t100, 0xf3b9cac2, t133;\n\t"
        "madc.hi.cc.u32 t134, t100, 0xf3b9cac2, t134;\n\t"
        "addc.u32 t135, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t112, t77, 0xa7179e84, t112;\n\t"
        "madc.hi.cc.u32 t113, t77, 0xa7179e84, t113;\n\t"
        "madc.lo.cc.u32 t114, t96, 0xa7179e84, t114;\n\t"
        "madc.hi.cc.u32 t115, t96, 0xa7179e84, t115;\n\t"
        "madc.lo.cc.u32 t116, t98, 0xa7179e84, t116;\n\t"
        "madc.hi.cc.u32 t117, t98, 0xa7179e84, t117;\n\t"
        "madc.lo.cc.u32 t118, t100, 0xa7179e84, t118;\n\t"
        "madc.hi.cc.u32 t119, t100, 0xa7179e84, t119;\n\t"
        "addc.u32 t120, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t129, t95, 0xa7179e84, t129;\n\t"
        "madc.hi.cc.u32 t130, t95, 0xa7179e84, t130;\n\t"
        "madc.lo.cc.u32 t131, t97, 0xa7179e84, t131;\n\t"
        "madc.hi.cc.u32 t132, t97, 0xa7179e84, t132;\n\t"
        "madc.lo.cc.u32 t133, t99, 0xa7179e84, t133;\n\t"
        "madc.hi.cc.u32 t134, t99, 0xa7179e84, t134;\n\t"
        "madc.lo.cc.u32 t135, t109, 0xa7179e84, t135;\n\t"
        "madc.hi.u32 t136, t109, 0xa7179e84, 0x0;\n\t"
        "mad.lo.cc.u32 t114, t95, 0xbce6faad, t114;\n\t"
        "madc.hi.cc.u32 t115, t95, 0xbce6faad, t115;\n\t"
        "madc.lo.cc.u32 t116, t97, 0xbce6faad, t116;\n\t"
        "madc.hi.cc.u32 t117, t97, 0xbce6faad, t117;\n\t"
        "madc.lo.cc.u32 t118, t99, 0xbce6faad, t118;\n\t"
        "madc.hi.cc.u32 t119, t99, 0xbce6faad, t119;\n\t"
        "madc.lo.cc.u32 t120, t109, 0xbce6faad, t120;\n\t"
        "madc.hi.u32 t121, t109, 0xbce6faad, 0x0;\n\t"
        "mad.lo.cc.u32 t129, t77, 0xbce6faad, t129;\n\t"
        "madc.hi.cc.u32 t130, t77, 0xbce6faad, t130;\n\t"
        "madc.lo.cc.u32 t131, t96, 0xbce6faad, t131;\n\t"
        "madc.hi.cc.u32 t132, t96, 0xbce6faad, t132;\n\t"
        "madc.lo.cc.u32 t133, t98, 0xbce6faad, t133;\n\t"
        "madc.hi.cc.u32 t134, t98, 0xbce6faad, t134;\n\t"
        "madc.lo.cc.u32 t135, t100, 0xbce6faad, t135;\n\t"
        "madc.hi.cc.u32 t136, t100, 0xbce6faad, t136;\n\t"
        "addc.u32 t137, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t114, t77, 0xffffffff, t114;\n\t"
        "madc.hi.cc.u32 t115, t77, 0xffffffff, t115;\n\t"
        "madc.lo.cc.u32 t116, t96, 0xffffffff, t116;\n\t"
        "madc.hi.cc.u32 t117, t96, 0xffffffff, t117;\n\t"
        "madc.lo.cc.u32 t118, t98, 0xffffffff, t118;\n\t"
        "madc.hi.cc.u32 t119, t98, 0xffffffff, t119;\n\t"
        "madc.lo.cc.u32 t120, t100, 0xffffffff, t120;\n\t"
        "madc.hi.cc.u32 t121, t100, 0xffffffff, t121;\n\t"
        "addc.u32 t122, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t131, t95, 0xffffffff, t131;\n\t"
        "madc.hi.cc.u32 t132, t95, 0xffffffff, t132;\n\t"
        "madc.lo.cc.u32 t133, t97, 0xffffffff, t133;\n\t"
        "madc.hi.cc.u32 t134, t97, 0xffffffff, t134;\n\t"
        "madc.lo.cc.u32 t135, t99, 0xffffffff, t135;\n\t"
        "madc.hi.cc.u32 t136, t99, 0xffffffff, t136;\n\t"
        "madc.lo.cc.u32 t137, t109, 0xffffffff, t137;\n\t"
        "madc.hi.u32 t138, t109, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t116, t95, 0xffffffff, t116;\n\t"
        "madc.hi.cc.u32 t117, t95, 0xffffffff, t117;\n\t"
        "madc.lo.cc.u32 t118, t97, 0xffffffff, t118;\n\t"
        "madc.hi.cc.u32 t119, t97, 0xffffffff, t119;\n\t"
        "madc.lo.cc.u32 t120, t99, 0xffffffff, t120;\n\t"
        "madc.hi.cc.u32 t121, t99, 0xffffffff, t121;\n\t"
        "madc.lo.cc.u32 t122, t109, 0xffffffff, t122;\n\t"
        "madc.hi.u32 t123, t109, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t131, t77, 0xffffffff, t131;\n\t"
        "madc.hi.cc.u32 t132, t77, 0xffffffff, t132;\n\t"
        "madc.lo.cc.u32 t133, t96, 0xffffffff, t133;\n\t"
        "madc.hi.cc.u32 t134, t96, 0xffffffff, t134;\n\t"
        "madc.lo.cc.u32 t135, t98, 0xffffffff, t135;\n\t"
        "madc.hi.cc.u32 t136, t98, 0xffffffff, t136;\n\t"
        "madc.lo.cc.u32 t137, t100, 0xffffffff, t137;\n\t"
        "madc.hi.cc.u32 t138, t100, 0xffffffff, t138;\n\t"
        "addc.u32 t139, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t116, t77, 0x0, t116;\n\t"
        "madc.hi.cc.u32 t117, t77, 0x0, t117;\n\t"
        "madc.lo.cc.u32 t118, t96, 0x0, t118;\n\t"
        "madc.hi.cc.u32 t119, t96, 0x0, t119;\n\t"
        "madc.lo.cc.u32 t120, t98, 0x0, t120;\n\t"
        "madc.hi.cc.u32 t121, t98, 0x0, t121;\n\t"
        "madc.lo.cc.u32 t122, t100, 0x0, t122;\n\t"
        "madc.hi.cc.u32 t123, t100, 0x0, t123;\n\t"
        "addc.u32 t124, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t133, t95, 0x0, t133;\n\t"
        "madc.hi.cc.u32 t134, t95, 0x0, t134;\n\t"
        "madc.lo.cc.u32 t135, t97, 0x0, t135;\n\t"
        "madc.hi.cc.u32 t136, t97, 0x0, t136;\n\t"
        "madc.lo.cc.u32 t137, t99, 0x0, t137;\n\t"
        "madc.hi.cc.u32 t138, t99, 0x0, t138;\n\t"
        "madc.lo.cc.u32 t139, t109, 0x0, t139;\n\t"
        "madc.hi.u32 t140, t109, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t118, t95, 0xffffffff, t118;\n\t"
        "madc.hi.cc.u32 t119, t95, 0xffffffff, t119;\n\t"
        "madc.lo.cc.u32 t120, t97, 0xffffffff, t120;\n\t"
        "madc.hi.cc.u32 t121, t97, 0xffffffff, t121;\n\t"
        "madc.lo.cc.u32 t122, t99, 0xffffffff, t122;\n\t"
        "madc.hi.cc.u32 t123, t99, 0xffffffff, t123;\n\t"
        "madc.lo.cc.u32 t124, t109, 0xffffffff, t124;\n\t"
        "madc.hi.u32 t125, t109, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t133, t77, 0xffffffff, t133;\n\t"
        "madc.hi.cc.u32 t134, t77, 0xffffffff, t134;\n\t"
        "madc.lo.cc.u32 t135, t96, 0xffffffff, t135;\n\t"
        "madc.hi.cc.u32 t136, t96, 0xffffffff, t136;\n\t"
        "madc.lo.cc.u32 t137, t98, 0xffffffff, t137;\n\t"
        "madc.hi.cc.u32 t138, t98, 0xffffffff, t138;\n\t"
        "madc.lo.cc.u32 t139, t100, 0xffffffff, t139;\n\t"
        "madc.hi.cc.u32 t140, t100, 0xffffffff, t140;\n\t"
        "addc.u32 t141, 0x0, 0x0;\n\t"
        "add.cc.u32 t142, t111, t127;\n\t"
        "addc.cc.u32 t143, t112, t128;\n\t"
        "addc.cc.u32 t144, t113, t129;\n\t"
        "addc.cc.u32 t145, t114, t130;\n\t"
        "addc.cc.u32 t146, t115, t131;\n\t"
        "addc.cc.u32 t147, t116, t132;\n\t"
        "addc.cc.u32 t148, t117, t133;\n\t"
        "addc.cc.u32 t149, t118, t134;\n\t"
        "addc.cc.u32 t150, t119, t135;\n\t"
        "addc.cc.u32 t151, t120, t136;\n\t"
        "addc.cc.u32 t152, t121, t137;\n\t"
        "addc.cc.u32 t153, t122, t138;\n\t"
        "addc.cc.u32 t154, t123, t139;\n\t"
        "addc.cc.u32 t155, t124, t140;\n\t"
        "addc.u32 t156, t125, t141;\n\t"
        "add.cc.u32 t157, t61, t110;\n\t"
        "addc.cc.u32 t158, t62, t142;\n\t"
        "addc.cc.u32 t159, t63, t143;\n\t"
        "addc.cc.u32 t160, t64, t144;\n\t"
        "addc.cc.u32 t161, t65, t145;\n\t"
        "addc.cc.u32 t162, t66, t146;\n\t"
        "addc.cc.u32 t163, t67, t147;\n\t"
        "addc.cc.u32 t164, t68, t148;\n\t"
        "addc.cc.u32 t165, t69, t149;\n\t"
        "addc.cc.u32 t166, t70, t150;\n\t"
        "addc.cc.u32 t167, t71, t151;\n\t"
        "addc.cc.u32 t168, t72, t152;\n\t"
        "addc.cc.u32 t169, t73, t153;\n\t"
        "addc.cc.u32 t170, t74, t154;\n\t"
        "addc.cc.u32 t171, t75, t155;\n\t"
        "addc.cc.u32 t172, t76, t156;\n\t"
        "addc.u32 t173, 0x0, 0x0;\n\t"
        "sub.cc.u32 t174, t165, 0xfc632551;\n\t"
        "subc.cc.u32 t175, t166, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t176, t167, 0xa7179e84;\n\t"
        "subc.cc.u32 t177, t168, 0xbce6faad;\n\t"
        "subc.cc.u32 t178, t169, 0xffffffff;\n\t"
        "subc.cc.u32 t179, t170, 0xffffffff;\n\t"
        "subc.cc.u32 t180, t171, 0x0;\n\t"
        "subc.cc.u32 t181, t172, 0xffffffff;\n\t"
        "subc.cc.u32 t182, t173, 0x0;\n\t"
        "subc.u32 t183, 0x0, 0x0;\n\t"
        "xor.b32 t184, t174, t165;\n\t"
        "and.b32 t185, t184, t183;\n\t"
        "xor.b32 t186, t185, t174;\n\t"
        "xor.b32 t187, t175, t166;\n\t"
        "and.b32 t188, t187, t183;\n\t"
        "xor.b32 t189, t188, t175;\n\t"
        "xor.b32 t190, t176, t167;\n\t"
        "and.b32 t191, t190, t183;\n\t"
        "xor.b32 t192, t191, t176;\n\t"
        "xor.b32 t193, t177, t168;\n\t"
        "and.b32 t194, t193, t183;\n\t"
        "xor.b32 t195, t194, t177;\n\t"
        "xor.b32 t196, t178, t169;\n\t"
        "and.b32 t197, t196, t183;\n\t"
        "xor.b32 t198, t197, t178;\n\t"
        "xor.b32 t199, t179, t170;\n\t"
        "and.b32 t200, t199, t183;\n\t"
        "xor.b32 t201, t200, t179;\n\t"
        "xor.b32 t202, t180, t171;\n\t"
        "and.b32 t203, t202, t183;\n\t"
        "xor.b32 t204, t203, t180;\n\t"
        "xor.b32 t205, t181, t172;\n\t"
        "and.b32 t206, t205, t183;\n\t"
        "xor.b32 t207, t206, t181;\n\t"
        "mov.u32 %0, t186;\n\t"
        "mov.u32 %1, t189;\n\t"
        "mov.u32 %2, t192;\n\t"
        "mov.u32 %3, t195;\n\t"
        "mov.u32 %4, t198;\n\t"
        "mov.u32 %5, t201;\n\t"
        "mov.u32 %6, t204;\n\t"
        "mov.u32 %7, t207;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127, t128, t129, t130, t131, t132, t133, t134, t135, t136, t137, t138, t139, t140, t141, t142, t143, t144, t145, t146, t147, t148, t149, t150, t151, t152, t153, t154, t155, t156, t157, t158, t159, t160, t161, t162, t163, t164, t165, t166, t167, t168, t169, t170, t171, t172, t173, t174, t175, t176, t177, t178, t179, t180, t181, t182, t183, t184, t185, t186, t187, t188, t189, t190, t191, t192, t193, t194, t195, t196, t197, t198, t199, t200, t201, t202, t203, t204, t205, t206, t207;
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
    t77 = (uint32_t)((uint32_t)(t61 * 0xee00bc4fu));
    t78 = (uint32_t)(((uint64_t)t61 * 0xee00bc4fu) >> 32);
    t79 = (uint32_t)((uint32_t)(t63 * 0xee00bc4fu));
    t80 = (uint32_t)(((uint64_t)t63 * 0xee00bc4fu) >> 32);
    t81 = (uint32_t)((uint32_t)(t65 * 0xee00bc4fu));
    t82 = (uint32_t)(((uint64_t)t65 * 0xee00bc4fu) >> 32);
    t83 = (uint32_t)((uint32_t)(t67 * 0xee00bc4fu));
    t84 = (uint32_t)(((uint64_t)t67 * 0xee00bc4fu) >> 32);
    t87 = (uint32_t)((uint32_t)(t62 * 0xee00bc4fu));
    t88 = (uint32_t)(((uint64_t)t62 * 0xee00bc4fu) >> 32);
    t89 = (uint32_t)((uint32_t)(t64 * 0xee00bc4fu));
    t90 = (uint32_t)(((uint64_t)t64 * 0xee00bc4fu) >> 32);
    t91 = (uint32_t)((uint32_t)(t66 * 0xee00bc4fu));
    t92 = (uint32_t)(((uint64_t)t66 * 0xee00bc4fu) >> 32);
    w_ = (uint64_t)(uint32_t)(t62 * 0xccd1c8aau) + t79; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t62 * 0xccd1c8aau) >> 32) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t64 * 0xccd1c8aau) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t64 * 0xccd1c8aau) >> 32) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t66 * 0xccd1c8aau) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t66 * 0xccd1c8aau) >> 32) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t85 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t61 * 0xccd1c8aau) + t87; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t61 * 0xccd1c8aau) >> 32) + t88 + cf_; t88 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t63 * 0xccd1c8aau) + t89 + cf_; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t63 * 0xccd1c8aau) >> 32) + t90 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t65 * 0xccd1c8aau) + t91 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t65 * 0xccd1c8aau) >> 32) + t92 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t93 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t61 * 0x7d74d2e4u) + t79; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t61 * 0x7d74d2e4u) >> 32) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t63 * 0x7d74d2e4u) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t63 * 0x7d74d2e4u) >> 32) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t65 * 0x7d74d2e4u) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t65 * 0x7d74d2e4u) >> 32) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t85 + 0x0u + cf_; t85 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t62 * 0x7d74d2e4u) + t89; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t62 * 0x7d74d2e4u) >> 32) + t90 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t64 * 0x7d74d2e4u) + t91 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t64 * 0x7d74d2e4u) >> 32) + t92 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t93 + 0x0u + cf_; t93 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t62 * 0x48c94408u) + t81; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t62 * 0x48c94408u) >> 32) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t64 * 0x48c94408u) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t64 * 0x48c94408u) >> 32) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t85 + 0x0u + cf_; t85 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t61 * 0x48c94408u) + t89; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t61 * 0x48c94408u) >> 32) + t90 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t63 * 0x48c94408u) + t91 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t63 * 0x48c94408u) >> 32) + t92 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t93 + 0x0u + cf_; t93 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t61 * 0xc588c6f6u) + t81; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t61 * 0xc588c6f6u) >> 32) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t63 * 0xc588c6f6u) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t63 * 0xc588c6f6u) >> 32) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t85 + 0x0u + cf_; t85 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t62 * 0xc588c6f6u) + t91; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t62 * 0xc588c6f6u) >> 32) + t92 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t93 + 0x0u + cf_; t93 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t62 * 0x50fe77ecu) + t83; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t62 * 0x50fe77ecu) >> 32) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t85 + 0x0u + cf_; t85 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t61 * 0x50fe77ecu) + t91; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t61 * 0x50fe77ecu) >> 32) + t92 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t93 + 0x0u + cf_; t93 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t61 * 0xa9d6281cu) + t83; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t61 * 0xa9d6281cu) >> 32) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t85 + 0x0u + cf_; t85 = (uint32_t)w_;
    w_ = (uint64_t)t78 + t87; t95 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t79 + t88 + cf_; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t80 + t89 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t81 + t90 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t82 + t91 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t83 + t92 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t84 + t93 + cf_; t101 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t68 * 0xee00bc4fu) + t101; t102 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t67 * 0xccd1c8aau) + t102; t103 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t66 * 0x7d74d2e4u) + t103; t104 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t65 * 0x48c94408u) + t104; t105 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t64 * 0xc588c6f6u) + t105; t106 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t63 * 0x50fe77ecu) + t106; t107 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t62 * 0xa9d6281cu) + t107; t108 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t61 * 0x60d06633u) + t108; t109 = (uint32_t)w_;
    t110 = (uint32_t)((uint32_t)(t77 * 0xfc632551u));
    t111 = (uint32_t)(((uint64_t)t77 * 0xfc632551u) >> 32);
    t112 = (uint32_t)((uint32_t)(t96 * 0xfc632551u));
    t113 = (uint32_t)(((uint64_t)t96 * 0xfc632551u) >> 32);
    t114 = (uint32_t)((uint32_t)(t98 * 0xfc632551u));
    t115 = (uint32_t)(((uint64_t)t98 * 0xfc632551u) >> 32);
    t116 = (uint32_t)((uint32_t)(t100 * 0xfc632551u));
    t117 = (uint32_t)(((uint64_t)t100 * 0xfc632551u) >> 32);
    t127 = (uint32_t)((uint32_t)(t95 * 0xfc632551u));
    t128 = (uint32_t)(((uint64_t)t95 * 0xfc632551u) >> 32);
    t129 = (uint32_t)((uint32_t)(t97 * 0xfc632551u));
    t130 = (uint32_t)(((uint64_t)t97 * 0xfc632551u) >> 32);
    t131 = (uint32_t)((uint32_t)(t99 * 0xfc632551u));
    t132 = (uint32_t)(((uint64_t)t99 * 0xfc632551u) >> 32);
    t133 = (uint32_t)((uint32_t)(t109 * 0xfc632551u));
    t134 = (uint32_t)(((uint64_t)t109 * 0xfc632551u) >> 32);
    w_ = (uint64_t)(uint32_t)(t95 * 0xf3b9cac2u) + t112; t112 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t95 * 0xf3b9cac2u) >> 32) + t113 + cf_; t113 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t97 * 0xf3b9cac2u) + t114 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t97 * 0xf3b9cac2u) >> 32) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t99 * 0xf3b9cac2u) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t99 * 0xf3b9cac2u) >> 32) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t109 * 0xf3b9cac2u) + 0x0u + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t109 * 0xf3b9cac2u) >> 32) + 0x0u + cf_; t119 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t77 * 0xf3b9cac2u) + t127; t127 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t77 * 0xf3b9cac2u) >> 32) + t128 + cf_; t128 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t96 * 0xf3b9cac2u) + t129 + cf_; t129 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t96 * 0xf3b9cac2u) >> 32) + t130 + cf_; t130 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t98 * 0xf3b9cac2u) + t131 + cf_; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t98 * 0xf3b9cac2u) >> 32) + t132 + cf_; t132 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t100 * 0xf3b9cac2u) + t133 + cf_; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t100 * 0xf3b9cac2u) >> 32) + t134 + cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t135 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t77 * 0xa7179e84u) + t112; t112 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t77 * 0xa7179e84u) >> 32) + t113 + cf_; t113 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t96 * 0xa7179e84u) + t114 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t96 * 0xa7179e84u) >> 32) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t98 * 0xa7179e84u) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t98 * 0xa7179e84u) >> 32) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t100 * 0xa7179e84u) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t100 * 0xa7179e84u) >> 32) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t120 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t95 * 0xa7179e84u) + t129; t129 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t95 * 0xa7179e84u) >> 32) + t130 + cf_; t130 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t97 * 0xa7179e84u) + t131 + cf_; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t97 * 0xa7179e84u) >> 32) + t132 + cf_; t132 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t99 * 0xa7179e84u) + t133 + cf_; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t99 * 0xa7179e84u) >> 32) + t134 + cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t109 * 0xa7179e84u) + t135 + cf_; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t109 * 0xa7179e84u) >> 32) + 0x0u + cf_; t136 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t95 * 0xbce6faadu) + t114; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t95 * 0xbce6faadu) >> 32) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t97 * 0xbce6faadu) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t97 * 0xbce6faadu) >> 32) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t99 * 0xbce6faadu) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t99 * 0xbce6faadu) >> 32) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t109 * 0xbce6faadu) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t109 * 0xbce6faadu) >> 32) + 0x0u + cf_; t121 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t77 * 0xbce6faadu) + t129; t129 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t77 * 0xbce6faadu) >> 32) + t130 + cf_; t130 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t96 * 0xbce6faadu) + t131 + cf_; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t96 * 0xbce6faadu) >> 32) + t132 + cf_; t132 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t98 * 0xbce6faadu) + t133 + cf_; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t98 * 0xbce6faadu) >> 32) + t134 + cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t100 * 0xbce6faadu) + t135 + cf_; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t100 * 0xbce6faadu) >> 32) + t136 + cf_; t136 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t137 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t77 * 0xffffffffu) + t114; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t77 * 0xffffffffu) >> 32) + t115 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t96 * 0xffffffffu) + t116 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t96 * 0xffffffffu) >> 32) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t98 * 0xffffffffu) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t98 * 0xffffffffu) >> 32) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t100 * 0xffffffffu) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t100 * 0xffffffffu) >> 32) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t122 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t95 * 0xffffffffu) + t131; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t95 * 0xffffffffu) >> 32) + t132 + cf_; t132 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t97 * 0xffffffffu) + t133 + cf_; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t97 * 0xffffffffu) >> 32) + t134 + cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t99 * 0xffffffffu) + t135 + cf_; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t99 * 0xffffffffu) >> 32) + t136 + cf_; t136 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t109 * 0xffffffffu) + t137 + cf_; t137 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t109 * 0xffffffffu) >> 32) + 0x0u + cf_; t138 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t95 * 0xffffffffu) + t116; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t95 * 0xffffffffu) >> 32) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t97 * 0xffffffffu) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t97 * 0xffffffffu) >> 32) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t99 * 0xffffffffu) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t99 * 0xffffffffu) >> 32) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t109 * 0xffffffffu) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t109 * 0xffffffffu) >> 32) + 0x0u + cf_; t123 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t77 * 0xffffffffu) + t131; t131 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t77 * 0xffffffffu) >> 32) + t132 + cf_; t132 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t96 * 0xffffffffu) + t133 + cf_; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t96 * 0xffffffffu) >> 32) + t134 + cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t98 * 0xffffffffu) + t135 + cf_; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t98 * 0xffffffffu) >> 32) + t136 + cf_; t136 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t100 * 0xffffffffu) + t137 + cf_; t137 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t100 * 0xffffffffu) >> 32) + t138 + cf_; t138 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t139 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t77 * 0x0u) + t116; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t77 * 0x0u) >> 32) + t117 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t96 * 0x0u) + t118 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t96 * 0x0u) >> 32) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t98 * 0x0u) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t98 * 0x0u) >> 32) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t100 * 0x0u) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t100 * 0x0u) >> 32) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t124 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t95 * 0x0u) + t133; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t95 * 0x0u) >> 32) + t134 + cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t97 * 0x0u) + t135 + cf_; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t97 * 0x0u) >> 32) + t136 + cf_; t136 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t99 * 0x0u) + t137 + cf_; t137 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t99 * 0x0u) >> 32) + t138 + cf_; t138 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t109 * 0x0u) + t139 + cf_; t139 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t109 * 0x0u) >> 32) + 0x0u + cf_; t140 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t95 * 0xffffffffu) + t118; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t95 * 0xffffffffu) >> 32) + t119 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t97 * 0xffffffffu) + t120 + cf_; t120 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t97 * 0xffffffffu) >> 32) + t121 + cf_; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t99 * 0xffffffffu) + t122 + cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t99 * 0xffffffffu) >> 32) + t123 + cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t109 * 0xffffffffu) + t124 + cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t109 * 0xffffffffu) >> 32) + 0x0u + cf_; t125 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t77 * 0xffffffffu) + t133; t133 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t77 * 0xffffffffu) >> 32) + t134 + cf_; t134 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t96 * 0xffffffffu) + t135 + cf_; t135 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t96 * 0xffffffffu) >> 32) + t136 + cf_; t136 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t98 * 0xffffffffu) + t137 + cf_; t137 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t98 * 0xffffffffu) >> 32) + t138 + cf_; t138 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t100 * 0xffffffffu) + t139 + cf_; t139 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t100 * 0xffffffffu) >> 32) + t140 + cf_; t140 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t141 = (uint32_t)w_;
    w_ = (uint64_t)t111 + t127; t142 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t112 + t128 + cf_; t143 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t113 + t129 + cf_; t144 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t114 + t130 + cf_; t145 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t115 + t131 + cf_; t146 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t116 + t132 + cf_; t147 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t117 + t133 + cf_; t148 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t118 + t134 + cf_; t149 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t119 + t135 + cf_; t150 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t120 + t136 + cf_; t151 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t121 + t137 + cf_; t152 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t122 + t138 + cf_; t153 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t123 + t139 + cf_; t154 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t124 + t140 + cf_; t155 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t125 + t141 + cf_; t156 = (uint32_t)w_;
    w_ = (uint64_t)t61 + t110; t157 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t62 + t142 + cf_; t158 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + t143 + cf_; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t64 + t144 + cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t65 + t145 + cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t66 + t146 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t67 + t147 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t68 + t148 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t69 + t149 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t70 + t150 + cf_; t166 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t71 + t151 + cf_; t167 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t72 + t152 + cf_; t168 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t73 + t153 + cf_; t169 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t74 + t154 + cf_; t170 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t75 + t155 + cf_; t171 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t76 + t156 + cf_; t172 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t173 = (uint32_t)w_;
    w_ = (uint64_t)t165 - 0xfc632551u; t174 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t166 - 0xf3b9cac2u - cf_; t175 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t167 - 0xa7179e84u - cf_; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t168 - 0xbce6faadu - cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t169 - 0xffffffffu - cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t170 - 0xffffffffu - cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t171 - 0x0u - cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t172 - 0xffffffffu - cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t173 - 0x0u - cf_; t182 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t183 = (uint32_t)w_;
    t184 = (uint32_t)(t174 ^ t165);
    t185 = (uint32_t)(t184 & t183);
    t186 = (uint32_t)(t185 ^ t174);
    t187 = (uint32_t)(t175 ^ t166);
    t188 = (uint32_t)(t187 & t183);
    t189 = (uint32_t)(t188 ^ t175);
    t190 = (uint32_t)(t176 ^ t167);
    t191 = (uint32_t)(t190 & t183);
    t192 = (uint32_t)(t191 ^ t176);
    t193 = (uint32_t)(t177 ^ t168);
    t194 = (uint32_t)(t193 & t183);
    t195 = (uint32_t)(t194 ^ t177);
    t196 = (uint32_t)(t178 ^ t169);
    t197 = (uint32_t)(t196 & t183);
    t198 = (uint32_t)(t197 ^ t178);
    t199 = (uint32_t)(t179 ^ t170);
    t200 = (uint32_t)(t199 & t183);
    t201 = (uint32_t)(t200 ^ t179);
    t202 = (uint32_t)(t180 ^ t171);
    t203 = (uint32_t)(t202 & t183);
    t204 = (uint32_t)(t203 ^ t180);
    t205 = (uint32_t)(t181 ^ t172);
    t206 = (uint32_t)(t205 & t183);
    t207 = (uint32_t)(t206 ^ t181);
    r[0] = t186;
    r[1] = t189;
    r[2] = t192;
    r[3] = t195;
    r[4] = t198;
    r[5] = t201;
    r[6] = t204;
    r[7] = t207;
#endif
  }

  // c = a*b for a small integer b (pseudo.py:705-728 / monty.py:876-978)
  static MAB_DEV void mli(uint32_t (&r)[8], const uint32_t (&a)[8], uint32_t b) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<333>;\n\t"
        "mul.lo.u32 t0, 0xbe79eea2, %16;\n\t"
        "mul.hi.u32 t8, 0xbe79eea2, %16;\n\t"
        "mul.lo.u32 t1, 0x83244c95, %16;\n\t"
        "mul.hi.u32 t9, 0x83244c95, %16;\n\t"
        "mul.lo.u32 t2, 0x49bd6fa6, %16;\n\t"
        "mul.hi.u32 t10, 0x49bd6fa6, %16;\n\t"
        "mul.lo.u32 t3, 0x4699799c, %16;\n\t"
        "mul.hi.u32 t11, 0x4699799c, %16;\n\t"
        "mul.lo.u32 t4, 0x2b6bec59, %16;\n\t"
        "mul.hi.u32 t12, 0x2b6bec59, %16;\n\t"
        "mul.lo.u32 t5, 0x2845b239, %16;\n\t"
        "mul.hi.u32 t13, 0x2845b239, %16;\n\t"
        "mul.lo.u32 t6, 0xf3d95620, %16;\n\t"
        "mul.hi.u32 t14, 0xf3d95620, %16;\n\t"
        "mul.lo.u32 t7, 0x66e12d94, %16;\n\t"
        "mul.hi.u32 t15, 0x66e12d94, %16;\n\t"
        "add.cc.u32 t16, t1, t8;\n\t"
        "addc.cc.u32 t17, t2, t9;\n\t"
        "addc.cc.u32 t18, t3, t10;\n\t"
        "addc.cc.u32 t19, t4, t11;\n\t"
        "addc.cc.u32 t20, t5, t12;\n\t"
        "addc.cc.u32 t21, t6, t13;\n\t"
        "addc.cc.u32 t22, t7, t14;\n\t"
        "addc.u32 t23, t15, 0x0;\n\t"
        "mul.lo.u32 t24, t0, 0xee00bc4f;\n\t"
        "mul.hi.u32 t25, t0, 0xee00bc4f;\n\t"
        "mul.lo.u32 t26, t17, 0xee00bc4f;\n\t"
        "mul.hi.u32 t27, t17, 0xee00bc4f;\n\t"
        "mul.lo.u32 t28, t19, 0xee00bc4f;\n\t"
        "mul.hi.u32 t29, t19, 0xee00bc4f;\n\t"
        "mul.lo.u32 t30, t21, 0xee00bc4f;\n\t"
        "mul.hi.u32 t31, t21, 0xee00bc4f;\n\t"
        "mul.lo.u32 t34, t16, 0xee00bc4f;\n\t"
        "mul.hi.u32 t35, t16, 0xee00bc4f;\n\t"
        "mul.lo.u32 t36, t18, 0xee00bc4f;\n\t"
        "mul.hi.u32 t37, t18, 0xee00bc4f;\n\t"
        "mul.lo.u32 t38, t20, 0xee00bc4f;\n\t"
        "mul.hi.u32 t39, t20, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t26, t16, 0xccd1c8aa, t26;\n\t"
        "madc.hi.cc.u32 t27, t16, 0xccd1c8aa, t27;\n\t"
        "madc.lo.cc.u32 t28, t18, 0xccd1c8aa, t28;\n\t"
        "madc.hi.cc.u32 t29, t18, 0xccd1c8aa, t29;\n\t"
        "madc.lo.cc.u32 t30, t20, 0xccd1c8aa, t30;\n\t"
        "madc.hi.cc.u32 t31, t20, 0xccd1c8aa, t31;\n\t"
        "addc.u32 t32, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t34, t0, 0xccd1c8aa, t34;\n\t"
        "madc.hi.cc.u32 t35, t0, 0xccd1c8aa, t35;\n\t"
        "madc.lo.cc.u32 t36, t17, 0xccd1c8aa, t36;\n\t"
        "madc.hi.cc.u32 t37, t17, 0xccd1c8aa, t37;\n\t"
        "madc.lo.cc.u32 t38, t19, 0xccd1c8aa, t38;\n\t"
        "madc.hi.cc.u32 t39, t19, 0xccd1c8aa, t39;\n\t"
        "addc.u32 t40, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t26, t0, 0x7d74d2e4, t26;\n\t"
        "madc.hi.cc.u32 t27, t0, 0x7d74d2e4, t27;\n\t"
        "madc.lo.cc.u32 t28, t17, 0x7d74d2e4, t28;\n\t"
        "madc.hi.cc.u32 t29, t17, 0x7d74d2e4, t29;\n\t"
        "madc.lo.cc.u32 t30, t19, 0x7d74d2e4, t30;\n\t"
        "madc.hi.cc.u32 t31, t19, 0x7d74d2e4, t31;\n\t"
        "addc.u32 t32, t32, 0x0;\n\t"
        "mad.lo.cc.u32 t36, t16, 0x7d74d2e4, t36;\n\t"
        "madc.hi.cc.u32 t37, t16, 0x7d74d2e4, t37;\n\t"
        "madc.lo.cc.u32 t38, t18, 0x7d74d2e4, t38;\n\t"
        "madc.hi.cc.u32 t39, t18, 0x7d74d2e4, t39;\n\t"
        "addc.u32 t40, t40, 0x0;\n\t"
        "mad.lo.cc.u32 t28, t16, 0x48c94408, t28;\n\t"
        "madc.hi.cc.u32 t29, t16, 0x48c94408, t29;\n\t"
        "madc.lo.cc.u32 t30, t18, 0x48c94408, t30;\n\t"
        "madc.hi.cc.u32 t31, t18, 0x48c94408, t31;\n\t"
        "addc.u32 t32, t32, 0x0;\n\t"
        "mad.lo.cc.u32 t36, t0, 0x48c94408, t36;\n\t"
        "madc.hi.cc.u32 t37, t0, 0x48c94408, t37;\n\t"
        "madc.lo.cc.u32 t38, t17, 0x48c94408, t38;\n\t"
        "madc.hi.cc.u32 t39, t17, 0x48c94408, t39;\n\t"
        "addc.u32 t40, t40, 0x0;\n\t"
        "mad.lo.cc.u32 t28, t0, 0xc588c6f6, t28;\n\t"
        "madc.hi.cc.u32 t29, t0, 0xc588c6f6, t29;\n\t"
        "madc.lo.cc.u32 t30, t17, 0xc588c6f6, t30;\n\t"
        "madc.hi.cc.u32 t31, t17, 0xc588c6f6, t31;\n\t"
        "addc.u32 t32, t32, 0x0;\n\t"
        "mad.lo.cc.u32 t38, t16, 0xc588c6f6, t38;\n\t"
        "madc.hi.cc.u32 t39, t16, 0xc588c6f6, t39;\n\t"
        "addc.u32 t40, t40, 0x0;\n\t"
        "mad.lo.cc.u32 t30, t16, 0x50fe77ec, t30;\n\t"
        "madc.hi.cc.u32 t31, t16, 0x50fe77ec, t31;\n\t"
        "addc.u32 t32, t32, 0x0;\n\t"
        "mad.lo.cc.u32 t38, t0, 0x50fe77ec, t38;\n\t"
        "madc.hi.cc.u32 t39, t0, 0x50fe77ec, t39;\n\t"
        "addc.u32 t40, t40, 0x0;\n\t"
        "mad.lo.cc.u32 t30, t0, 0xa9d6281c, t30;\n\t"
        "madc.hi.cc.u32 t31, t0, 0xa9d6281c, t31;\n\t"
        "addc.u32 t32, t32, 0x0;\n\t"
        "add.cc.u32 t42, t25, t34;\n\t"
        "addc.cc.u32 t43, t26, t35;\n\t"
        "addc.cc.u32 t44, t27, t36;\n\t"
        "addc.cc.u32 t45, t28, t37;\n\t"
        "addc.cc.u32 t46, t29, t38;\n\t"
        "addc.cc.u32 t47, t30, t39;\n\t"
        "addc.u32 t48, t31, t40;\n\t"
        "mad.lo.u32 t49, t22, 0xee00bc4f, t48;\n\t"
        "mad.lo.u32 t50, t21, 0xccd1c8aa, t49;\n\t"
        "mad.lo.u32 t51, t20, 0x7d74d2e4, t50;\n\t"
        "mad.lo.u32 t52, t19, 0x48c94408, t51;\n\t"
        "mad.lo.u32 t53, t18, 0xc588c6f6, t52;\n\t"
        "mad.lo.u32 t54, t17, 0x50fe77ec, t53;\n\t"
        "mad.lo.u32 t55, t16, 0xa9d6281c, t54;\n\t"
        "mad.lo.u32 t56, t0, 0x60d06633, t55;\n\t"
        "mul.lo.u32 t57, t24, 0xfc632551;\n\t"
        "mul.hi.u32 t58, t24, 0xfc632551;\n\t"
        "mul.lo.u32 t59, t43, 0xfc632551;\n\t"
        "mul.hi.u32 t60, t43, 0xfc632551;\n\t"
        "mul.lo.u32 t61, t45, 0xfc632551;\n\t"
        "mul.hi.u32 t62, t45, 0xfc632551;\n\t"
        "mul.lo.u32 t63, t47, 0xfc632551;\n\t"
        "mul.hi.u32 t64, t47, 0xfc632551;\n\t"
        "mul.lo.u32 t74, t42, 0xfc632551;\n\t"
        "mul.hi.u32 t75, t42, 0xfc632551;\n\t"
        "mul.lo.u32 t76, t44, 0xfc632551;\n\t"
        "mul.hi.u32 t77, t44, 0xfc632551;\n\t"
        "mul.lo.u32 t78, t46, 0xfc632551;\n\t"
        "mul.hi.u32 t79, t46, 0xfc632551;\n\t"
        "mul.lo.u32 t80, t56, 0xfc632551;\n\t"
        "mul.hi.u32 t81, t56, 0xfc632551;\n\t"
        "mad.lo.cc.u32 t59, t42, 0xf3b9cac2, t59;\n\t"
        "madc.hi.cc.u32 t60, t42, 0xf3b9cac2, t60;\n\t"
        "madc.lo.cc.u32 t61, t44, 0xf3b9cac2, t61;\n\t"
        "madc.hi.cc.u32 t62, t44, 0xf3b9cac2, t62;\n\t"
        "madc.lo.cc.u32 t63, t46, 0xf3b9cac2, t63;\n\t"
        "madc.hi.cc.u32 t64, t46, 0xf3b9cac2, t64;\n\t"
        "madc.lo.cc.u32 t65, t56, 0xf3b9cac2, 0x0;\n\t"
        "madc.hi.u32 t66, t56, 0xf3b9cac2, 0x0;\n\t"
        "mad.lo.cc.u32 t74, t24, 0xf3b9cac2, t74;\n\t"
        "madc.hi.cc.u32 t75, t24, 0xf3b9cac2, t75;\n\t"
        "madc.lo.cc.u32 t76, t43, 0xf3b9cac2, t76;\n\t"
        "madc.hi.cc.u32 t77, t43, 0xf3b9cac2, t77;\n\t"
        "madc.lo.cc.u32 t78, t45, 0xf3b9cac2, t78;\n\t"
        "madc.hi.cc.u32 t79, t45, 0xf3b9cac2, t79;\n\t"
        "madc.lo.cc.u32 t80, t47, 0xf3b9cac2, t80;\n\t"
        "madc.hi.cc.u32 t81, t47, 0xf3b9cac2, t81;\n\t"
        "addc.u32 t82, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t59, t24, 0xa7179e84, t59;\n\t"
        "madc.hi.cc.u32 t60, t24, 0xa7179e84, t60;\n\t"
        "madc.lo.cc.u32 t61, t43, 0xa7179e84, t61;\n\t"
        "madc.hi.cc.u32 t62, t43, 0xa7179e84, t62;\n\t"
        "madc.lo.cc.u32 t63, t45, 0xa7179e84, t63;\n\t"
        "madc.hi.cc.u32 t64, t45, 0xa7179e84, t64;\n\t"
        "madc.lo.cc.u32 t65, t47, 0xa7179e84, t65;\n\t"
        "madc.hi.cc.u32 t66, t47, 0xa7179e84, t66;\n\t"
        "addc.u32 t67, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t76, t42, 0xa7179e84, t76;\n\t"
        "madc.hi.cc.u32 t77, t42, 0xa7179e84, t77;\n\t"
        "madc.lo.cc.u32 t78, t44, 0xa7179e84, t78;\n\t"
        "madc.hi.cc.u32 t79, t44, 0xa7179e84, t79;\n\t"
        "madc.lo.cc.u32 t80, t46, 0xa7179e84, t80;\n\t"
        "madc.hi.cc.u32 t81, t46, 0xa7179e84, t81;\n\t"
        "madc.lo.cc.u32 t82, t56, 0xa7179e84, t82;\n\t"
        "madc.hi.u32 t83, t56, 0xa7179e84, 0x0;\n\t"
        "mad.lo.cc.u32 t61, t42, 0xbce6faad, t61;\n\t"
        "madc.hi.cc.u32 t62, t42, 0xbce6faad, t62;\n\t"
        "madc.lo.cc.u32 t63, t44, 0xbce6faad, t63;\n\t"
        "madc.hi.cc.u32 t64, t44, 0xbce6faad, t64;\n\t"
        "madc.lo.cc.u32 t65, t46, 0xbce6faad, t65;\n\t"
        "madc.hi.cc.u32 t66, t46, 0xbce6faad, t66;\n\t"
        "madc.lo.cc.u32 t67, t56, 0xbce6faad, t67;\n\t"
        "madc.hi.u32 t68, t56, 0xbce6faad, 0x0;\n\t"
        "mad.lo.cc.u32 t76, t24, 0xbce6faad, t76;\n\t"
        "madc.hi.cc.u32 t77, t24, 0xbce6faad, t77;\n\t"
        "madc.lo.cc.u32 t78, t43, 0xbce6faad, t78;\n\t"
        "madc.hi.cc.u32 t79, t43, 0xbce6faad, t79;\n\t"
        "madc.lo.cc.u32 t80, t45, 0xbce6faad, t80;\n\t"
        "madc.hi.cc.u32 t81, t45, 0xbce6faad, t81;\n\t"
        "madc.lo.cc.u32 t82, t47, 0xbce6faad, t82;\n\t"
        "madc.hi.cc.u32 t83, t47, 0xbce6faad, t83;\n\t"
        "addc.u32 t84, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t61, t24, 0xffffffff, t61;\n\t"
        "madc.hi.cc.u32 t62, t24, 0xffffffff, t62;\n\t"
        "madc.lo.cc.u32 t63, t43, 0xffffffff, t63;\n\t"
        "madc.hi.cc.u32 t64, t43, 0xffffffff, t64;\n\t"
        "madc.lo.cc.u32 t65, t45, 0xffffffff, t65;\n\t"
        "madc.hi.cc.u32 t66, t45, 0xffffffff, t66;\n\t"
        "madc.lo.cc.u32 t67, t47, 0xffffffff, t67;\n\t"
        "madc.hi.cc.u32 t68, t47, 0xffffffff, t68;\n\t"
        "addc.u32 t69, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t78, t42, 0xffffffff, t78;\n\t"
        "madc.hi.cc.u32 t79, t42, 0xffffffff, t79;\n\t"
        "madc.lo.cc.u32 t80, t44, 0xffffffff, t80;\n\t"
        "madc.hi.cc.u32 t81, t44, 0xffffffff, t81;\n\t"
        "madc.lo.cc.u32 t82, t46, 0xffffffff, t82;\n\t"
        "madc.hi.cc.u32 t83, t46, 0xffffffff, t83;\n\t"
        "madc.lo.cc.u32 t84, t56, 0xffffffff, t84;\n\t"
        "madc.hi.u32 t85, t56, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t63, t42, 0xffffffff, t63;\n\t"
        "madc.hi.cc.u32 t64, t42, 0xffffffff, t64;\n\t"
        "madc.lo.cc.u32 t65, t44, 0xffffffff, t65;\n\t"
        "madc.hi.cc.u32 t66, t44, 0xffffffff, t66;\n\t"
        "madc.lo.cc.u32 t67, t46, 0xffffffff, t67;\n\t"
        "madc.hi.cc.u32 t68, t46, 0xffffffff, t68;\n\t"
        "madc.lo.cc.u32 t69, t56, 0xffffffff, t69;\n\t"
        "madc.hi.u32 t70, t56, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t78, t24, 0xffffffff, t78;\n\t"
        "madc.hi.cc.u32 t79, t24, 0xffffffff, t79;\n\t"
        "madc.lo.cc.u32 t80, t43, 0xffffffff, t80;\n\t"
        "madc.hi.cc.u32 t81, t43, 0xffffffff, t81;\n\t"
        "madc.lo.cc.u32 t82, t45, 0xffffffff, t82;\n\t"
        "madc.hi.cc.u32 t83, t45, 0xffffffff, t83;\n\t"
        "madc.lo.cc.u32 t84, t47, 0xffffffff, t84;\n\t"
        "madc.hi.cc.u32 t85, t47, 0xffffffff, t85;\n\t"
        "addc.u32 t86, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t63, t24, 0x0, t63;\n\t"
        "madc.hi.cc.u32 t64, t24, 0x0, t64;\n\t"
        "madc.lo.cc.u32 t65, t43, 0x0, t65;\n\t"
        "madc.hi.cc.u32 t66, t43, 0x0, t66;\n\t"
        "madc.lo.cc.u32 t67, t45, 0x0, t67;\n\t"
        "madc.hi.cc.u32 t68, t45, 0x0, t68;\n\t"
        "madc.lo.cc.u32 t69, t47, 0x0, t69;\n\t"
        "madc.hi.cc.u32 t70, t47, 0x0, t70;\n\t"
        "addc.u32 t71, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t80, t42, 0x0, t80;\n\t"
        "madc.hi.cc.u32 t81, t42, 0x0, t81;\n\t"
        "madc.lo.cc.u32 t82, t44, 0x0, t82;\n\t"
        "madc.hi.cc.u32 t83, t44, 0x0, t83;\n\t"
        "madc.lo.cc.u32 t84, t46, 0x0, t84;\n\t"
        "madc.hi.cc.u32 t85, t46, 0x0, t85;\n\t"
        "madc.lo.cc.u32 t86, t56, 0x0, t86;\n\t"
        "madc.hi.u32 t87, t56, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t65, t42, 0xffffffff, t65;\n\t"
        "madc.hi.cc.u32 t66, t42, 0xffffffff, t66;\n\t"
        "madc.lo.cc.u32 t67, t44, 0xffffffff, t67;\n\t"
        "madc.hi.cc.u32 t68, t44, 0xffffffff, t68;\n\t"
        "madc.lo.cc.u32 t69, t46, 0xffffffff, t69;\n\t"
        "madc.hi.cc.u32 t70, t46, 0xffffffff, t70;\n\t"
        "madc.lo.cc.u32 t71, t56, 0xffffffff, t71;\n\t"
        "madc.hi.u32 t72, t56, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t80, t24, 0xffffffff, t80;\n\t"
        "madc.hi.cc.u32 t81, t24, 0xffffffff, t81;\n\t"
        "madc.lo.cc.u32 t82, t43, 0xffffffff, t82;\n\t"
        "madc.hi.cc.u32 t83, t43, 0xffffffff, t83;\n\t"
        "madc.lo.cc.u32 t84, t45, 0xffffffff, t84;\n\t"
        "madc.hi.cc.u32 t85, t45, 0xffffffff, t85;\n\t"
        "madc.lo.cc.u32 t86, t47, 0xffffffff, t86;\n\t"
        "madc.hi.cc.u32 t87, t47, 0xffffffff, t87;\n\t"
        "addc.u32 t88, 0x0, 0x0;\n\t"
        "add.cc.u32 t89, t58, t74;\n\t"
        "addc.cc.u32 t90, t59, t75;\n\t"
        "addc.cc.u32 t91, t60, t76;\n\t"
        "addc.cc.u32 t92, t61, t77;\n\t"
        "addc.cc.u32 t93, t62, t78;\n\t"
        "addc.cc.u32 t94, t63, t79;\n\t"
        "addc.cc.u32 t95, t64, t80;\n\t"
        "addc.cc.u32 t96, t65, t81;\n\t"
        "addc.cc.u32 t97, t66, t82;\n\t"
        "addc.cc.u32 t98, t67, t83;\n\t"
        "addc.cc.u32 t99, t68, t84;\n\t"
        "addc.cc.u32 t100, t69, t85;\n\t"
        "addc.cc.u32 t101, t70, t86;\n\t"
        "addc.cc.u32 t102, t71, t87;\n\t"
        "addc.u32 t103, t72, t88;\n\t"
        "add.cc.u32 t104, t0, t57;\n\t"
        "addc.cc.u32 t105, t16, t89;\n\t"
        "addc.cc.u32 t106, t17, t90;\n\t"
        "addc.cc.u32 t107, t18, t91;\n\t"
        "addc.cc.u32 t108, t19, t92;\n\t"
        "addc.cc.u32 t109, t20, t93;\n\t"
        "addc.cc.u32 t110, t21, t94;\n\t"
        "addc.cc.u32 t111, t22, t95;\n\t"
        "addc.cc.u32 t112, t23, t96;\n\t"
        "addc.cc.u32 t113, 0x0, t97;\n\t"
        "addc.cc.u32 t114, 0x0, t98;\n\t"
        "addc.cc.u32 t115, 0x0, t99;\n\t"
        "addc.cc.u32 t116, 0x0, t100;\n\t"
        "addc.cc.u32 t117, 0x0, t101;\n\t"
        "addc.cc.u32 t118, 0x0, t102;\n\t"
        "addc.cc.u32 t119, 0x0, t103;\n\t"
        "addc.u32 t120, 0x0, 0x0;\n\t"
        "sub.cc.u32 t121, t112, 0xfc632551;\n\t"
        "subc.cc.u32 t122, t113, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t123, t114, 0xa7179e84;\n\t"
        "subc.cc.u32 t124, t115, 0xbce6faad;\n\t"
        "subc.cc.u32 t125, t116, 0xffffffff;\n\t"
        "subc.cc.u32 t126, t117, 0xffffffff;\n\t"
        "subc.cc.u32 t127, t118, 0x0;\n\t"
        "subc.cc.u32 t128, t119, 0xffffffff;\n\t"
        "subc.cc.u32 t129, t120, 0x0;\n\t"
        "subc.u32 t130, 0x0, 0x0;\n\t"
        "xor.b32 t131, t121, t112;\n\t"
        "and.b32 t132, t131, t130;\n\t"
        "xor.b32 t133, t132, t121;\n\t"
        "xor.b32 t134, t122, t113;\n\t"
        "and.b32 t135, t134, t130;\n\t"
        "xor.b32 t136, t135, t122;\n\t"
        "xor.b32 t137, t123, t114;\n\t"
        "and.b32 t138, t137, t130;\n\t"
        "xor.b32 t139, t138, t123;\n\t"
        "xor.b32 t140, t124, t115;\n\t"
        "and.b32 t141, t140, t130;\n\t"
        "xor.b32 t142, t141, t124;\n\t"
        "xor.b32 t143, t125, t116;\n\t"
        "and.b32 t144, t143, t130;\n\t"
        "xor.b32 t145, t144, t125;\n\t"
        "xor.b32 t146, t126, t117;\n\t"
        "and.b32 t147, t146, t130;\n\t"
        "xor.b32 t148, t147, t126;\n\t"
        "xor.b32 t149, t127, t118;\n\t"
        "and.b32 t150, t149, t130;\n\t"
        "xor.b32 t151, t150, t127;\n\t"
        "xor.b32 t152, t128, t119;\n\t"
        "and.b32 t153, t152, t130;\n\t"
        "xor.b32 t154, t153, t128;\n\t"
        "mul.lo.u32 t155, %8, t133;\n\t"
        "mul.hi.u32 t156, %8, t133;\n\t"
        "mul.lo.u32 t157, %10, t133;\n\t"
        "mul.hi.u32 t158, %10, t133;\n\t"
        "mul.lo.u32 t159, %12, t133;\n\t"
        "mul.hi.u32 t160, %12, t133;\n\t"
        "mul.lo.u32 t161, %14, t133;\n\t"
        "mul.hi.u32 t162, %14, t133;\n\t"
        "mul.lo.u32 t172, %9, t133;\n\t"
        "mul.hi.u32 t173, %9, t133;\n\t"
        "mul.lo.u32 t174, %11, t133;\n\t"
        "mul.hi.u32 t175, %11, t133;\n\t"
        "mul.lo.u32 t176, %13, t133;\n\t"
        "mul.hi.u32 t177, %13, t133;\n\t"
        "mul.lo.u32 t178, %15, t133;\n\t"
        "mul.hi.u32 t179, %15, t133;\n\t"
        "mad.lo.cc.u32 t157, %9, t136, t157;\n\t"
        "madc.hi.cc.u32 t158, %9, t136, t158;\n\t"
        "madc.lo.cc.u32 t159, %11, t136, t159;\n\t"
        "madc.hi.cc.u32 t160, %11, t136, t160;\n\t"
        "madc.lo.cc.u32 t161, %13, t136, t161;\n\t"
        "madc.hi.cc.u32 t162, %13, t136, t162;\n\t"
        "madc.lo.cc.u32 t163, %15, t136, 0x0;\n\t"
        "madc.hi.u32 t164, %15, t136, 0x0;\n\t"
        "mad.lo.cc.u32 t172, %8, t136, t172;\n\t"
        "madc.hi.cc.u32 t173, %8, t136, t173;\n\t"
        "madc.lo.cc.u32 t174, %10, t136, t174;\n\t"
        "madc.hi.cc.u32 t175, %10, t136, t175;\n\t"
        "madc.lo.cc.u32 t176, %12, t136, t176;\n\t"
        "madc.hi.cc.u32 t177, %12, t136, t177;\n\t"
        "madc.lo.cc.u32 t178, %14, t136, t178;\n\t"
        "madc.hi.cc.u32 t179, %14, t136, t179;\n\t"
        "addc.u32 t180, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t157, %8, t139, t157;\n\t"
        "madc.hi.cc.u32 t158, %8, t139, t158;\n\t"
        "madc.lo.cc.u32 t159, %10, t139, t159;\n\t"
        "madc.hi.cc.u32 t160, %10, t139, t160;\n\t"
        "madc.lo.cc.u32 t161, %12, t139, t161;\n\t"
        "madc.hi.cc.u32 t162, %12, t139, t162;\n\t"
        "madc.lo.cc.u32 t163, %14, t139, t163;\n\t"
        "madc.hi.cc.u32 t164, %14, t139, t164;\n\t"
        "addc.u32 t165, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t174, %9, t139, t174;\n\t"
        "madc.hi.cc.u32 t175, %9, t139, t175;\n\t"
        "madc.lo.cc.u32 t176, %11, t139, t176;\n\t"
        "madc.hi.cc.u32 t177, %11, t139, t177;\n\t"
        "madc.lo.cc.u32 t178, %13, t139, t178;\n\t"
        "madc.hi.cc.u32 t179, %13, t139, t179;\n\t"
        "madc.lo.cc.u32 t180, %15, t139, t180;\n\t"
        "madc.hi.u32 t181, %15, t139, 0x0;\n\t"
        "mad.lo.cc.u32 t159, %9, t142, t159;\n\t"
        "madc.hi.cc.u32 t160, %9, t142, t160;\n\t"
        "madc.lo.cc.u32 t161, %11, t142, t161;\n\t"
        "madc.hi.cc.u32 t162, %11, t142, t162;\n\t"
        "madc.lo.cc.u32 t163, %13, t142, t163;\n\t"
        "madc.hi.cc.u32 t164, %13, t142, t164;\n\t"
        "madc.lo.cc.u32 t165, %15, t142, t165;\n\t"
        "madc.hi.u32 t166, %15, t142, 0x0;\n\t"
        "mad.lo.cc.u32 t174, %8, t142, t174;\n\t"
        "madc.hi.cc.u32 t175, %8, t142, t175;\n\t"
        "madc.lo.cc.u32 t176, %10, t142, t176;\n\t"
        "madc.hi.cc.u32 t177, %10, t142, t177;\n\t"
        "madc.lo.cc.u32 t178, %12, t142, t178;\n\t"
        "madc.hi.cc.u32 t179, %12, t142, t179;\n\t"
        "madc.lo.cc.u32 t180, %14, t142, t180;\n\t"
        "madc.hi.cc.u32 t181, %14, t142, t181;\n\t"
        "addc.u32 t182, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t159, %8, t145, t159;\n\t"
        "madc.hi.cc.u32 t160, %8, t145, t160;\n\t"
        "madc.lo.cc.u32 t161, %10, t145, t161;\n\t"
        "madc.hi.cc.u32 t162, %10, t145, t162;\n\t"
        "madc.lo.cc.u32 t163, %12, t145, t163;\n\t"
        "madc.hi.cc.u32 t164, %12, t145, t164;\n\t"
        "madc.lo.cc.u32 t165, %14, t145, t165;\n\t"
        "madc.hi.cc.u32 t166, %14, t145, t166;\n\t"
        "addc.u32 t167, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t176, %9, t145, t176;\n\t"
        "madc.hi.cc.u32 t177, %9, t145, t177;\n\t"
        "madc.lo.cc.u32 t178, %11, t145, t178;\n\t"
        "madc.hi.cc.u32 t179, %11, t145, t179;\n\t"
        "madc.lo.cc.u32 t180, %13, t145, t180;\n\t"
        "madc.hi.cc.u32 t181, %13, t145, t181;\n\t"
        "madc.lo.cc.u32 t182, %15, t145, t182;\n\t"
        "madc.hi.u32 t183, %15, t145, 0x0;\n\t"
        "mad.lo.cc.u32 t161, %9, t148, t161;\n\t"
        "madc.hi.cc.u32 t162, %9, t148, t162;\n\t"
        "madc.lo.cc.u32 t163, %11, t148, t163;\n\t"
        "madc.hi.cc.u32 t164, %11, t148, t164;\n\t"
        "madc.lo.cc.u32 t165, %13, t148, t165;\n\t"
        "madc.hi.cc.u32 t166, %13, t148, t166;\n\t"
        "madc.lo.cc.u32 t167, %15, t148, t167;\n\t"
        "madc.hi.u32 t168, %15, t148, 0x0;\n\t"
        "mad.lo.cc.u32 t176, %8, t148, t176;\n\t"
        "madc.hi.cc.u32 t177, %8, t148, t177;\n\t"
        "madc.lo.cc.u32 t178, %10, t148, t178;\n\t"
        "madc.hi.cc.u32 t179, %10, t148, t179;\n\t"
        "madc.lo.cc.u32 t180, %12, t148, t180;\n\t"
        "madc.hi.cc.u32 t181, %12, t148, t181;\n\t"
        "madc.lo.cc.u32 t182, %14, t148, t182;\n\t"
        "madc.hi.cc.u32 t183, %14, t148, t183;\n\t"
        "addc.u32 t184, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t161, %8, t151, t161;\n\t"
        "madc.hi.cc.u32 t162, %8, t151, t162;\n\t"
        "madc.lo.cc.u32 t163, %10, t151, t163;\n\t"
        "madc.hi.cc.u32 t164, %10, t151, t164;\n\t"
        "madc.lo.cc.u32 t165, %12, t151, t165;\n\t"
        "madc.hi.cc.u32 t166, %12, t151, t166;\n\t"
        "madc.lo.cc.u32 t167, %14, t151, t167;\n\t"
        "madc.hi.cc.u32 t168, %14, t151, t168;\n\t"
        "addc.u32 t169, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t178, %9, t151, t178;\n\t"
        "madc.hi.cc.u32 t179, %9, t151, t179;\n\t"
        "madc.lo.cc.u32 t180, %11, t151, t180;\n\t"
        "madc.hi.cc.u32 t181, %11, t151, t181;\n\t"
        "madc.lo.cc.u32 t182, %13, t151, t182;\n\t"
        "madc.hi.cc.u32 t183, %13, t151, t183;\n\t"
        "madc.lo.cc.u32 t184, %15, t151, t184;\n\t"
        "madc.hi.u32 t185, %15, t151, 0x0;\n\t"
        "mad.lo.cc.u32 t163, %9, t154, t163;\n\t"
        "madc.hi.cc.u32 t164, %9, t154, t164;\n\t"
        "madc.lo.cc.u32 t165, %11, t154, t165;\n\t"
        "madc.hi.cc.u32 t166, %11, t154, t166;\n\t"
        "madc.lo.cc.u32 t167, %13, t154, t167;\n\t"
        "madc.hi.cc.u32 t168, %13, t154, t168;\n\t"
        "madc.lo.cc.u32 t169, %15, t154, t169;\n\t"
        "madc.hi.u32 t170, %15, t154, 0x0;\n\t"
        "mad.lo.cc.u32 t178, %8, t154, t178;\n\t"
        "madc.hi.cc.u32 t179, %8, t154, t179;\n\t"
        "madc.lo.cc.u32 t180, %10, t154, t180;\n\t"
        "madc.hi.cc.u32 t181, %10, t154, t181;\n\t"
        "madc.lo.cc.u32 t182, %12, t154, t182;\n\t"
        "madc.hi.cc.u32 t183, %12, t154, t183;\n\t"
        "madc.lo.cc.u32 t184, %14, t154, t184;\n\t"
        "madc.hi.cc.u32 t185, %14, t154, t185;\n\t"
        "addc.u32 t186, 0x0, 0x0;\n\t"
        "add.cc.u32 t187, t156, t172;\n\t"
        "addc.cc.u32 t188, t157, t173;\n\t"
        "addc.cc.u32 t189, t158, t174;\n\t"
        "addc.cc.u32 t190, t159, t175;\n\t"
        "addc.cc.u32 t191, t160, t176;\n\t"
        "addc.cc.u32 t192, t161, t177;\n\t"
        "addc.cc.u32 t193, t162, t178;\n\t"
        "addc.cc.u32 t194, t163, t179;\n\t"
        "addc.cc.u32 t195, t164, t180;\n\t"
        "addc.cc.u32 t196, t165, t181;\n\t"
        "addc.cc.u32 t197, t166, t182;\n\t"
        "addc.cc.u32 t198, t167, t183;\n\t"
        "addc.cc.u32 t199, t168, t184;\n\t"
        "addc.cc.u32 t200, t169, t185;\n\t"
        "addc.u32 t201, t170, t186;\n\t"
        "mul.lo.u32 t202, t155, 0xee00bc4f;\n\t"
        "mul.hi.u32 t203, t155, 0xee00bc4f;\n\t"
        "mul.lo.u32 t204, t188, 0xee00bc4f;\n\t"
        "mul.hi.u32 t205, t188, 0xee00bc4f;\n\t"
        "mul.lo.u32 t206, t190, 0xee00bc4f;\n\t"
        "mul.hi.u32 t207, t190, 0xee00bc4f;\n\t"
        "mul.lo.u32 t208, t192, 0xee00bc4f;\n\t"
        "mul.hi.u32 t209, t192, 0xee00bc4f;\n\t"
        "mul.lo.u32 t212, t187, 0xee00bc4f;\n\t"
        "mul.hi.u32 t213, t187, 0xee00bc4f;\n\t"
        "mul.lo.u32 t214, t189, 0xee00bc4f;\n\t"
        "mul.hi.u32 t215, t189, 0xee00bc4f;\n\t"
        "mul.lo.u32 t216, t191, 0xee00bc4f;\n\t"
        "mul.hi.u32 t217, t191, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t204, t187, 0xccd1c8aa, t204;\n\t"
        "madc.hi.cc.u32 t205, t187, 0xccd1c8aa, t205;\n\t"
        "madc.lo.cc.u32 t206, t189, 0xccd1c8aa, t206;\n\t"
        "madc.hi.cc.u32 t207, t189, 0xccd1c8aa, t207;\n\t"
        "madc.lo.cc.u32 t208, t191, 0xccd1c8aa, t208;\n\t"
        "madc.hi.cc.u32 t209, t191, 0xccd1c8aa, t209;\n\t"
        "addc.u32 t210, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t212, t155, 0xccd1c8aa, t212;\n\t"
        "madc.hi.cc.u32 t213, t155, 0xccd1c8aa, t213;\n\t"
        "madc.lo.cc.u32 t214, t188, 0xccd1c8aa, t214;\n\t"
        "madc.hi.cc.u32 t215, t188, 0xccd1c8aa, t215;\n\t"
        "madc.lo.cc.u32 t216, t190, 0xccd1c8aa, t216;\n\t"
        "madc.hi.cc.u32 t217, t190, 0xccd1c8aa, t217;\n\t"
        "addc.u32 t218, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t204, t155, 0x7d74d2e4, t204;\n\t"
        "madc.hi.cc.u32 t205, t155, 0x7d74d2e4, t205;\n\t"
        "madc.lo.cc.u32 t206, t188, 0x7d74d2e4, t206;\n\t"
        "madc.hi.cc.u32 t207, t188, 0x7d74d2e4, t207;\n\t"
        "madc.lo.cc.u32 t208, t190, 0x7d74d2e4, t208;\n\t"
        "madc.hi.cc.u32 t209, t190, 0x7d74d2e4, t209;\n\t"
        "addc.u32 t210, t210, 0x0;\n\t"
        "mad.lo.cc.u32 t214, t187, 0x7d74d2e4, t214;\n\t"
        "madc.hi.cc.u32 t215, t187, 0x7d74d2e4, t215;\n\t"
        "madc.lo.cc.u32 t216, t189, 0x7d74d2e4, t216;\n\t"
        "madc.hi.cc.u32 t217, t189, 0x7d74d2e4, t217;\n\t"
        "addc.u32 t218, t218, 0x0;\n\t"
        "mad.lo.cc.u32 t206, t187, 0x48c94408, t206;\n\t"
        "madc.hi.cc.u32 t207, t187, 0x48c94408, t207;\n\t"
        "madc.lo.cc.u32 t208, t189, 0x48c94408, t208;\n\t"
        "madc.hi.cc.u32 t209, t189, 0x48c94408, t209;\n\t"
        "addc.u32 t210, t210, 0x0;\n\t"
        "mad.lo.cc.u32 t214, t155, 0x48c94408, t214;\n\t"
        "madc.hi.cc.u32 t215, t155, 0x48c94408, t215;\n\t"
        "madc.lo.cc.u32 t216, t188, 0x48c94408, t216;\n\t"
        "madc.hi.cc.u32 t217, t188, 0x48c94408, t217;\n\t"
        "addc.u32 t218, t218, 0x0;\n\t"
        "mad.lo.cc.u32 t206, t155, 0xc588c6f6, t206;\n\t"
        "madc.hi.cc.u32 t207, t155, 0xc588c6f6, t207;\n\t"
        "madc.lo.cc.u32 t208, t188, 0xc588c6f6, t208;\n\t"
        "madc.hi.cc.u32 t209, t188, 0xc588c6f6, t209;\n\t"
        "addc.u32 t210, t210, 0x0;\n\t"
        "mad.lo.cc.u32 t216, t187, 0xc588c6f6, t216;\n\t"
        "madc.hi.cc.u32 t217, t187, 0xc588c6f6, t217;\n\t"
        "addc.u32 t218, t218, 0x0;\n\t"
        "mad.lo.cc.u32 t208, t187, 0x50fe77ec, t208;\n\t"
        "madc.hi.cc.u32 t209, t187, 0x50fe77ec, t209;\n\t"
        "addc.u32 t210, t210, 0x0;\n\t"
        "mad.lo.cc.u32 t216, t155, 0x50fe77ec, t216;\n\t"
        "madc.hi.cc.u32 t217, t155, 0x50fe77ec, t217;\n\t"
        "addc.u32 t218, t218, 0x0;\n\t"
        "mad.lo.cc.u32 t208, t155, 0xa9d6281c, t208;\n\t"
        "madc.hi.cc.u32 t209, t155, 0xa9d6281c, t209;\n\t"
        "addc.u32 t210, t210, 0x0;\n\t"
        "add.cc.u32 t220, t203, t212;\n\t"
        "addc.cc.u32 t221, t204, t213;\n\t"
        "addc.cc.u32 t222, t205, t214;\n\t"
        "addc.cc.u32 t223, t206, t215;\n\t"
        "addc.cc.u32 t224, t207, t216;\n\t"
        "addc.cc.u32 t225, t208, t217;\n\t"
        "addc.u32 t226, t209, t218;\n\t"
        "mad.lo.u32 t227, t193, 0xee00bc4f, t226;\n\t"
        "mad.lo.u32 t228, t192, 0xccd1c8aa, t227;\n\t"
        "mad.lo.u32 t229, t191, 0x7d74d2e4, t228;\n\t"
        "mad.lo.u32 t230, t190, 0x48c94408, t229;\n\t"
        "mad.lo.u32 t231, t189, 0xc588c6f6, t230;\n\t"
        "mad.lo.u32 t232, t188, 0x50fe77ec, t231;\n\t"
        "mad.lo.u32 t233, t187, 0xa9d6281c, t232;\n\t"
        "mad.lo.u32 t234, t155, 0x60d06633, t233;\n\t"
        "mul.lo.u32 t235, t202, 0xfc632551;\n\t"
        "mul.hi.u32 t236, t202, 0xfc632551;\n\t"
        "mul.lo.u32 t237, t221, 0xfc632551;\n\t"
        "mul.hi.u32 t238, t221, 0xfc632551;\n\t"
        "mul.lo.u32 t239, t223, 0xfc632551;\n\t"
        "mul.hi.u32 t240, t223, 0xfc632551;\n\t"
        "mul.lo.u32 t241, t225, 0xfc632551;\n\t"
        "mul.hi.u32 t242, t225, 0xfc632551;\n\t"
        "mul.lo.u32 t252, t220, 0xfc632551;\n\t"
        "mul.hi.u32 t253, t220, 0xfc632551;\n\t"
        "mul.lo.u32 t254, t222, 0xfc632551;\n\t"
        "mul.hi.u32 t255, t222, 0xfc632551;\n\t"
        "mul.lo.u32 t256, t224, 0xfc632551;\n\t"
        "mul.hi.u32 t257, t224, 0xfc632551;\n\t"
        "mul.lo.u32 t258, t234, 0xfc632551;\n\t"
        "mul.hi.u32 t259, t234, 0xfc632551;\n\t"
        "mad.lo.cc.u32 t237, t220, 0xf3b9cac2, t237;\n\t"
        "madc.hi.cc.u32 t238, t220, 0xf3b9cac2, t238;\n\t"
        "madc.lo.cc.u32 t239, t222, 0xf3b9cac2, t239;\n\t"
        "madc.hi.cc.u32 t240, t222, 0xf3b9cac2, t240;\n\t"
        "madc.lo.cc.u32 t241, t224, 0xf3b9cac2, t241;\n\t"
        "madc.hi.cc.u32 t242, t224, 0xf3b9cac2, t242;\n\t"
        "madc.lo.cc.u32 t243, t234, 0xf3b9cac2, 0x0;\n\t"
        "madc.hi.u32 t244, t234, 0xf3b9cac2, 0x0;\n\t"
        "mad.lo.cc.u32 t252, t202, 0xf3b9cac2, t252;\n\t"
        "madc.hi.cc.u32 t253, t202, 0xf3b9cac2, t253;\n\t"
        "madc.lo.cc.u32 t254, t221, 0xf3b9cac2, t254;\n\t"
        "madc.hi.cc.u32 t255, t221, 0xf3b9cac2, t255;\n\t"
        "madc.lo.cc.u32 t256, t223, 0xf3b9cac2, t256;\n\t"
        "madc.hi.cc.u32 t257, t223, 0xf3b9cac2, t257;\n\t"
        "madc.lo.cc.u32 t258, t225, 0xf3b9cac2, t258;\n\t"
        "madc.hi.cc.u32 t259, t225, 0xf3b9cac2, t259;\n\t"
        "addc.u32 t260, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t237, t202, 0xa7179e84, t237;\n\t"
        "madc.hi.cc.u32 t238, t202, 0xa7179e84, t238;\n\t"
        "madc.lo.cc.u32 t239, t221, 0xa7179e84, t239;\n\t"
        "madc.hi.cc.u32 t240, t221, 0xa7179e84, t240;\n\t"
        "madc.lo.cc.u32 t241, t223, 0xa7179e84, t241;\n\t"
        "madc.hi.cc.u32 t242, t223, 0xa7179e84, t242;\n\t"
        "madc.lo.cc.u32 t243, t225, 0xa7179e84, t243;\n\t"
        "madc.hi.cc.u32 t244, t225, 0xa7179e84, t244;\n\t"
        "addc.u32 t245, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t254, t220, 0xa7179e84, t254;\n\t"
        "madc.hi.cc.u32 t255, t220, 0xa7179e84, t255;\n\t"
        "madc.lo.cc.u32 t256, t222, 0xa7179e84, t256;\n\t"
        "madc.hi.cc.u32 t257, t222, 0xa7179e84, t257;\n\t"
        "madc.lo.cc.u32 t258, t224, 0xa7179e84, t258;\n\t"
        "madc.hi.cc.u32 t259, t224, 0xa7179e84, t259;\n\t"
        "madc.lo.cc.u32 t260, t234, 0xa7179e84, t260;\n\t"
        "madc.hi.u32 t261, t234, 0xa7179e84, 0x0;\n\t"
        "mad.lo.cc.u32 t239, t220, 0xbce6faad, t239;\n\t"
        "madc.hi.cc.u32 t240, t220, 0xbce6faad, t240;\n\t"
        "madc.lo.cc.u32 t241, t222, 0xbce6faad, t241;\n\t"
        "madc.hi.cc.u32 t242, t222, 0xbce6faad, t242;\n\t"
        "madc.lo.cc.u32 t243, t224, 0xbce6faad, t243;\n\t"
        "madc.hi.cc.u32 t244, t224, 0xbce6faad, t244;\n\t"
        "madc.lo.cc.u32 t245, t234, 0xbce6faad, t245;\n\t"
        "madc.hi.u32 t246, t234, 0xbce6faad, 0x0;\n\t"
        "mad.lo.cc.u32 t254, t202, 0xbce6faad, t254;\n\t"
        "madc.hi.cc.u32 t255, t202, 0xbce6faad, t255;\n\t"
        "madc.lo.cc.u32 t256, t221, 0xbce6faad, t256;\n\t"
        "madc.hi.cc.u32 t257, t221, 0xbce6faad, t257;\n\t"
        "madc.lo.cc.u32 t258, t223, 0xbce6faad, t258;\n\t"
        "madc.hi.cc.u32 t259, t223, 0xbce6faad, t259;\n\t"
        "madc.lo.cc.u32 t260, t225, 0xbce6faad, t260;\n\t"
        "madc.hi.cc.u32 t261, t225, 0xbce6faad, t261;\n\t"
        "addc.u32 t262, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t239, t202, 0xffffffff, t239;\n\t"
        "madc.hi.cc.u32 t240, t202, 0xffffffff, t240;\n\t"
        "madc.lo.cc.u32 t241, t221, 0xffffffff, t241;\n\t"
        "madc.hi.cc.u32 t242, t221, 0xffffffff, t242;\n\t"
        "madc.lo.cc.u32 t243, t223, 0xffffffff, t243;\n\t"
        "madc.hi.cc.u32 t244, t223, 0xffffffff, t244;\n\t"
        "madc.lo.cc.u32 t245, t225, 0xffffffff, t245;\n\t"
        "madc.hi.cc.u32 t246, t225, 0xffffffff, t246;\n\t"
        "addc.u32 t247, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t256, t220, 0xffffffff, t256;\n\t"
        "madc.hi.cc.u32 t257, t220, 0xffffffff, t257;\n\t"
        "madc.lo.cc.u32 t258, t222, 0xffffffff, t258;\n\t"
        "madc.hi.cc.u32 t259, t222, 0xffffffff, t259;\n\t"
        "madc.lo.cc.u32 t260, t224, 0xffffffff, t260;\n\t"
        "madc.hi.cc.u32 t261, t224, 0xffffffff, t261;\n\t"
        "madc.lo.cc.u32 t262, t234, 0xffffffff, t262;\n\t"
        "madc.hi.u32 t263, t234, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t241, t220, 0xffffffff, t241;\n\t"
        "madc.hi.cc.u32 t242, t220, 0xffffffff, t242;\n\t"
        "madc.lo.cc.u32 t243, t222, 0xffffffff, t243;\n\t"
        "madc.hi.cc.u32 t244, t222, 0xffffffff, t244;\n\t"
        "madc.lo.cc.u32 t245, t224, 0xffffffff, t245;\n\t"
        "madc.hi.cc.u32 t246, t224, 0xffffffff, t246;\n\t"
        "madc.lo.cc.u32 t247, t234, 0xffffffff, t247;\n\t"
        "madc.hi.u32 t248, t234, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t256, t202, 0xffffffff, t256;\n\t"
        "madc.hi.cc.u32 t257, t202, 0xffffffff, t257;\n\t"
        "madc.lo.cc.u32 t258, t221, 0xffffffff, t258;\n\t"
        "madc.hi.cc.u32 t259, t221, 0xffffffff, t259;\n\t"
        "madc.lo.cc.u32 t260, t223, 0xffffffff, t260;\n\t"
        "madc.hi.cc.u32 t261, t223, 0xffffffff, t261;\n\t"
        "madc.lo.cc.u32 t262, t225, 0xffffffff, t262;\n\t"
        "madc.hi.cc.u32 t263, t225, 0xffffffff, t263;\n\t"
        "addc.u32 t264, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t241, t202, 0x0, t241;\n\t"
        "madc.hi.cc.u32 t242, t202, 0x0, t242;\n\t"
        "madc.lo.cc.u32 t243, t221, 0x0, t243;\n\t"
        "madc.hi.cc.u32 t244, t221, 0x0, t244;\n\t"
        "madc.lo.cc.u32 t245, t223, 0x0, t245;\n\t"
        "madc.hi.cc.u32 t246, t223, 0x0, t246;\n\t"
        "madc.lo.cc.u32 t247, t225, 0x0, t247;\n\t"
        "madc.hi.cc.u32 t248, t225, 0x0, t248;\n\t"
        "addc.u32 t249, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t258, t220, 0x0, t258;\n\t"
        "madc.hi.cc.u32 t259, t220, 0x0, t259;\n\t"
        "madc.lo.cc.u32 t260, t222, 0x0, t260;\n\t"
        "madc.hi.cc.u32 t261, t222, 0x0, t261;\n\t"
        "madc.lo.cc.u32 t262, t224, 0x0, t262;\n\t"
        "madc.hi.cc.u32 t263, t224, 0x0, t263;\n\t"
        "madc.lo.cc.u32 t264, t234, 0x0, t264;\n\t"
        "madc.hi.u32 t265, t234, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t243, t220, 0xffffffff, t243;\n\t"
        "madc.hi.cc.u32 t244, t220, 0xffffffff, t244;\n\t"
        "madc.lo.cc.u32 t245, t222, 0xffffffff, t245;\n\t"
        "madc.hi.cc.u32 t246, t222, 0xffffffff, t246;\n\t"
        "madc.lo.cc.u32 t247, t224, 0xffffffff, t247;\n\t"
        "madc.hi.cc.u32 t248, t224, 0xffffffff, t248;\n\t"
        "madc.lo.cc.u32 t249, t234, 0xffffffff, t249;\n\t"
        "madc.hi.u32 t250, t234, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t258, t202, 0xffffffff, t258;\n\t"
        "madc.hi.cc.u32 t259, t202, 0xffffffff, t259;\n\t"
        "madc.lo.cc.u32 t260, t221, 0xffffffff, t260;\n\t"
        "madc.hi.cc.u32 t261, t221, 0xffffffff, t261;\n\t"
        "madc.lo.cc.u32 t262, t223, 0xffffffff, t262;\n\t"
        "madc.hi.cc.u32 t263, t223, 0xffffffff, t263;\n\t"
        "madc.lo.cc.u32 t264, t225, 0xffffffff, t264;\n\t"
        "madc.hi.cc.u32 t265, t225, 0xffffffff, t265;\n\t"
        "addc.u32 t266, 0x0, 0x0;\n\t"
        "add.cc.u32 t267, t236, t252;\n\t"
        "addc.cc.u32 t268, t237, t253;\n\t"
        "addc.cc.u32 t269, t238, t254;\n\t"
        "addc.cc.u32 t270, t239, t255;\n\t"
        "addc.cc.u32 t271, t240, t256;\n\t"
        "addc.cc.u32 t272, t241, t257;\n\t"
        "addc.cc.u32 t273, t242, t258;\n\t"
        "addc.cc.u32 t274, t243, t259;\n\t"
        "addc.cc.u32 t275, t244, t260;\n\t"
        "addc.cc.u32 t276, t245, t261;\n\t"
        "addc.cc.u32 t277, t246, t262;\n\t"
        "addc.cc.u32 t278, t247, t263;\n\t"
        "addc.cc.u32 t279, t248, t264;\n\t"
        "addc.cc.u32 t280, t249, t265;\n\t"
        "addc.u32 t281, t250, t266;\n\t"
        "add.cc.u32 t282, t155, t235;\n\t"
        "addc.cc.u32 t283, t187, t267;\n\t"
        "addc.cc.u32 t284, t188, t268;\n\t"
        "addc.cc.u32 t285, t189, t269;\n\t"
        "addc.cc.u32 t286, t190, t270;\n\t"
        "addc.cc.u32 t287, t191, t271;\n\t"
        "addc.cc.u32 t288, t192, t272;\n\t"
        "addc.cc.u32 t289, t193, t273;\n\t"
        "addc.cc.u32 t290, t194, t274;\n\t"
        "addc.cc.u32 t291, t195, t275;\n\t"
        "addc.cc.u32 t292, t196, t276;\n\t"
        "addc.cc.u32 t293, t197, t277;\n\t"
        "addc.cc.u32 t294, t198, t278;\n\t"
        "addc.cc.u32 t295, t199, t279;\n\t"
        "addc.cc.u32 t296, t200, t280;\n\t"
        "addc.cc.u32 t297, t201, t281;\n\t"
        "addc.u32 t298, 0x0, 0x0;\n\t"
        "sub.cc.u32 t299, t290, 0xfc632551;\n\t"
        "subc.cc.u32 t300, t291, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t301, t292, 0xa7179e84;\n\t"
        "subc.cc.u32 t302, t293, 0xbce6faad;\n\t"
        "subc.cc.u32 t303, t294, 0xffffffff;\n\t"
        "subc.cc.u32 t304, t295, 0xffffffff;\n\t"
        "subc.cc.u32 t305, t296, 0x0;\n\t"
        "subc.cc.u32 t306, t297, 0xffffffff;\n\t"
        "subc.cc.u32 t307, t298, 0x0;\n\t"
        "subc.u32 t308, 0x0, 0x0;\n\t"
        "xor.b32 t309, t299, t290;\n\t"
        "and.b32 t310, t309, t308;\n\t"
        "xor.b32 t311, t310, t299;\n\t"
        "xor.b32 t312, t300, t291;\n\t"
        "and.b32 t313, t312, t308;\n\t"
        "xor.b32 t314, t313, t300;\n\t"
        "xor.b32 t315, t301, t292;\n\t"
        "and.b32 t316, t315, t308;\n\t"
        "xor.b32 t317, t316, t301;\n\t"
        "xor.b32 t318, t302, t293;\n\t"
        "and.b32 t319, t318, t308;\n\t"
        "xor.b32 t320, t319, t302;\n\t"
        "xor.b32 t321, t303, t294;\n\t"
        "and.b32 t322, t321, t308;\n\t"
        "xor.b32 t323, t322, t303;\n\t"
        "xor.b32 t324, t304, t295;\n\t"
        "and.b32 t325, t324, t308;\n\t"
        "xor.b32 t326, t325, t304;\n\t"
        "xor.b32 t327, t305, t296;\n\t"
        "and.b32 t328, t327, t308;\n\t"
        "xor.b32 t329, t328, t305;\n\t"
        "xor.b32 t330, t306, t297;\n\t"
        "and.b32 t331, t330, t308;\n\t"
        "xor.b32 t332, t331, t306;\n\t"
        "mov.u32 %0, t311;\n\t"
        "mov.u32 %1, t314;\n\t"
        "mov.u32 %2, t317;\n\t"
        "mov.u32 %3, t320;\n\t"
        "mov.u32 %4, t323;\n\t"
        "mov.u32 %5, t326;\n\t"
        "mov.u32 %6, t329;\n\t"
        "mov.u32 %7, t332;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127, t128, t129, t130, t131, t132, t133, t134, t135, t136, t137, t138, t139, t140, t141, t142, t143, t144, t145, t146, t147, t148, t149, t150, t151, t152, t153, t154, t155, t156, t157, t158, t159, t160, t161, t162, t163, t164, t165, t166, t167, t168, t169, t170, t171, t172, t173, t174, t175, t176, t177, t178, t179, t180, t181, t182, t183, t184, t185, t186, t187, t188, t189, t190, t191, t192, t193, t194, t195, t196, t197, t198, t199, t200, t201, t202, t203, t204, t205, t206, t207, t208, t209, t210, t211, t212, t213, t214, t215, t216, t217, t218, t219, t220, t221, t222, t223, t224, t225, t226, t227, t228, t229, t230, t231, t232, t233, t234, t235, t236, t237, t238, t239, t240, t241, t242, t243, t244, t245, t246, t247, t248, t249, t250, t251, t252, t253, t254, t255, t256, t257, t258, t259, t260, t261, t262, t263, t264, t265, t266, t267, t268, t269, t270, t271, t272, t273, t274, t275, t276, t277, t278, t279, t280, t281, t282, t283, t284, t285, t286, t287, t288, t289, t290, t291, t292, t293, t294, t295, t296, t297, t298, t299, t300, t301, t302, t303, t304, t305, t306, t307, t308, t309, t310, t311, t312, t313, t314, t315, t316, t317, t318, t319, t320, t321, t322, t323, t324, t325, t326, t327, t328, t329, t330, t331, t332;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t0 = (uint32_t)((uint32_t)(0xbe79eea2u * b_i));
    t8 = (uint32_t)(((uint64_t)0xbe79eea2u * b_i) >> 32);
    t1 = (uint32_t)((uint32_t)(0x83244c95u * b_i));
    t9 = (uint32_t)(((uint64_t)0x83244c95u * b_i) >> 32);
    t2 = (uint32_t)((uint32_t)(0x49bd6fa6u * b_i));
    t10 = (uint32_t)(((uint64_t)0x49bd6fa6u * b_i) >> 32);
    t3 = (uint32_t)((uint32_t)(0x4699799cu * b_i));
    t11 = (uint32_t)(((uint64_t)0x4699799cu * b_i) >> 32);
    t4 = (uint32_t)((uint32_t)(0x2b6bec59u * b_i));
    t12 = (uint32_t)(((uint64_t)0x2b6bec59u * b_i) >> 32);
    t5 = (uint32_t)((uint32_t)(0x2845b239u * b_i));
    t13 = (uint32_t)(((uint64_t)0x2845b239u * b_i) >> 32);
    t6 = (uint32_t)((uint32_t)(0xf3d95620u * b_i));
    t14 = (uint32_t)(((uint64_t)0xf3d95620u * b_i) >> 32);
    t7 = (uint32_t)((uint32_t)(0x66e12d94u * b_i));
    t15 = (uint32_t)(((uint64_t)0x66e12d94u * b_i) >> 32);
    w_ = (uint64_t)t1 + t8; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t9 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t10 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t11 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t12 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t13 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t14 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + 0x0u + cf_; t23 = (uint32_t)w_;
    t24 = (uint32_t)((uint32_t)(t0 * 0xee00bc4fu));
    t25 = (uint32_t)(((uint64_t)t0 * 0xee00bc4fu) >> 32);
    t26 = (uint32_t)((uint32_t)(t17 * 0xee00bc4fu));
    t27 = (uint32_t)(((uint64_t)t17 * 0xee00bc4fu) >> 32);
    t28 = (uint32_t)((uint32_t)(t19 * 0xee00bc4fu));
    t29 = (uint32_t)(((uint64_t)t19 * 0xee00bc4fu) >> 32);
    t30 = (uint32_t)((uint32_t)(t21 * 0xee00bc4fu));
    t31 = (uint32_t)(((uint64_t)t21 * 0xee00bc4fu) >> 32);
    t34 = (uint32_t)((uint32_t)(t16 * 0xee00bc4fu));
    t35 = (uint32_t)(((uint64_t)t16 * 0xee00bc4fu) >> 32);
    t36 = (uint32_t)((uint32_t)(t18 * 0xee00bc4fu));
    t37 = (uint32_t)(((uint64_t)t18 * 0xee00bc4fu) >> 32);
    t38 = (uint32_t)((uint32_t)(t20 * 0xee00bc4fu));
    t39 = (uint32_t)(((uint64_t)t20 * 0xee00bc4fu) >> 32);
    w_ = (uint64_t)(uint32_t)(t16 * 0xccd1c8aau) + t26; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t16 * 0xccd1c8aau) >> 32) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t18 * 0xccd1c8aau) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t18 * 0xccd1c8aau) >> 32) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t20 * 0xccd1c8aau) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t20 * 0xccd1c8aau) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0xccd1c8aau) + t34; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0xccd1c8aau) >> 32) + t35 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t17 * 0xccd1c8aau) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t17 * 0xccd1c8aau) >> 32) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t19 * 0xccd1c8aau) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t19 * 0xccd1c8aau) >> 32) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t40 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x7d74d2e4u) + t26; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0x7d74d2e4u) >> 32) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t17 * 0x7d74d2e4u) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t17 * 0x7d74d2e4u) >> 32) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t19 * 0x7d74d2e4u) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t19 * 0x7d74d2e4u) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t16 * 0x7d74d2e4u) + t36; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t16 * 0x7d74d2e4u) >> 32) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t18 * 0x7d74d2e4u) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t18 * 0x7d74d2e4u) >> 32) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + 0x0u + cf_; t40 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t16 * 0x48c94408u) + t28; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t16 * 0x48c94408u) >> 32) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t18 * 0x48c94408u) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t18 * 0x48c94408u) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x48c94408u) + t36; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0x48c94408u) >> 32) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t17 * 0x48c94408u) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t17 * 0x48c94408u) >> 32) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + 0x0u + cf_; t40 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0xc588c6f6u) + t28; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0xc588c6f6u) >> 32) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t17 * 0xc588c6f6u) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t17 * 0xc588c6f6u) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t16 * 0xc588c6f6u) + t38; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t16 * 0xc588c6f6u) >> 32) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + 0x0u + cf_; t40 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t16 * 0x50fe77ecu) + t30; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t16 * 0x50fe77ecu) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x50fe77ecu) + t38; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0x50fe77ecu) >> 32) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + 0x0u + cf_; t40 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0xa9d6281cu) + t30; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0xa9d6281cu) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)t25 + t34; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t26 + t35 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t27 + t36 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t28 + t37 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t29 + t38 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t30 + t39 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t31 + t40 + cf_; t48 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t22 * 0xee00bc4fu) + t48; t49 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t21 * 0xccd1c8aau) + t49; t50 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t20 * 0x7d74d2e4u) + t50; t51 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t19 * 0x48c94408u) + t51; t52 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t18 * 0xc588c6f6u) + t52; t53 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t17 * 0x50fe77ecu) + t53; t54 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t16 * 0xa9d6281cu) + t54; t55 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x60d06633u) + t55; t56 = (uint32_t)w_;
    t57 = (uint32_t)((uint32_t)(t24 * 0xfc632551u));
    t58 = (uint32_t)(((uint64_t)t24 * 0xfc632551u) >> 32);
    t59 = (uint32_t)((uint32_t)(t43 * 0xfc632551u));
    t60 = (uint32_t)(((uint64_t)t43 * 0xfc632551u) >> 32);
    t61 = (uint32_t)((uint32_t)(t45 * 0xfc632551u));
    t62 = (uint32_t)(((uint64_t)t45 * 0xfc632551u) >> 32);
    t63 = (uint32_t)((uint32_t)(t47 * 0xfc632551u));
    t64 = (uint32_t)(((uint64_t)t47 * 0xfc632551u) >> 32);
    t74 = (uint32_t)((uint32_t)(t42 * 0xfc632551u));
    t75 = (uint32_t)(((uint64_t)t42 * 0xfc632551u) >> 32);
    t76 = (uint32_t)((uint32_t)(t44 * 0xfc632551u));
    t77 = (uint32_t)(((uint64_t)t44 * 0xfc632551u) >> 32);
    t78 = (uint32_t)((uint32_t)(t46 * 0xfc632551u));
    t79 = (uint32_t)(((uint64_t)t46 * 0xfc632551u) >> 32);
    t80 = (uint32_t)((uint32_t)(t56 * 0xfc632551u));
    t81 = (uint32_t)(((uint64_t)t56 * 0xfc632551u) >> 32);
    w_ = (uint64_t)(uint32_t)(t42 * 0xf3b9cac2u) + t59; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xf3b9cac2u) >> 32) + t60 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xf3b9cac2u) + t61 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xf3b9cac2u) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xf3b9cac2u) + t63 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xf3b9cac2u) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xf3b9cac2u) + 0x0u + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xf3b9cac2u) >> 32) + 0x0u + cf_; t66 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xf3b9cac2u) + t74; t74 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xf3b9cac2u) >> 32) + t75 + cf_; t75 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xf3b9cac2u) + t76 + cf_; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xf3b9cac2u) >> 32) + t77 + cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xf3b9cac2u) + t78 + cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xf3b9cac2u) >> 32) + t79 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xf3b9cac2u) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xf3b9cac2u) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t82 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xa7179e84u) + t59; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xa7179e84u) >> 32) + t60 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xa7179e84u) + t61 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xa7179e84u) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xa7179e84u) + t63 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xa7179e84u) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xa7179e84u) + t65 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xa7179e84u) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t67 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0xa7179e84u) + t76; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xa7179e84u) >> 32) + t77 + cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xa7179e84u) + t78 + cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xa7179e84u) >> 32) + t79 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xa7179e84u) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xa7179e84u) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xa7179e84u) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xa7179e84u) >> 32) + 0x0u + cf_; t83 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0xbce6faadu) + t61; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xbce6faadu) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xbce6faadu) + t63 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xbce6faadu) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xbce6faadu) + t65 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xbce6faadu) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xbce6faadu) + t67 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xbce6faadu) >> 32) + 0x0u + cf_; t68 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xbce6faadu) + t76; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xbce6faadu) >> 32) + t77 + cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xbce6faadu) + t78 + cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xbce6faadu) >> 32) + t79 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xbce6faadu) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xbce6faadu) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xbce6faadu) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xbce6faadu) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t84 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xffffffffu) + t61; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xffffffffu) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xffffffffu) + t63 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xffffffffu) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xffffffffu) + t65 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xffffffffu) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xffffffffu) + t67 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xffffffffu) >> 32) + t68 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t69 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0xffffffffu) + t78; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xffffffffu) >> 32) + t79 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xffffffffu) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xffffffffu) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xffffffffu) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xffffffffu) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xffffffffu) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xffffffffu) >> 32) + 0x0u + cf_; t85 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0xffffffffu) + t63; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xffffffffu) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xffffffffu) + t65 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xffffffffu) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xffffffffu) + t67 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xffffffffu) >> 32) + t68 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xffffffffu) + t69 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xffffffffu) >> 32) + 0x0u + cf_; t70 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xffffffffu) + t78; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xffffffffu) >> 32) + t79 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xffffffffu) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xffffffffu) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xffffffffu) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xffffffffu) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xffffffffu) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xffffffffu) >> 32) + t85 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t86 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0x0u) + t63; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0x0u) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0x0u) + t65 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0x0u) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0x0u) + t67 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0x0u) >> 32) + t68 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0x0u) + t69 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0x0u) >> 32) + t70 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t71 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0x0u) + t80; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0x0u) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0x0u) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0x0u) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0x0u) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0x0u) >> 32) + t85 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0x0u) + t86 + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0x0u) >> 32) + 0x0u + cf_; t87 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0xffffffffu) + t65; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xffffffffu) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xffffffffu) + t67 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xffffffffu) >> 32) + t68 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xffffffffu) + t69 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xffffffffu) >> 32) + t70 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xffffffffu) + t71 + cf_; t71 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xffffffffu) >> 32) + 0x0u + cf_; t72 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xffffffffu) + t80; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xffffffffu) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xffffffffu) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xffffffffu) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xffffffffu) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xffffffffu) >> 32) + t85 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xffffffffu) + t86 + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xffffffffu) >> 32) + t87 + cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t88 = (uint32_t)w_;
    w_ = (uint64_t)t58 + t74; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t59 + t75 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t60 + t76 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t61 + t77 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t62 + t78 + cf_; t93 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + t79 + cf_; t94 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t64 + t80 + cf_; t95 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t65 + t81 + cf_; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t66 + t82 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t67 + t83 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t68 + t84 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t69 + t85 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t70 + t86 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t71 + t87 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t72 + t88 + cf_; t103 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t57; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t16 + t89 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t17 + t90 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t18 + t91 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t19 + t92 + cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t20 + t93 + cf_; t109 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t21 + t94 + cf_; t110 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t22 + t95 + cf_; t111 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t23 + t96 + cf_; t112 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t97 + cf_; t113 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t98 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t99 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t100 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t101 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t102 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t103 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t120 = (uint32_t)w_;
    w_ = (uint64_t)t112 - 0xfc632551u; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t113 - 0xf3b9cac2u - cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t114 - 0xa7179e84u - cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t115 - 0xbce6faadu - cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t116 - 0xffffffffu - cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t117 - 0xffffffffu - cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t118 - 0x0u - cf_; t127 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t119 - 0xffffffffu - cf_; t128 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t120 - 0x0u - cf_; t129 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t130 = (uint32_t)w_;
    t131 = (uint32_t)(t121 ^ t112);
    t132 = (uint32_t)(t131 & t130);
    t133 = (uint32_t)(t132 ^ t121);
    t134 = (uint32_t)(t122 ^ t113);
    t135 = (uint32_t)(t134 & t130);
    t136 = (uint32_t)(t135 ^ t122);
    t137 = (uint32_t)(t123 ^ t114);
    t138 = (uint32_t)(t137 & t130);
    t139 = (uint32_t)(t138 ^ t123);
    t140 = (uint32_t)(t124 ^ t115);
    t141 = (uint32_t)(t140 & t130);
    t142 = (uint32_t)(t141 ^ t124);
    t143 = (uint32_t)(t125 ^ t116);
    t144 = (uint32_t)(t143 & t130);
    t145 = (uint32_t)(t144 ^ t125);
    t146 = (uint32_t)(t126 ^ t117);
    t147 = (uint32_t)(t146 & t130);
    t148 = (uint32_t)(t147 ^ t126);
    t149 = (uint32_t)(t127 ^ t118);
    t150 = (uint32_t)(t149 & t130);
    t151 = (uint32_t)(t150 ^ t127);
    t152 = (uint32_t)(t128 ^ t119);
    t153 = (uint32_t)(t152 & t130);
    t154 = (uint32_t)(t153 ^ t128);
    t155 = (uint32_t)((uint32_t)(a_0_i * t133));
    t156 = (uint32_t)(((uint64_t)a_0_i * t133) >> 32);
    t157 = (uint32_t)((uint32_t)(a_2_i * t133));
    t158 = (uint32_t)(((uint64_t)a_2_i * t133) >> 32);
    t159 = (uint32_t)((uint32_t)(a_4_i * t133));
    t160 = (uint32_t)(((uint64_t)a_4_i * t133) >> 32);
    t161 = (uint32_t)((uint32_t)(a_6_i * t133));
    t162 = (uint32_t)(((uint64_t)a_6_i * t133) >> 32);
    t172 = (uint32_t)((uint32_t)(a_1_i * t133));
    t173 = (uint32_t)(((uint64_t)a_1_i * t133) >> 32);
    t174 = (uint32_t)((uint32_t)(a_3_i * t133));
    t175 = (uint32_t)(((uint64_t)a_3_i * t133) >> 32);
    t176 = (uint32_t)((uint32_t)(a_5_i * t133));
    t177 = (uint32_t)(((uint64_t)a_5_i * t133) >> 32);
    t178 = (uint32_t)((uint32_t)(a_7_i * t133));
    t179 = (uint32_t)(((uint64_t)a_7_i * t133) >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * t136) + t157; t157 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t136) >> 32) + t158 + cf_; t158 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t136) + t159 + cf_; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t136) >> 32) + t160 + cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t136) + t161 + cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t136) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t136) + 0x0u + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t136) >> 32) + 0x0u + cf_; t164 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t136) + t172; t172 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t136) >> 32) + t173 + cf_; t173 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t136) + t174 + cf_; t174 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t136) >> 32) + t175 + cf_; t175 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t136) + t176 + cf_; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t136) >> 32) + t177 + cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t136) + t178 + cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t136) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t180 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t139) + t157; t157 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t139) >> 32) + t158 + cf_; t158 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t139) + t159 + cf_; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t139) >> 32) + t160 + cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t139) + t161 + cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t139) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t139) + t163 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t139) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t165 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t139) + t174; t174 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t139) >> 32) + t175 + cf_; t175 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t139) + t176 + cf_; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t139) >> 32) + t177 + cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t139) + t178 + cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t139) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t139) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t139) >> 32) + 0x0u + cf_; t181 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t142) + t159; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t142) >> 32) + t160 + cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t142) + t161 + cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t142) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t142) + t163 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t142) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t142) + t165 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t142) >> 32) + 0x0u + cf_; t166 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t142) + t174; t174 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t142) >> 32) + t175 + cf_; t175 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t142) + t176 + cf_; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t142) >> 32) + t177 + cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t142) + t178 + cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t142) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t142) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t142) >> 32) + t181 + cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t182 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t145) + t159; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t145) >> 32) + t160 + cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t145) + t161 + cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t145) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t145) + t163 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t145) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t145) + t165 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t145) >> 32) + t166 + cf_; t166 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t167 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t145) + t176; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t145) >> 32) + t177 + cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t145) + t178 + cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t145) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t145) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t145) >> 32) + t181 + cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t145) + t182 + cf_; t182 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t145) >> 32) + 0x0u + cf_; t183 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t148) + t161; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t148) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t148) + t163 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t148) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t148) + t165 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t148) >> 32) + t166 + cf_; t166 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t148) + t167 + cf_; t167 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t148) >> 32) + 0x0u + cf_; t168 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t148) + t176; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t148) >> 32) + t177 + cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t148) + t178 + cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t148) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t148) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t148) >> 32) + t181 + cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t148) + t182 + cf_; t182 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t148) >> 32) + t183 + cf_; t183 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t184 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t151) + t161; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t151) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t151) + t163 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t151) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t151) + t165 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t151) >> 32) + t166 + cf_; t166 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t151) + t167 + cf_; t167 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t151) >> 32) + t168 + cf_; t168 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t169 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t151) + t178; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t151) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t151) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t151) >> 32) + t181 + cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t151) + t182 + cf_; t182 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t151) >> 32) + t183 + cf_; t183 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t151) + t184 + cf_; t184 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t151) >> 32) + 0x0u + cf_; t185 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t154) + t163; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t154) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t154) + t165 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t154) >> 32) + t166 + cf_; t166 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t154) + t167 + cf_; t167 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t154) >> 32) + t168 + cf_; t168 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t154) + t169 + cf_; t169 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t154) >> 32) + 0x0u + cf_; t170 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t154) + t178; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t154) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t154) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t154) >> 32) + t181 + cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t154) + t182 + cf_; t182 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t154) >> 32) + t183 + cf_; t183 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t154) + t184 + cf_; t184 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t154) >> 32) + t185 + cf_; t185 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t186 = (uint32_t)w_;
    w_ = (uint64_t)t156 + t172; t187 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t157 + t173 + cf_; t188 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t158 + t174 + cf_; t189 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t159 + t175 + cf_; t190 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t160 + t176 + cf_; t191 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t161 + t177 + cf_; t192 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t162 + t178 + cf_; t193 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t163 + t179 + cf_; t194 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t164 + t180 + cf_; t195 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t165 + t181 + cf_; t196 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t166 + t182 + cf_; t197 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t167 + t183 + cf_; t198 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t168 + t184 + cf_; t199 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t169 + t185 + cf_; t200 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t170 + t186 + cf_; t201 = (uint32_t)w_;
    t202 = (uint32_t)((uint32_t)(t155 * 0xee00bc4fu));
    t203 = (uint32_t)(((uint64_t)t155 * 0xee00bc4fu) >> 32);
    t204 = (uint32_t)((uint32_t)(t188 * 0xee00bc4fu));
    t205 = (uint32_t)(((uint64_t)t188 * 0xee00bc4fu) >> 32);
    t206 = (uint32_t)((uint32_t)(t190 * 0xee00bc4fu));
    t207 = (uint32_t)(((uint64_t)t190 * 0xee00bc4fu) >> 32);
    t208 = (uint32_t)((uint32_t)(t192 * 0xee00bc4fu));
    t209 = (uint32_t)(((uint64_t)t192 * 0xee00bc4fu) >> 32);
    t212 = (uint32_t)((uint32_t)(t187 * 0xee00bc4fu));
    t213 = (uint32_t)(((uint64_t)t187 * 0xee00bc4fu) >> 32);
    t214 = (uint32_t)((uint32_t)(t189 * 0xee00bc4fu));
    t215 = (uint32_t)(((uint64_t)t189 * 0xee00bc4fu) >> 32);
    t216 = (uint32_t)((uint32_t)(t191 * 0xee00bc4fu));
    t217 = (uint32_t)(((uint64_t)t191 * 0xee00bc4fu) >> 32);
    w_ = (uint64_t)(uint32_t)(t187 * 0xccd1c8aau) + t204; t204 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t187 * 0xccd1c8aau) >> 32) + t205 + cf_; t205 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t189 * 0xccd1c8aau) + t206 + cf_; t206 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t189 * 0xccd1c8aau) >> 32) + t207 + cf_; t207 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t191 * 0xccd1c8aau) + t208 + cf_; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t191 * 0xccd1c8aau) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0xccd1c8aau) + t212; t212 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0xccd1c8aau) >> 32) + t213 + cf_; t213 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t188 * 0xccd1c8aau) + t214 + cf_; t214 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t188 * 0xccd1c8aau) >> 32) + t215 + cf_; t215 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t190 * 0xccd1c8aau) + t216 + cf_; t216 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t190 * 0xccd1c8aau) >> 32) + t217 + cf_; t217 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t218 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0x7d74d2e4u) + t204; t204 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0x7d74d2e4u) >> 32) + t205 + cf_; t205 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t188 * 0x7d74d2e4u) + t206 + cf_; t206 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t188 * 0x7d74d2e4u) >> 32) + t207 + cf_; t207 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t190 * 0x7d74d2e4u) + t208 + cf_; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t190 * 0x7d74d2e4u) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t210 + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t187 * 0x7d74d2e4u) + t214; t214 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t187 * 0x7d74d2e4u) >> 32) + t215 + cf_; t215 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t189 * 0x7d74d2e4u) + t216 + cf_; t216 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t189 * 0x7d74d2e4u) >> 32) + t217 + cf_; t217 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t218 + 0x0u + cf_; t218 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t187 * 0x48c94408u) + t206; t206 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t187 * 0x48c94408u) >> 32) + t207 + cf_; t207 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t189 * 0x48c94408u) + t208 + cf_; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t189 * 0x48c94408u) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t210 + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0x48c94408u) + t214; t214 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0x48c94408u) >> 32) + t215 + cf_; t215 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t188 * 0x48c94408u) + t216 + cf_; t216 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t188 * 0x48c94408u) >> 32) + t217 + cf_; t217 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t218 + 0x0u + cf_; t218 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0xc588c6f6u) + t206; t206 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0xc588c6f6u) >> 32) + t207 + cf_; t207 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t188 * 0xc588c6f6u) + t208 + cf_; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t188 * 0xc588c6f6u) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t210 + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t187 * 0xc588c6f6u) + t216; t216 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t187 * 0xc588c6f6u) >> 32) + t217 + cf_; t217 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t218 + 0x0u + cf_; t218 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t187 * 0x50fe77ecu) + t208; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t187 * 0x50fe77ecu) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t210 + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0x50fe77ecu) + t216; t216 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0x50fe77ecu) >> 32) + t217 + cf_; t217 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t218 + 0x0u + cf_; t218 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0xa9d6281cu) + t208; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0xa9d6281cu) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t210 + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)t203 + t212; t220 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t204 + t213 + cf_; t221 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t205 + t214 + cf_; t222 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t206 + t215 + cf_; t223 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t207 + t216 + cf_; t224 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t208 + t217 + cf_; t225 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t209 + t218 + cf_; t226 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t193 * 0xee00bc4fu) + t226; t227 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t192 * 0xccd1c8aau) + t227; t228 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t191 * 0x7d74d2e4u) + t228; t229 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t190 * 0x48c94408u) + t229; t230 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t189 * 0xc588c6f6u) + t230; t231 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t188 * 0x50fe77ecu) + t231; t232 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t187 * 0xa9d6281cu) + t232; t233 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0x60d06633u) + t233; t234 = (uint32_t)w_;
    t235 = (uint32_t)((uint32_t)(t202 * 0xfc632551u));
    t236 = (uint32_t)(((uint64_t)t202 * 0xfc632551u) >> 32);
    t237 = (uint32_t)((uint32_t)(t221 * 0xfc632551u));
    t238 = (uint32_t)(((uint64_t)t221 * 0xfc632551u) >> 32);
    t239 = (uint32_t)((uint32_t)(t223 * 0xfc632551u));
    t240 = (uint32_t)(((uint64_t)t223 * 0xfc632551u) >> 32);
    t241 = (uint32_t)((uint32_t)(t225 * 0xfc632551u));
    t242 = (uint32_t)(((uint64_t)t225 * 0xfc632551u) >> 32);
    t252 = (uint32_t)((uint32_t)(t220 * 0xfc632551u));
    t253 = (uint32_t)(((uint64_t)t220 * 0xfc632551u) >> 32);
    t254 = (uint32_t)((uint32_t)(t222 * 0xfc632551u));
    t255 = (uint32_t)(((uint64_t)t222 * 0xfc632551u) >> 32);
    t256 = (uint32_t)((uint32_t)(t224 * 0xfc632551u));
    t257 = (uint32_t)(((uint64_t)t224 * 0xfc632551u) >> 32);
    t258 = (uint32_t)((uint32_t)(t234 * 0xfc632551u));
    t259 = (uint32_t)(((uint64_t)t234 * 0xfc632551u) >> 32);
    w_ = (uint64_t)(uint32_t)(t220 * 0xf3b9cac2u) + t237; t237 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xf3b9cac2u) >> 32) + t238 + cf_; t238 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xf3b9cac2u) + t239 + cf_; t239 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xf3b9cac2u) >> 32) + t240 + cf_; t240 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xf3b9cac2u) + t241 + cf_; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xf3b9cac2u) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xf3b9cac2u) + 0x0u + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xf3b9cac2u) >> 32) + 0x0u + cf_; t244 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xf3b9cac2u) + t252; t252 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xf3b9cac2u) >> 32) + t253 + cf_; t253 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xf3b9cac2u) + t254 + cf_; t254 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xf3b9cac2u) >> 32) + t255 + cf_; t255 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xf3b9cac2u) + t256 + cf_; t256 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xf3b9cac2u) >> 32) + t257 + cf_; t257 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xf3b9cac2u) + t258 + cf_; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xf3b9cac2u) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t260 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xa7179e84u) + t237; t237 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xa7179e84u) >> 32) + t238 + cf_; t238 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xa7179e84u) + t239 + cf_; t239 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xa7179e84u) >> 32) + t240 + cf_; t240 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xa7179e84u) + t241 + cf_; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xa7179e84u) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xa7179e84u) + t243 + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xa7179e84u) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t245 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0xa7179e84u) + t254; t254 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xa7179e84u) >> 32) + t255 + cf_; t255 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xa7179e84u) + t256 + cf_; t256 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xa7179e84u) >> 32) + t257 + cf_; t257 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xa7179e84u) + t258 + cf_; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xa7179e84u) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xa7179e84u) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xa7179e84u) >> 32) + 0x0u + cf_; t261 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0xbce6faadu) + t239; t239 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xbce6faadu) >> 32) + t240 + cf_; t240 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xbce6faadu) + t241 + cf_; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xbce6faadu) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xbce6faadu) + t243 + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xbce6faadu) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xbce6faadu) + t245 + cf_; t245 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xbce6faadu) >> 32) + 0x0u + cf_; t246 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xbce6faadu) + t254; t254 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xbce6faadu) >> 32) + t255 + cf_; t255 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xbce6faadu) + t256 + cf_; t256 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xbce6faadu) >> 32) + t257 + cf_; t257 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xbce6faadu) + t258 + cf_; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xbce6faadu) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xbce6faadu) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xbce6faadu) >> 32) + t261 + cf_; t261 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t262 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xffffffffu) + t239; t239 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xffffffffu) >> 32) + t240 + cf_; t240 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xffffffffu) + t241 + cf_; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xffffffffu) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xffffffffu) + t243 + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xffffffffu) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xffffffffu) + t245 + cf_; t245 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xffffffffu) >> 32) + t246 + cf_; t246 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t247 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0xffffffffu) + t256; t256 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xffffffffu) >> 32) + t257 + cf_; t257 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xffffffffu) + t258 + cf_; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xffffffffu) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xffffffffu) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xffffffffu) >> 32) + t261 + cf_; t261 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xffffffffu) + t262 + cf_; t262 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xffffffffu) >> 32) + 0x0u + cf_; t263 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0xffffffffu) + t241; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xffffffffu) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xffffffffu) + t243 + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xffffffffu) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xffffffffu) + t245 + cf_; t245 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xffffffffu) >> 32) + t246 + cf_; t246 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xffffffffu) + t247 + cf_; t247 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xffffffffu) >> 32) + 0x0u + cf_; t248 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xffffffffu) + t256; t256 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xffffffffu) >> 32) + t257 + cf_; t257 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xffffffffu) + t258 + cf_; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xffffffffu) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xffffffffu) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xffffffffu) >> 32) + t261 + cf_; t261 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xffffffffu) + t262 + cf_; t262 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xffffffffu) >> 32) + t263 + cf_; t263 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t264 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0x0u) + t241; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0x0u) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0x0u) + t243 + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0x0u) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0x0u) + t245 + cf_; t245 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0x0u) >> 32) + t246 + cf_; t246 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0x0u) + t247 + cf_; t247 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0x0u) >> 32) + t248 + cf_; t248 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t249 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0x0u) + t258; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0x0u) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0x0u) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0x0u) >> 32) + t261 + cf_; t261 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0x0u) + t262 + cf_; t262 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0x0u) >> 32) + t263 + cf_; t263 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0x0u) + t264 + cf_; t264 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0x0u) >> 32) + 0x0u + cf_; t265 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0xffffffffu) + t243; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xffffffffu) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xffffffffu) + t245 + cf_; t245 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xffffffffu) >> 32) + t246 + cf_; t246 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xffffffffu) + t247 + cf_; t247 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xffffffffu) >> 32) + t248 + cf_; t248 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xffffffffu) + t249 + cf_; t249 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xffffffffu) >> 32) + 0x0u + cf_; t250 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xffffffffu) + t258; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xffffffffu) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xffffffffu) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xffffffffu) >> 32) + t261 + cf_; t261 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xffffffffu) + t262 + cf_; t262 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xffffffffu) >> 32) + t263 + cf_; t263 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xffffffffu) + t264 + cf_; t264 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xffffffffu) >> 32) + t265 + cf_; t265 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t266 = (uint32_t)w_;
    w_ = (uint64_t)t236 + t252; t267 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t237 + t253 + cf_; t268 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t238 + t254 + cf_; t269 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t239 + t255 + cf_; t270 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t240 + t256 + cf_; t271 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t241 + t257 + cf_; t272 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t242 + t258 + cf_; t273 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t243 + t259 + cf_; t274 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t244 + t260 + cf_; t275 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t245 + t261 + cf_; t276 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t246 + t262 + cf_; t277 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t247 + t263 + cf_; t278 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t248 + t264 + cf_; t279 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t249 + t265 + cf_; t280 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t250 + t266 + cf_; t281 = (uint32_t)w_;
    w_ = (uint64_t)t155 + t235; t282 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t187 + t267 + cf_; t283 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t188 + t268 + cf_; t284 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t189 + t269 + cf_; t285 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t190 + t270 + cf_; t286 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t191 + t271 + cf_; t287 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t192 + t272 + cf_; t288 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t193 + t273 + cf_; t289 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t194 + t274 + cf_; t290 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t195 + t275 + cf_; t291 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t196 + t276 + cf_; t292 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t197 + t277 + cf_; t293 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t198 + t278 + cf_; t294 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t199 + t279 + cf_; t295 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t200 + t280 + cf_; t296 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t201 + t281 + cf_; t297 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t298 = (uint32_t)w_;
    w_ = (uint64_t)t290 - 0xfc632551u; t299 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t291 - 0xf3b9cac2u - cf_; t300 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t292 - 0xa7179e84u - cf_; t301 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t293 - 0xbce6faadu - cf_; t302 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t294 - 0xffffffffu - cf_; t303 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t295 - 0xffffffffu - cf_; t304 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t296 - 0x0u - cf_; t305 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t297 - 0xffffffffu - cf_; t306 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t298 - 0x0u - cf_; t307 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t308 = (uint32_t)w_;
    t309 = (uint32_t)(t299 ^ t290);
    t310 = (uint32_t)(t309 & t308);
    t311 = (uint32_t)(t310 ^ t299);
    t312 = (uint32_t)(t300 ^ t291);
    t313 = (uint32_t)(t312 & t308);
    t314 = (uint32_t)(t313 ^ t300);
    t315 = (uint32_t)(t301 ^ t292);
    t316 = (uint32_t)(t315 & t308);
    t317 = (uint32_t)(t316 ^ t301);
    t318 = (uint32_t)(t302 ^ t293);
    t319 = (uint32_t)(t318 & t308);
    t320 = (uint32_t)(t319 ^ t302);
    t321 = (uint32_t)(t303 ^ t294);
    t322 = (uint32_t)(t321 & t308);
    t323 = (uint32_t)(t322 ^ t303);
    t324 = (uint32_t)(t304 ^ t295);
    t325 = (uint32_t)(t324 & t308);
    t326 = (uint32_t)(t325 ^ t304);
    t327 = (uint32_t)(t305 ^ t296);
    t328 = (uint32_t)(t327 & t308);
    t329 = (uint32_t)(t328 ^ t305);
    t330 = (uint32_t)(t306 ^ t297);
    t331 = (uint32_t)(t330 & t308);
    t332 = (uint32_t)(t331 ^ t306);
    r[0] = t311;
    r[1] = t314;
    r[2] = t317;
    r[3] = t320;
    r[4] = t323;
    r[5] = t326;
    r[6] = t329;
    r[7] = t332;
#endif
  }

  // r = a*b + c, small integer b: modmli + modadd fused (rfc7748.c:209,212)
  static MAB_DEV void mla(uint32_t (&r)[8], const uint32_t (&a)[8], uint32_t b, const uint32_t (&c)[8]) {
#ifndef MAB_HOSTSIM
    asm("{\n\t"
        ".reg .u32 t<376>;\n\t"
        "mul.lo.u32 t0, 0xbe79eea2, %24;\n\t"
        "mul.hi.u32 t8, 0xbe79eea2, %24;\n\t"
        "mul.lo.u32 t1, 0x83244c95, %24;\n\t"
        "mul.hi.u32 t9, 0x83244c95, %24;\n\t"
        "mul.lo.u32 t2, 0x49bd6fa6, %24;\n\t"
        "mul.hi.u32 t10, 0x49bd6fa6, %24;\n\t"
        "mul.lo.u32 t3, 0x4699799c, %24;\n\t"
        "mul.hi.u32 t11, 0x4699799c, %24;\n\t"
        "mul.lo.u32 t4, 0x2b6bec59, %24;\n\t"
        "mul.hi.u32 t12, 0x2b6bec59, %24;\n\t"
        "mul.lo.u32 t5, 0x2845b239, %24;\n\t"
        "mul.hi.u32 t13, 0x2845b239, %24;\n\t"
        "mul.lo.u32 t6, 0xf3d95620, %24;\n\t"
        "mul.hi.u32 t14, 0xf3d95620, %24;\n\t"
        "mul.lo.u32 t7, 0x66e12d94, %24;\n\t"
        "mul.hi.u32 t15, 0x66e12d94, %24;\n\t"
        "add.cc.u32 t16, t1, t8;\n\t"
        "addc.cc.u32 t17, t2, t9;\n\t"
        "addc.cc.u32 t18, t3, t10;\n\t"
        "addc.cc.u32 t19, t4, t11;\n\t"
        "addc.cc.u32 t20, t5, t12;\n\t"
        "addc.cc.u32 t21, t6, t13;\n\t"
        "addc.cc.u32 t22, t7, t14;\n\t"
        "addc.u32 t23, t15, 0x0;\n\t"
        "mul.lo.u32 t24, t0, 0xee00bc4f;\n\t"
        "mul.hi.u32 t25, t0, 0xee00bc4f;\n\t"
        "mul.lo.u32 t26, t17, 0xee00bc4f;\n\t"
        "mul.hi.u32 t27, t17, 0xee00bc4f;\n\t"
        "mul.lo.u32 t28, t19, 0xee00bc4f;\n\t"
        "mul.hi.u32 t29, t19, 0xee00bc4f;\n\t"
        "mul.lo.u32 t30, t21, 0xee00bc4f;\n\t"
        "mul.hi.u32 t31, t21, 0xee00bc4f;\n\t"
        "mul.lo.u32 t34, t16, 0xee00bc4f;\n\t"
        "mul.hi.u32 t35, t16, 0xee00bc4f;\n\t"
        "mul.lo.u32 t36, t18, 0xee00bc4f;\n\t"
        "mul.hi.u32 t37, t18, 0xee00bc4f;\n\t"
        "mul.lo.u32 t38, t20, 0xee00bc4f;\n\t"
        "mul.hi.u32 t39, t20, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t26, t16, 0xccd1c8aa, t26;\n\t"
        "madc.hi.cc.u32 t27, t16, 0xccd1c8aa, t27;\n\t"
        "madc.lo.cc.u32 t28, t18, 0xccd1c8aa, t28;\n\t"
        "madc.hi.cc.u32 t29, t18, 0xccd1c8aa, t29;\n\t"
        "madc.lo.cc.u32 t30, t20, 0xccd1c8aa, t30;\n\t"
        "madc.hi.cc.u32 t31, t20, 0xccd1c8aa, t31;\n\t"
        "addc.u32 t32, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t34, t0, 0xccd1c8aa, t34;\n\t"
        "madc.hi.cc.u32 t35, t0, 0xccd1c8aa, t35;\n\t"
        "madc.lo.cc.u32 t36, t17, 0xccd1c8aa, t36;\n\t"
        "madc.hi.cc.u32 t37, t17, 0xccd1c8aa, t37;\n\t"
        "madc.lo.cc.u32 t38, t19, 0xccd1c8aa, t38;\n\t"
        "madc.hi.cc.u32 t39, t19, 0xccd1c8aa, t39;\n\t"
        "addc.u32 t40, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t26, t0, 0x7d74d2e4, t26;\n\t"
        "madc.hi.cc.u32 t27, t0, 0x7d74d2e4, t27;\n\t"
        "madc.lo.cc.u32 t28, t17, 0x7d74d2e4, t28;\n\t"
        "madc.hi.cc.u32 t29, t17, 0x7d74d2e4, t29;\n\t"
        "madc.lo.cc.u32 t30, t19, 0x7d74d2e4, t30;\n\t"
        "madc.hi.cc.u32 t31, t19, 0x7d74d2e4, t31;\n\t"
        "addc.u32 t32, t32, 0x0;\n\t"
        "mad.lo.cc.u32 t36, t16, 0x7d74d2e4, t36;\n\t"
        "madc.hi.cc.u32 t37, t16, 0x7d74d2e4, t37;\n\t"
        "madc.lo.cc.u32 t38, t18, 0x7d74d2e4, t38;\n\t"
        "madc.hi.cc.u32 t39, t18, 0x7d74d2e4, t39;\n\t"
        "addc.u32 t40, t40, 0x0;\n\t"
        "mad.lo.cc.u32 t28, t16, 0x48c94408, t28;\n\t"
        "madc.hi.cc.u32 t29, t16, 0x48c94408, t29;\n\t"
        "madc.lo.cc.u32 t30, t18, 0x48c94408, t30;\n\t"
        "madc.hi.cc.u32 t31, t18, 0x48c94408, t31;\n\t"
        "addc.u32 t32, t32, 0x0;\n\t"
        "mad.lo.cc.u32 t36, t0, 0x48c94408, t36;\n\t"
        "madc.hi.cc.u32 t37, t0, 0x48c94408, t37;\n\t"
        "madc.lo.cc.u32 t38, t17, 0x48c94408, t38;\n\t"
        "madc.hi.cc.u32 t39, t17, 0x48c94408, t39;\n\t"
        "addc.u32 t40, t40, 0x0;\n\t"
        "mad.lo.cc.u32 t28, t0, 0xc588c6f6, t28;\n\t"
        "madc.hi.cc.u32 t29, t0, 0xc588c6f6, t29;\n\t"
        "madc.lo.cc.u32 t30, t17, 0xc588c6f6, t30;\n\t"
        "madc.hi.cc.u32 t31, t17, 0xc588c6f6, t31;\n\t"
        "addc.u32 t32, t32, 0x0;\n\t"
        "mad.lo.cc.u32 t38, t16, 0xc588c6f6, t38;\n\t"
        "madc.hi.cc.u32 t39, t16, 0xc588c6f6, t39;\n\t"
        "addc.u32 t40, t40, 0x0;\n\t"
        "mad.lo.cc.u32 t30, t16, 0x50fe77ec, t30;\n\t"
        "madc.hi.cc.u32 t31, t16, 0x50fe77ec, t31;\n\t"
        "addc.u32 t32, t32, 0x0;\n\t"
        "mad.lo.cc.u32 t38, t0, 0x50fe77ec, t38;\n\t"
        "madc.hi.cc.u32 t39, t0, 0x50fe77ec, t39;\n\t"
        "addc.u32 t40, t40, 0x0;\n\t"
        "mad.lo.cc.u32 t30, t0, 0xa9d6281c, t30;\n\t"
        "madc.hi.cc.u32 t31, t0, 0xa9d6281c, t31;\n\t"
        "addc.u32 t32, t32, 0x0;\n\t"
        "add.cc.u32 t42, t25, t34;\n\t"
        "addc.cc.u32 t43, t26, t35;\n\t"
        "addc.cc.u32 t44, t27, t36;\n\t"
        "addc.cc.u32 t45, t28, t37;\n\t"
        "addc.cc.u32 t46, t29, t38;\n\t"
        "addc.cc.u32 t47, t30, t39;\n\t"
        "addc.u32 t48, t31, t40;\n\t"
        "mad.lo.u32 t49, t22, 0xee00bc4f, t48;\n\t"
        "mad.lo.u32 t50, t21, 0xccd1c8aa, t49;\n\t"
        "mad.lo.u32 t51, t20, 0x7d74d2e4, t50;\n\t"
        "mad.lo.u32 t52, t19, 0x48c94408, t51;\n\t"
        "mad.lo.u32 t53, t18, 0xc588c6f6, t52;\n\t"
        "mad.lo.u32 t54, t17, 0x50fe77ec, t53;\n\t"
        "mad.lo.u32 t55, t16, 0xa9d6281c, t54;\n\t"
        "mad.lo.u32 t56, t0, 0x60d06633, t55;\n\t"
        "mul.lo.u32 t57, t24, 0xfc632551;\n\t"
        "mul.hi.u32 t58, t24, 0xfc632551;\n\t"
        "mul.lo.u32 t59, t43, 0xfc632551;\n\t"
        "mul.hi.u32 t60, t43, 0xfc632551;\n\t"
        "mul.lo.u32 t61, t45, 0xfc632551;\n\t"
        "mul.hi.u32 t62, t45, 0xfc632551;\n\t"
        "mul.lo.u32 t63, t47, 0xfc632551;\n\t"
        "mul.hi.u32 t64, t47, 0xfc632551;\n\t"
        "mul.lo.u32 t74, t42, 0xfc632551;\n\t"
        "mul.hi.u32 t75, t42, 0xfc632551;\n\t"
        "mul.lo.u32 t76, t44, 0xfc632551;\n\t"
        "mul.hi.u32 t77, t44, 0xfc632551;\n\t"
        "mul.lo.u32 t78, t46, 0xfc632551;\n\t"
        "mul.hi.u32 t79, t46, 0xfc632551;\n\t"
        "mul.lo.u32 t80, t56, 0xfc632551;\n\t"
        "mul.hi.u32 t81, t56, 0xfc632551;\n\t"
        "mad.lo.cc.u32 t59, t42, 0xf3b9cac2, t59;\n\t"
        "madc.hi.cc.u32 t60, t42, 0xf3b9cac2, t60;\n\t"
        "madc.lo.cc.u32 t61, t44, 0xf3b9cac2, t61;\n\t"
        "madc.hi.cc.u32 t62, t44, 0xf3b9cac2, t62;\n\t"
        "madc.lo.cc.u32 t63, t46, 0xf3b9cac2, t63;\n\t"
        "madc.hi.cc.u32 t64, t46, 0xf3b9cac2, t64;\n\t"
        "madc.lo.cc.u32 t65, t56, 0xf3b9cac2, 0x0;\n\t"
        "madc.hi.u32 t66, t56, 0xf3b9cac2, 0x0;\n\t"
        "mad.lo.cc.u32 t74, t24, 0xf3b9cac2, t74;\n\t"
        "madc.hi.cc.u32 t75, t24, 0xf3b9cac2, t75;\n\t"
        "madc.lo.cc.u32 t76, t43, 0xf3b9cac2, t76;\n\t"
        "madc.hi.cc.u32 t77, t43, 0xf3b9cac2, t77;\n\t"
        "madc.lo.cc.u32 t78, t45, 0xf3b9cac2, t78;\n\t"
        "madc.hi.cc.u32 t79, t45, 0xf3b9cac2, t79;\n\t"
        "madc.lo.cc.u32 t80, t47, 0xf3b9cac2, t80;\n\t"
        "madc.hi.cc.u32 t81, t47, 0xf3b9cac2, t81;\n\t"
        "addc.u32 t82, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t59, t24, 0xa7179e84, t59;\n\t"
        "madc.hi.cc.u32 t60, t24, 0xa7179e84, t60;\n\t"
        "madc.lo.cc.u32 t61, t43, 0xa7179e84, t61;\n\t"
        "madc.hi.cc.u32 t62, t43, 0xa7179e84, t62;\n\t"
        "madc.lo.cc.u32 t63, t45, 0xa7179e84, t63;\n\t"
        "madc.hi.cc.u32 t64, t45, 0xa7179e84, t64;\n\t"
        "madc.lo.cc.u32 t65, t47, 0xa7179e84, t65;\n\t"
        "madc.hi.cc.u32 t66, t47, 0xa7179e84, t66;\n\t"
        "addc.u32 t67, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t76, t42, 0xa7179e84, t76;\n\t"
        "madc.hi.cc.u32 t77, t42, 0xa7179e84, t77;\n\t"
        "madc.lo.cc.u32 t78, t44, 0xa7179e84, t78;\n\t"
        "madc.hi.cc.u32 t79, t44, 0xa7179e84, t79;\n\t"
        "madc.lo.cc.u32 t80, t46, 0xa7179e84, t80;\n\t"
        "madc.hi.cc.u32 t81, t46, 0xa7179e84, t81;\n\t"
        "madc.lo.cc.u32 t82, t56, 0xa7179e84, t82;\n\t"
        "madc.hi.u32 t83, t56, 0xa7179e84, 0x0;\n\t"
        "mad.lo.cc.u32 t61, t42, 0xbce6faad, t61;\n\t"
        "madc.hi.cc.u32 t62, t42, 0xbce6faad, t62;\n\t"
        "madc.lo.cc.u32 t63, t44, 0xbce6faad, t63;\n\t"
        "madc.hi.cc.u32 t64, t44, 0xbce6faad, t64;\n\t"
        "madc.lo.cc.u32 t65, t46, 0xbce6faad, t65;\n\t"
        "madc.hi.cc.u32 t66, t46, 0xbce6faad, t66;\n\t"
        "madc.lo.cc.u32 t67, t56, 0xbce6faad, t67;\n\t"
        "madc.hi.u32 t68, t56, 0xbce6faad, 0x0;\n\t"
        "mad.lo.cc.u32 t76, t24, 0xbce6faad, t76;\n\t"
        "madc.hi.cc.u32 t77, t24, 0xbce6faad, t77;\n\t"
        "madc.lo.cc.u32 t78, t43, 0xbce6faad, t78;\n\t"
        "madc.hi.cc.u32 t79, t43, 0xbce6faad, t79;\n\t"
        "madc.lo.cc.u32 t80, t45, 0xbce6faad, t80;\n\t"
        "madc.hi.cc.u32 t81, t45, 0xbce6faad, t81;\n\t"
        "madc.lo.cc.u32 t82, t47, 0xbce6faad, t82;\n\t"
        "madc.hi.cc.u32 t83, t47, 0xbce6faad, t83;\n\t"
        "addc.u32 t84, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t61, t24, 0xffffffff, t61;\n\t"
        "madc.hi.cc.u32 t62, t24, 0xffffffff, t62;\n\t"
        "madc.lo.cc.u32 t63, t43, 0xffffffff, t63;\n\t"
        "madc.hi.cc.u32 t64, t43, 0xffffffff, t64;\n\t"
        "madc.lo.cc.u32 t65, t45, 0xffffffff, t65;\n\t"
        "madc.hi.cc.u32 t66, t45, 0xffffffff, t66;\n\t"
        "madc.lo.cc.u32 t67, t47, 0xffffffff, t67;\n\t"
        "madc.hi.cc.u32 t68, t47, 0xffffffff, t68;\n\t"
        "addc.u32 t69, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t78, t42, 0xffffffff, t78;\n\t"
        "madc.hi.cc.u32 t79, t42, 0xffffffff, t79;\n\t"
        "madc.lo.cc.u32 t80, t44, 0xffffffff, t80;\n\t"
        "madc.hi.cc.u32 t81, t44, 0xffffffff, t81;\n\t"
        "madc.lo.cc.u32 t82, t46, 0xffffffff, t82;\n\t"
        "madc.hi.cc.u32 t83, t46, 0xffffffff, t83;\n\t"
        "madc.lo.cc.u32 t84, t56, 0xffffffff, t84;\n\t"
        "madc.hi.u32 t85, t56, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t63, t42, 0xffffffff, t63;\n\t"
        "madc.hi.cc.u32 t64, t42, 0xffffffff, t64;\n\t"
        "madc.lo.cc.u32 t65, t44, 0xffffffff, t65;\n\t"
        "madc.hi.cc.u32 t66, t44, 0xffffffff, t66;\n\t"
        "madc.lo.cc.u32 t67, t46, 0xffffffff, t67;\n\t"
        "madc.hi.cc.u32 t68, t46, 0xffffffff, t68;\n\t"
        "madc.lo.cc.u32 t69, t56, 0xffffffff, t69;\n\t"
        "madc.hi.u32 t70, t56, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t78, t24, 0xffffffff, t78;\n\t"
        "madc.hi.cc.u32 t79, t24, 0xffffffff, t79;\n\t"
        "madc.lo.cc.u32 t80, t43, 0xffffffff, t80;\n\t"
        "madc.hi.cc.u32 t81, t43, 0xffffffff, t81;\n\t"
        "madc.lo.cc.u32 t82, t45, 0xffffffff, t82;\n\t"
        "madc.hi.cc.u32 t83, t45, 0xffffffff, t83;\n\t"
        "madc.lo.cc.u32 t84, t47, 0xffffffff, t84;\n\t"
        "madc.hi.cc.u32 t85, t47, 0xffffffff, t85;\n\t"
        "addc.u32 t86, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t63, t24, 0x0, t63;\n\t"
        "madc.hi.cc.u32 t64, t24, 0x0, t64;\n\t"
        "madc.lo.cc.u32 t65, t43, 0x0, t65;\n\t"
        "madc.hi.cc.u32 t66, t43, 0x0, t66;\n\t"
        "madc.lo.cc.u32 t67, t45, 0x0, t67;\n\t"
        "madc.hi.cc.u32 t68, t45, 0x0, t68;\n\t"
        "madc.lo.cc.u32 t69, t47, 0x0, t69;\n\t"
        "madc.hi.cc.u32 t70, t47, 0x0, t70;\n\t"
        "addc.u32 t71, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t80, t42, 0x0, t80;\n\t"
        "madc.hi.cc.u32 t81, t42, 0x0, t81;\n\t"
        "madc.lo.cc.u32 t82, t44, 0x0, t82;\n\t"
        "madc.hi.cc.u32 t83, t44, 0x0, t83;\n\t"
        "madc.lo.cc.u32 t84, t46, 0x0, t84;\n\t"
        "madc.hi.cc.u32 t85, t46, 0x0, t85;\n\t"
        "madc.lo.cc.u32 t86, t56, 0x0, t86;\n\t"
        "madc.hi.u32 t87, t56, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t65, t42, 0xffffffff, t65;\n\t"
        "madc.hi.cc.u32 t66, t42, 0xffffffff, t66;\n\t"
        "madc.lo.cc.u32 t67, t44, 0xffffffff, t67;\n\t"
        "madc.hi.cc.u32 t68, t44, 0xffffffff, t68;\n\t"
        "madc.lo.cc.u32 t69, t46, 0xffffffff, t69;\n\t"
        "madc.hi.cc.u32 t70, t46, 0xffffffff, t70;\n\t"
        "madc.lo.cc.u32 t71, t56, 0xffffffff, t71;\n\t"
        "madc.hi.u32 t72, t56, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t80, t24, 0xffffffff, t80;\n\t"
        "madc.hi.cc.u32 t81, t24, 0xffffffff, t81;\n\t"
        "madc.lo.cc.u32 t82, t43, 0xffffffff, t82;\n\t"
        "madc.hi.cc.u32 t83, t43, 0xffffffff, t83;\n\t"
        "madc.lo.cc.u32 t84, t45, 0xffffffff, t84;\n\t"
        "madc.hi.cc.u32 t85, t45, 0xffffffff, t85;\n\t"
        "madc.lo.cc.u32 t86, t47, 0xffffffff, t86;\n\t"
        "madc.hi.cc.u32 t87, t47, 0xffffffff, t87;\n\t"
        "addc.u32 t88, 0x0, 0x0;\n\t"
        "add.cc.u32 t89, t58, t74;\n\t"
        "addc.cc.u32 t90, t59, t75;\n\t"
        "addc.cc.u32 t91, t60, t76;\n\t"
        "addc.cc.u32 t92, t61, t77;\n\t"
        "addc.cc.u32 t93, t62, t78;\n\t"
        "addc.cc.u32 t94, t63, t79;\n\t"
        "addc.cc.u32 t95, t64, t80;\n\t"
        "addc.cc.u32 t96, t65, t81;\n\t"
        "addc.cc.u32 t97, t66, t82;\n\t"
        "addc.cc.u32 t98, t67, t83;\n\t"
        "addc.cc.u32 t99, t68, t84;\n\t"
        "addc.cc.u32 t100, t69, t85;\n\t"
        "addc.cc.u32 t101, t70, t86;\n\t"
        "addc.cc.u32 t102, t71, t87;\n\t"
        "addc.u32 t103, t72, t88;\n\t"
        "add.cc.u32 t104, t0, t57;\n\t"
        "addc.cc.u32 t105, t16, t89;\n\t"
        "addc.cc.u32 t106, t17, t90;\n\t"
        "addc.cc.u32 t107, t18, t91;\n\t"
        "addc.cc.u32 t108, t19, t92;\n\t"
        "addc.cc.u32 t109, t20, t93;\n\t"
        "addc.cc.u32 t110, t21, t94;\n\t"
        "addc.cc.u32 t111, t22, t95;\n\t"
        "addc.cc.u32 t112, t23, t96;\n\t"
        "addc.cc.u32 t113, 0x0, t97;\n\t"
        "addc.cc.u32 t114, 0x0, t98;\n\t"
        "addc.cc.u32 t115, 0x0, t99;\n\t"
        "addc.cc.u32 t116, 0x0, t100;\n\t"
        "addc.cc.u32 t117, 0x0, t101;\n\t"
        "addc.cc.u32 t118, 0x0, t102;\n\t"
        "addc.cc.u32 t119, 0x0, t103;\n\t"
        "addc.u32 t120, 0x0, 0x0;\n\t"
        "sub.cc.u32 t121, t112, 0xfc632551;\n\t"
        "subc.cc.u32 t122, t113, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t123, t114, 0xa7179e84;\n\t"
        "subc.cc.u32 t124, t115, 0xbce6faad;\n\t"
        "subc.cc.u32 t125, t116, 0xffffffff;\n\t"
        "subc.cc.u32 t126, t117, 0xffffffff;\n\t"
        "subc.cc.u32 t127, t118, 0x0;\n\t"
        "subc.cc.u32 t128, t119, 0xffffffff;\n\t"
        "subc.cc.u32 t129, t120, 0x0;\n\t"
        "subc.u32 t130, 0x0, 0x0;\n\t"
        "xor.b32 t131, t121, t112;\n\t"
        "and.b32 t132, t131, t130;\n\t"
        "xor.b32 t133, t132, t121;\n\t"
        "xor.b32 t134, t122, t113;\n\t"
        "and.b32 t135, t134, t130;\n\t"
        "xor.b32 t136, t135, t122;\n\t"
        "xor.b32 t137, t123, t114;\n\t"
        "and.b32 t138, t137, t130;\n\t"
        "xor.b32 t139, t138, t123;\n\t"
        "xor.b32 t140, t124, t115;\n\t"
        "and.b32 t141, t140, t130;\n\t"
        "xor.b32 t142, t141, t124;\n\t"
        "xor.b32 t143, t125, t116;\n\t"
        "and.b32 t144, t143, t130;\n\t"
        "xor.b32 t145, t144, t125;\n\t"
        "xor.b32 t146, t126, t117;\n\t"
        "and.b32 t147, t146, t130;\n\t"
        "xor.b32 t148, t147, t126;\n\t"
        "xor.b32 t149, t127, t118;\n\t"
        "and.b32 t150, t149, t130;\n\t"
        "xor.b32 t151, t150, t127;\n\t"
        "xor.b32 t152, t128, t119;\n\t"
        "and.b32 t153, t152, t130;\n\t"
        "xor.b32 t154, t153, t128;\n\t"
        "mul.lo.u32 t155, %8, t133;\n\t"
        "mul.hi.u32 t156, %8, t133;\n\t"
        "mul.lo.u32 t157, %10, t133;\n\t"
        "mul.hi.u32 t158, %10, t133;\n\t"
        "mul.lo.u32 t159, %12, t133;\n\t"
        "mul.hi.u32 t160, %12, t133;\n\t"
        "mul.lo.u32 t161, %14, t133;\n\t"
        "mul.hi.u32 t162, %14, t133;\n\t"
        "mul.lo.u32 t172, %9, t133;\n\t"
        "mul.hi.u32 t173, %9, t133;\n\t"
        "mul.lo.u32 t174, %11, t133;\n\t"
        "mul.hi.u32 t175, %11, t133;\n\t"
        "mul.lo.u32 t176, %13, t133;\n\t"
        "mul.hi.u32 t177, %13, t133;\n\t"
        "mul.lo.u32 t178, %15, t133;\n\t"
        "mul.hi.u32 t179, %15, t133;\n\t"
        "mad.lo.cc.u32 t157, %9, t136, t157;\n\t"
        "madc.hi.cc.u32 t158, %9, t136, t158;\n\t"
        "madc.lo.cc.u32 t159, %11, t136, t159;\n\t"
        "madc.hi.cc.u32 t160, %11, t136, t160;\n\t"
        "madc.lo.cc.u32 t161, %13, t136, t161;\n\t"
        "madc.hi.cc.u32 t162, %13, t136, t162;\n\t"
        "madc.lo.cc.u32 t163, %15, t136, 0x0;\n\t"
        "madc.hi.u32 t164, %15, t136, 0x0;\n\t"
        "mad.lo.cc.u32 t172, %8, t136, t172;\n\t"
        "madc.hi.cc.u32 t173, %8, t136, t173;\n\t"
        "madc.lo.cc.u32 t174, %10, t136, t174;\n\t"
        "madc.hi.cc.u32 t175, %10, t136, t175;\n\t"
        "madc.lo.cc.u32 t176, %12, t136, t176;\n\t"
        "madc.hi.cc.u32 t177, %12, t136, t177;\n\t"
        "madc.lo.cc.u32 t178, %14, t136, t178;\n\t"
        "madc.hi.cc.u32 t179, %14, t136, t179;\n\t"
        "addc.u32 t180, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t157, %8, t139, t157;\n\t"
        "madc.hi.cc.u32 t158, %8, t139, t158;\n\t"
        "madc.lo.cc.u32 t159, %10, t139, t159;\n\t"
        "madc.hi.cc.u32 t160, %10, t139, t160;\n\t"
        "madc.lo.cc.u32 t161, %12, t139, t161;\n\t"
        "madc.hi.cc.u32 t162, %12, t139, t162;\n\t"
        "madc.lo.cc.u32 t163, %14, t139, t163;\n\t"
        "madc.hi.cc.u32 t164, %14, t139, t164;\n\t"
        "addc.u32 t165, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t174, %9, t139, t174;\n\t"
        "madc.hi.cc.u32 t175, %9, t139, t175;\n\t"
        "madc.lo.cc.u32 t176, %11, t139, t176;\n\t"
        "madc.hi.cc.u32 t177, %11, t139, t177;\n\t"
        "madc.lo.cc.u32 t178, %13, t139, t178;\n\t"
        "madc.hi.cc.u32 t179, %13, t139, t179;\n\t"
        "madc.lo.cc.u32 t180, %15, t139, t180;\n\t"
        "madc.hi.u32 t181, %15, t139, 0x0;\n\t"
        "mad.lo.cc.u32 t159, %9, t142, t159;\n\t"
        "madc.hi.cc.u32 t160, %9, t142, t160;\n\t"
        "madc.lo.cc.u32 t161, %11, t142, t161;\n\t"
        "madc.hi.cc.u32 t162, %11, t142, t162;\n\t"
        "madc.lo.cc.u32 t163, %13, t142, t163;\n\t"
        "madc.hi.cc.u32 t164, %13, t142, t164;\n\t"
        "madc.lo.cc.u32 t165, %15, t142, t165;\n\t"
        "madc.hi.u32 t166, %15, t142, 0x0;\n\t"
        "mad.lo.cc.u32 t174, %8, t142, t174;\n\t"
        "madc.hi.cc.u32 t175, %8, t142, t175;\n\t"
        "madc.lo.cc.u32 t176, %10, t142, t176;\n\t"
        "madc.hi.cc.u32 t177, %10, t142, t177;\n\t"
        "madc.lo.cc.u32 t178, %12, t142, t178;\n\t"
        "madc.hi.cc.u32 t179, %12, t142, t179;\n\t"
        "madc.lo.cc.u32 t180, %14, t142, t180;\n\t"
        "madc.hi.cc.u32 t181, %14, t142, t181;\n\t"
        "addc.u32 t182, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t159, %8, t145, t159;\n\t"
        "madc.hi.cc.u32 t160, %8, t145, t160;\n\t"
        "madc.lo.cc.u32 t161, %10, t145, t161;\n\t"
        "madc.hi.cc.u32 t162, %10, t145, t162;\n\t"
        "madc.lo.cc.u32 t163, %12, t145, t163;\n\t"
        "madc.hi.cc.u32 t164, %12, t145, t164;\n\t"
        "madc.lo.cc.u32 t165, %14, t145, t165;\n\t"
        "madc.hi.cc.u32 t166, %14, t145, t166;\n\t"
        "addc.u32 t167, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t176, %9, t145, t176;\n\t"
        "madc.hi.cc.u32 t177, %9, t145, t177;\n\t"
        "madc.lo.cc.u32 t178, %11, t145, t178;\n\t"
        "madc.hi.cc.u32 t179, %11, t145, t179;\n\t"
        "madc.lo.cc.u32 t180, %13, t145, t180;\n\t"
        "madc.hi.cc.u32 t181, %13, t145, t181;\n\t"
        "madc.lo.cc.u32 t182, %15, t145, t182;\n\t"
        "madc.hi.u32 t183, %15, t145, 0x0;\n\t"
        "mad.lo.cc.u32 t161, %9, t148, t161;\n\t"
        "madc.hi.cc.u32 t162, %9, t148, t162;\n\t"
        "madc.lo.cc.u32 t163, %11, t148, t163;\n\t"
        "madc.hi.cc.u32 t164, %11, t148, t164;\n\t"
        "madc.lo.cc.u32 t165, %13, t148, t165;\n\t"
        "madc.hi.cc.u32 t166, %13, t148, t166;\n\t"
        "madc.lo.cc.u32 t167, %15, t148, t167;\n\t"
        "madc.hi.u32 t168, %15, t148, 0x0;\n\t"
        "mad.lo.cc.u32 t176, %8, t148, t176;\n\t"
        "madc.hi.cc.u32 t177, %8, t148, t177;\n\t"
        "madc.lo.cc.u32 t178, %10, t148, t178;\n\t"
        "madc.hi.cc.u32 t179, %10, t148, t179;\n\t"
        "madc.lo.cc.u32 t180, %12, t148, t180;\n\t"
        "madc.hi.cc.u32 t181, %12, t148, t181;\n\t"
        "madc.lo.cc.u32 t182, %14, t148, t182;\n\t"
        "madc.hi.cc.u32 t183, %14, t148, t183;\n\t"
        "addc.u32 t184, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t161, %8, t151, t161;\n\t"
        "madc.hi.cc.u32 t162, %8, t151, t162;\n\t"
        "madc.lo.cc.u32 t163, %10, t151, t163;\n\t"
        "madc.hi.cc.u32 t164, %10, t151, t164;\n\t"
        "madc.lo.cc.u32 t165, %12, t151, t165;\n\t"
        "madc.hi.cc.u32 t166, %12, t151, t166;\n\t"
        "madc.lo.cc.u32 t167, %14, t151, t167;\n\t"
        "madc.hi.cc.u32 t168, %14, t151, t168;\n\t"
        "addc.u32 t169, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t178, %9, t151, t178;\n\t"
        "madc.hi.cc.u32 t179, %9, t151, t179;\n\t"
        "madc.lo.cc.u32 t180, %11, t151, t180;\n\t"
        "madc.hi.cc.u32 t181, %11, t151, t181;\n\t"
        "madc.lo.cc.u32 t182, %13, t151, t182;\n\t"
        "madc.hi.cc.u32 t183, %13, t151, t183;\n\t"
        "madc.lo.cc.u32 t184, %15, t151, t184;\n\t"
        "madc.hi.u32 t185, %15, t151, 0x0;\n\t"
        "mad.lo.cc.u32 t163, %9, t154, t163;\n\t"
        "madc.hi.cc.u32 t164, %9, t154, t164;\n\t"
        "madc.lo.cc.u32 t165, %11, t154, t165;\n\t"
        "madc.hi.cc.u32 t166, %11, t154, t166;\n\t"
        "madc.lo.cc.u32 t167, %13, t154, t167;\n\t"
        "madc.hi.cc.u32 t168, %13, t154, t168;\n\t"
        "madc.lo.cc.u32 t169, %15, t154, t169;\n\t"
        "madc.hi.u32 t170, %15, t154, 0x0;\n\t"
        "mad.lo.cc.u32 t178, %8, t154, t178;\n\t"
        "madc.hi.cc.u32 t179, %8, t154, t179;\n\t"
        "madc.lo.cc.u32 t180, %10, t154, t180;\n\t"
        "madc.hi.cc.u32 t181, %10, t154, t181;\n\t"
        "madc.lo.cc.u32 t182, %12, t154, t182;\n\t"
        "madc.hi.cc.u32 t183, %12, t154, t183;\n\t"
        "madc.lo.cc.u32 t184, %14, t154, t184;\n\t"
        "madc.hi.cc.u32 t185, %14, t154, t185;\n\t"
        "addc.u32 t186, 0x0, 0x0;\n\t"
        "add.cc.u32 t187, t156, t172;\n\t"
        "addc.cc.u32 t188, t157, t173;\n\t"
        "addc.cc.u32 t189, t158, t174;\n\t"
        "addc.cc.u32 t190, t159, t175;\n\t"
        "addc.cc.u32 t191, t160, t176;\n\t"
        "addc.cc.u32 t192, t161, t177;\n\t"
        "addc.cc.u32 t193, t162, t178;\n\t"
        "addc.cc.u32 t194, t163, t179;\n\t"
        "addc.cc.u32 t195, t164, t180;\n\t"
        "addc.cc.u32 t196, t165, t181;\n\t"
        "addc.cc.u32 t197, t166, t182;\n\t"
        "addc.cc.u32 t198, t167, t183;\n\t"
        "addc.cc.u32 t199, t168, t184;\n\t"
        "addc.cc.u32 t200, t169, t185;\n\t"
        "addc.u32 t201, t170, t186;\n\t"
        "mul.lo.u32 t202, t155, 0xee00bc4f;\n\t"
        "mul.hi.u32 t203, t155, 0xee00bc4f;\n\t"
        "mul.lo.u32 t204, t188, 0xee00bc4f;\n\t"
        "mul.hi.u32 t205, t188, 0xee00bc4f;\n\t"
        "mul.lo.u32 t206, t190, 0xee00bc4f;\n\t"
        "mul.hi.u32 t207, t190, 0xee00bc4f;\n\t"
        "mul.lo.u32 t208, t192, 0xee00bc4f;\n\t"
        "mul.hi.u32 t209, t192, 0xee00bc4f;\n\t"
        "mul.lo.u32 t212, t187, 0xee00bc4f;\n\t"
        "mul.hi.u32 t213, t187, 0xee00bc4f;\n\t"
        "mul.lo.u32 t214, t189, 0xee00bc4f;\n\t"
        "mul.hi.u32 t215, t189, 0xee00bc4f;\n\t"
        "mul.lo.u32 t216, t191, 0xee00bc4f;\n\t"
        "mul.hi.u32 t217, t191, 0xee00bc4f;\n\t"
        "mad.lo.cc.u32 t204, t187, 0xccd1c8aa, t204;\n\t"
        "madc.hi.cc.u32 t205, t187, 0xccd1c8aa, t205;\n\t"
        "madc.lo.cc.u32 t206, t189, 0xccd1c8aa, t206;\n\t"
        "madc.hi.cc.u32 t207, t189, 0xccd1c8aa, t207;\n\t"
        "madc.lo.cc.u32 t208, t191, 0xccd1c8aa, t208;\n\t"
        "madc.hi.cc.u32 t209, t191, 0xccd1c8aa, t209;\n\t"
        "addc.u32 t210, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t212, t155, 0xccd1c8aa, t212;\n\t"
        "madc.hi.cc.u32 t213, t155, 0xccd1c8aa, t213;\n\t"
        "madc.lo.cc.u32 t214, t188, 0xccd1c8aa, t214;\n\t"
        "madc.hi.cc.u32 t215, t188, 0xccd1c8aa, t215;\n\t"
        "madc.lo.cc.u32 t216, t190, 0xccd1c8aa, t216;\n\t"
        "madc.hi.cc.u32 t217, t190, 0xccd1c8aa, t217;\n\t"
        "addc.u32 t218, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t204, t155, 0x7d74d2e4, t204;\n\t"
        "madc.hi.cc.u32 t205, t155, 0x7d74d2e4, t205;\n\t"
        "madc.lo.cc.u32 t206, t188, 0x7d74d2e4, t206;\n\t"
        "madc.hi.cc.u32 t207, t188, 0x7d74d2e4, t207;\n\t"
        "madc.lo.cc.u32 t208, t190, 0x7d74d2e4, t208;\n\t"
        "madc.hi.cc.u32 t209, t190, 0x7d74d2e4, t209;\n\t"
        "addc.u32 t210, t210, 0x0;\n\t"
        "mad.lo.cc.u32 t214, t187, 0x7d74d2e4, t214;\n\t"
        "madc.hi.cc.u32 t215, t187, 0x7d74d2e4, t215;\n\t"
        "madc.lo.cc.u32 t216, t189, 0x7d74d2e4, t216;\n\t"
        "madc.hi.cc.u32 t217, t189, 0x7d74d2e4, t217;\n\t"
        "addc.u32 t218, t218, 0x0;\n\t"
        "mad.lo.cc.u32 t206, t187, 0x48c94408, t206;\n\t"
        "madc.hi.cc.u32 t207, t187, 0x48c94408, t207;\n\t"
        "madc.lo.cc.u32 t208, t189, 0x48c94408, t208;\n\t"
        "madc.hi.cc.u32 t209, t189, 0x48c94408, t209;\n\t"
        "addc.u32 t210, t210, 0x0;\n\t"
        "mad.lo.cc.u32 t214, t155, 0x48c94408, t214;\n\t"
        "madc.hi.cc.u32 t215, t155, 0x48c94408, t215;\n\t"
        "madc.lo.cc.u32 t216, t188, 0x48c94408, t216;\n\t"
        "madc.hi.cc.u32 t217, t188, 0x48c94408, t217;\n\t"
        "addc.u32 t218, t218, 0x0;\n\t"
        "mad.lo.cc.u32 t206, t155, 0xc588c6f6, t206;\n\t"
        "madc.hi.cc.u32 t207, t155, 0xc588c6f6, t207;\n\t"
        "madc.lo.cc.u32 t208, t188, 0xc588c6f6, t208;\n\t"
        "madc.hi.cc.u32 t209, t188, 0xc588c6f6, t209;\n\t"
        "addc.u32 t210, t210, 0x0;\n\t"
        "mad.lo.cc.u32 t216, t187, 0xc588c6f6, t216;\n\t"
        "madc.hi.cc.u32 t217, t187, 0xc588c6f6, t217;\n\t"
        "addc.u32 t218, t218, 0x0;\n\t"
        "mad.lo.cc.u32 t208, t187, 0x50fe77ec, t208;\n\t"
        "madc.hi.cc.u32 t209, t187, 0x50fe77ec, t209;\n\t"
        "addc.u32 t210, t210, 0x0;\n\t"
        "mad.lo.cc.u32 t216, t155, 0x50fe77ec, t216;\n\t"
        "madc.hi.cc.u32 t217, t155, 0x50fe77ec, t217;\n\t"
        "addc.u32 t218, t218, 0x0;\n\t"
        "mad.lo.cc.u32 t208, t155, 0xa9d6281c, t208;\n\t"
        "madc.hi.cc.u32 t209, t155, 0xa9d6281c, t209;\n\t"
        "addc.u32 t210, t210, 0x0;\n\t"
        "add.cc.u32 t220, t203, t212;\n\t"
        "addc.cc.u32 t221, t204, t213;\n\t"
        "addc.cc.u32 t222, t205, t214;\n\t"
        "addc.cc.u32 t223, t206, t215;\n\t"
        "addc.cc.u32 t224, t207, t216;\n\t"
        "addc.cc.u32 t225, t208, t217;\n\t"
        "addc.u32 t226, t209, t218;\n\t"
        "mad.lo.u32 t227, t193, 0xee00bc4f, t226;\n\t"
        "mad.lo.u32 t228, t192, 0xccd1c8aa, t227;\n\t"
        "mad.lo.u32 t229, t191, 0x7d74d2e4, t228;\n\t"
        "mad.lo.u32 t230, t190, 0x48c94408, t229;\n\t"
        "mad.lo.u32 t231, t189, 0xc588c6f6, t230;\n\t"
        "mad.lo.u32 t232, t188, 0x50fe77ec, t231;\n\t"
        "mad.lo.u32 t233, t187, 0xa9d6281c, t232;\n\t"
        "mad.lo.u32 t234, t155, 0x60d06633, t233;\n\t"
        "mul.lo.u32 t235, t202, 0xfc632551;\n\t"
        "mul.hi.u32 t236, t202, 0xfc632551;\n\t"
        "mul.lo.u32 t237, t221, 0xfc632551;\n\t"
        "mul.hi.u32 t238, t221, 0xfc632551;\n\t"
        "mul.lo.u32 t239, t223, 0xfc632551;\n\t"
        "mul.hi.u32 t240, t223, 0xfc632551;\n\t"
        "mul.lo.u32 t241, t225, 0xfc632551;\n\t"
        "mul.hi.u32 t242, t225, 0xfc632551;\n\t"
        "mul.lo.u32 t252, t220, 0xfc632551;\n\t"
        "mul.hi.u32 t253, t220, 0xfc632551;\n\t"
        "mul.lo.u32 t254, t222, 0xfc632551;\n\t"
        "mul.hi.u32 t255, t222, 0xfc632551;\n\t"
        "mul.lo.u32 t256, t224, 0xfc632551;\n\t"
        "mul.hi.u32 t257, t224, 0xfc632551;\n\t"
        "mul.lo.u32 t258, t234, 0xfc632551;\n\t"
        "mul.hi.u32 t259, t234, 0xfc632551;\n\t"
        "mad.lo.cc.u32 t237, t220, 0xf3b9cac2, t237;\n\t"
        "madc.hi.cc.u32 t238, t220, 0xf3b9cac2, t238;\n\t"
        "madc.lo.cc.u32 t239, t222, 0xf3b9cac2, t239;\n\t"
        "madc.hi.cc.u32 t240, t222, 0xf3b9cac2, t240;\n\t"
        "madc.lo.cc.u32 t241, t224, 0xf3b9cac2, t241;\n\t"
        "madc.hi.cc.u32 t242, t224, 0xf3b9cac2, t242;\n\t"
        "madc.lo.cc.u32 t243, t234, 0xf3b9cac2, 0x0;\n\t"
        "madc.hi.u32 t244, t234, 0xf3b9cac2, 0x0;\n\t"
        "mad.lo.cc.u32 t252, t202, 0xf3b9cac2, t252;\n\t"
        "madc.hi.cc.u32 t253, t202, 0xf3b9cac2, t253;\n\t"
        "madc.lo.cc.u32 t254, t221, 0xf3b9cac2, t254;\n\t"
        "madc.hi.cc.u32 t255, t221, 0xf3b9cac2, t255;\n\t"
        "madc.lo.cc.u32 t256, t223, 0xf3b9cac2, t256;\n\t"
        "madc.hi.cc.u32 t257, t223, 0xf3b9cac2, t257;\n\t"
        "madc.lo.cc.u32 t258, t225, 0xf3b9cac2, t258;\n\t"
        "madc.hi.cc.u32 t259, t225, 0xf3b9cac2, t259;\n\t"
        "addc.u32 t260, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t237, t202, 0xa7179e84, t237;\n\t"
        "madc.hi.cc.u32 t238, t202, 0xa7179e84, t238;\n\t"
        "madc.lo.cc.u32 t239, t221, 0xa7179e84, t239;\n\t"
        "madc.hi.cc.u32 t240, t221, 0xa7179e84, t240;\n\t"
        "madc.lo.cc.u32 t241, t223, 0xa7179e84, t241;\n\t"
        "madc.hi.cc.u32 t242, t223, 0xa7179e84, t242;\n\t"
        "madc.lo.cc.u32 t243, t225, 0xa7179e84, t243;\n\t"
        "madc.hi.cc.u32 t244, t225, 0xa7179e84, t244;\n\t"
        "addc.u32 t245, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t254, t220, 0xa7179e84, t254;\n\t"
        "madc.hi.cc.u32 t255, t220, 0xa7179e84, t255;\n\t"
        "madc.lo.cc.u32 t256, t222, 0xa7179e84, t256;\n\t"
        "madc.hi.cc.u32 t257, t222, 0xa7179e84, t257;\n\t"
        "madc.lo.cc.u32 t258, t224, 0xa7179e84, t258;\n\t"
        "madc.hi.cc.u32 t259, t224, 0xa7179e84, t259;\n\t"
        "madc.lo.cc.u32 t260, t234, 0xa7179e84, t260;\n\t"
        "madc.hi.u32 t261, t234, 0xa7179e84, 0x0;\n\t"
        "mad.lo.cc.u32 t239, t220, 0xbce6faad, t239;\n\t"
        "madc.hi.cc.u32 t240, t220, 0xbce6faad, t240;\n\t"
        "madc.lo.cc.u32 t241, t222, 0xbce6faad, t241;\n\t"
        "madc.hi.cc.u32 t242, t222, 0xbce6faad, t242;\n\t"
        "madc.lo.cc.u32 t243, t224, 0xbce6faad, t243;\n\t"
        "madc.hi.cc.u32 t244, t224, 0xbce6faad, t244;\n\t"
        "madc.lo.cc.u32 t245, t234, 0xbce6faad, t245;\n\t"
        "madc.hi.u32 t246, t234, 0xbce6faad, 0x0;\n\t"
        "mad.lo.cc.u32 t254, t202, 0xbce6faad, t254;\n\t"
        "madc.hi.cc.u32 t255, t202, 0xbce6faad, t255;\n\t"
        "madc.lo.cc.u32 t256, t221, 0xbce6faad, t256;\n\t"
        "madc.hi.cc.u32 t257, t221, 0xbce6faad, t257;\n\t"
        "madc.lo.cc.u32 t258, t223, 0xbce6faad, t258;\n\t"
        "madc.hi.cc.u32 t259, t223, 0xbce6faad, t259;\n\t"
        "madc.lo.cc.u32 t260, t225, 0xbce6faad, t260;\n\t"
        "madc.hi.cc.u32 t261, t225, 0xbce6faad, t261;\n\t"
        "addc.u32 t262, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t239, t202, 0xffffffff, t239;\n\t"
        "madc.hi.cc.u32 t240, t202, 0xffffffff, t240;\n\t"
        "madc.lo.cc.u32 t241, t221, 0xffffffff, t241;\n\t"
        "madc.hi.cc.u32 t242, t221, 0xffffffff, t242;\n\t"
        "madc.lo.cc.u32 t243, t223, 0xffffffff, t243;\n\t"
        "madc.hi.cc.u32 t244, t223, 0xffffffff, t244;\n\t"
        "madc.lo.cc.u32 t245, t225, 0xffffffff, t245;\n\t"
        "madc.hi.cc.u32 t246, t225, 0xffffffff, t246;\n\t"
        "addc.u32 t247, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t256, t220, 0xffffffff, t256;\n\t"
        "madc.hi.cc.u32 t257, t220, 0xffffffff, t257;\n\t"
        "madc.lo.cc.u32 t258, t222, 0xffffffff, t258;\n\t"
        "madc.hi.cc.u32 t259, t222, 0xffffffff, t259;\n\t"
        "madc.lo.cc.u32 t260, t224, 0xffffffff, t260;\n\t"
        "madc.hi.cc.u32 t261, t224, 0xffffffff, t261;\n\t"
        "madc.lo.cc.u32 t262, t234, 0xffffffff, t262;\n\t"
        "madc.hi.u32 t263, t234, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t241, t220, 0xffffffff, t241;\n\t"
        "madc.hi.cc.u32 t242, t220, 0xffffffff, t242;\n\t"
        "madc.lo.cc.u32 t243, t222, 0xffffffff, t243;\n\t"
        "madc.hi.cc.u32 t244, t222, 0xffffffff, t244;\n\t"
        "madc.lo.cc.u32 t245, t224, 0xffffffff, t245;\n\t"
        "madc.hi.cc.u32 t246, t224, 0xffffffff, t246;\n\t"
        "madc.lo.cc.u32 t247, t234, 0xffffffff, t247;\n\t"
        "madc.hi.u32 t248, t234, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t256, t202, 0xffffffff, t256;\n\t"
        "madc.hi.cc.u32 t257, t202, 0xffffffff, t257;\n\t"
        "madc.lo.cc.u32 t258, t221, 0xffffffff, t258;\n\t"
        "madc.hi.cc.u32 t259, t221, 0xffffffff, t259;\n\t"
        "madc.lo.cc.u32 t260, t223, 0xffffffff, t260;\n\t"
        "madc.hi.cc.u32 t261, t223, 0xffffffff, t261;\n\t"
        "madc.lo.cc.u32 t262, t225, 0xffffffff, t262;\n\t"
        "madc.hi.cc.u32 t263, t225, 0xffffffff, t263;\n\t"
        "addc.u32 t264, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t241, t202, 0x0, t241;\n\t"
        "madc.hi.cc.u32 t242, t202, 0x0, t242;\n\t"
        "madc.lo.cc.u32 t243, t221, 0x0, t243;\n\t"
        "madc.hi.cc.u32 t244, t221, 0x0, t244;\n\t"
        "madc.lo.cc.u32 t245, t223, 0x0, t245;\n\t"
        "madc.hi.cc.u32 t246, t223, 0x0, t246;\n\t"
        "madc.lo.cc.u32 t247, t225, 0x0, t247;\n\t"
        "madc.hi.cc.u32 t248, t225, 0x0, t248;\n\t"
        "addc.u32 t249, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t258, t220, 0x0, t258;\n\t"
        "madc.hi.cc.u32 t259, t220, 0x0, t259;\n\t"
        "madc.lo.cc.u32 t260, t222, 0x0, t260;\n\t"
        "madc.hi.cc.u32 t261, t222, 0x0, t261;\n\t"
        "madc.lo.cc.u32 t262, t224, 0x0, t262;\n\t"
        "madc.hi.cc.u32 t263, t224, 0x0, t263;\n\t"
        "madc.lo.cc.u32 t264, t234, 0x0, t264;\n\t"
        "madc.hi.u32 t265, t234, 0x0, 0x0;\n\t"
        "mad.lo.cc.u32 t243, t220, 0xffffffff, t243;\n\t"
        "madc.hi.cc.u32 t244, t220, 0xffffffff, t244;\n\t"
        "madc.lo.cc.u32 t245, t222, 0xffffffff, t245;\n\t"
        "madc.hi.cc.u32 t246, t222, 0xffffffff, t246;\n\t"
        "madc.lo.cc.u32 t247, t224, 0xffffffff, t247;\n\t"
        "madc.hi.cc.u32 t248, t224, 0xffffffff, t248;\n\t"
        "madc.lo.cc.u32 t249, t234, 0xffffffff, t249;\n\t"
        "madc.hi.u32 t250, t234, 0xffffffff, 0x0;\n\t"
        "mad.lo.cc.u32 t258, t202, 0xffffffff, t258;\n\t"
        "madc.hi.cc.u32 t259, t202, 0xffffffff, t259;\n\t"
        "madc.lo.cc.u32 t260, t221, 0xffffffff, t260;\n\t"
        "madc.hi.cc.u32 t261, t221, 0xffffffff, t261;\n\t"
        "madc.lo.cc.u32 t262, t223, 0xffffffff, t262;\n\t"
        "madc.hi.cc.u32 t263, t223, 0xffffffff, t263;\n\t"
        "madc.lo.cc.u32 t264, t225, 0xffffffff, t264;\n\t"
        "madc.hi.cc.u32 t265, t225, 0xffffffff, t265;\n\t"
        "addc.u32 t266, 0x0, 0x0;\n\t"
        "add.cc.u32 t267, t236, t252;\n\t"
        "addc.cc.u32 t268, t237, t253;\n\t"
        "addc.cc.u32 t269, t238, t254;\n\t"
        "addc.cc.u32 t270, t239, t255;\n\t"
        "addc.cc.u32 t271, t240, t256;\n\t"
        "addc.cc.u32 t272, t241, t257;\n\t"
        "addc.cc.u32 t273, t242, t258;\n\t"
        "addc.cc.u32 t274, t243, t259;\n\t"
        "addc.cc.u32 t275, t244, t260;\n\t"
        "addc.cc.u32 t276, t245, t261;\n\t"
        "addc.cc.u32 t277, t246, t262;\n\t"
        "addc.cc.u32 t278, t247, t263;\n\t"
        "addc.cc.u32 t279, t248, t264;\n\t"
        "addc.cc.u32 t280, t249, t265;\n\t"
        "addc.u32 t281, t250, t266;\n\t"
        "add.cc.u32 t282, t155, t235;\n\t"
        "addc.cc.u32 t283, t187, t267;\n\t"
        "addc.cc.u32 t284, t188, t268;\n\t"
        "addc.cc.u32 t285, t189, t269;\n\t"
        "addc.cc.u32 t286, t190, t270;\n\t"
        "addc.cc.u32 t287, t191, t271;\n\t"
        "addc.cc.u32 t288, t192, t272;\n\t"
        "addc.cc.u32 t289, t193, t273;\n\t"
        "addc.cc.u32 t290, t194, t274;\n\t"
        "addc.cc.u32 t291, t195, t275;\n\t"
        "addc.cc.u32 t292, t196, t276;\n\t"
        "addc.cc.u32 t293, t197, t277;\n\t"
        "addc.cc.u32 t294, t198, t278;\n\t"
        "addc.cc.u32 t295, t199, t279;\n\t"
        "addc.cc.u32 t296, t200, t280;\n\t"
        "addc.cc.u32 t297, t201, t281;\n\t"
        "addc.u32 t298, 0x0, 0x0;\n\t"
        "sub.cc.u32 t299, t290, 0xfc632551;\n\t"
        "subc.cc.u32 t300, t291, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t301, t292, 0xa7179e84;\n\t"
        "subc.cc.u32 t302, t293, 0xbce6faad;\n\t"
        "subc.cc.u32 t303, t294, 0xffffffff;\n\t"
        "subc.cc.u32 t304, t295, 0xffffffff;\n\t"
        "subc.cc.u32 t305, t296, 0x0;\n\t"
        "subc.cc.u32 t306, t297, 0xffffffff;\n\t"
        "subc.cc.u32 t307, t298, 0x0;\n\t"
        "subc.u32 t308, 0x0, 0x0;\n\t"
        "xor.b32 t309, t299, t290;\n\t"
        "and.b32 t310, t309, t308;\n\t"
        "xor.b32 t311, t310, t299;\n\t"
        "xor.b32 t312, t300, t291;\n\t"
        "and.b32 t313, t312, t308;\n\t"
        "xor.b32 t314, t313, t300;\n\t"
        "xor.b32 t315, t301, t292;\n\t"
        "and.b32 t316, t315, t308;\n\t"
        "xor.b32 t317, t316, t301;\n\t"
        "xor.b32 t318, t302, t293;\n\t"
        "and.b32 t319, t318, t308;\n\t"
        "xor.b32 t320, t319, t302;\n\t"
        "xor.b32 t321, t303, t294;\n\t"
        "and.b32 t322, t321, t308;\n\t"
        "xor.b32 t323, t322, t303;\n\t"
        "xor.b32 t324, t304, t295;\n\t"
        "and.b32 t325, t324, t308;\n\t"
        "xor.b32 t326, t325, t304;\n\t"
        "xor.b32 t327, t305, t296;\n\t"
        "and.b32 t328, t327, t308;\n\t"
        "xor.b32 t329, t328, t305;\n\t"
        "xor.b32 t330, t306, t297;\n\t"
        "and.b32 t331, t330, t308;\n\t"
        "xor.b32 t332, t331, t306;\n\t"
        "add.cc.u32 t333, t311, %16;\n\t"
        "addc.cc.u32 t334, t314, %17;\n\t"
        "addc.cc.u32 t335, t317, %18;\n\t"
        "addc.cc.u32 t336, t320, %19;\n\t"
        "addc.cc.u32 t337, t323, %20;\n\t"
        "addc.cc.u32 t338, t326, %21;\n\t"
        "addc.cc.u32 t339, t329, %22;\n\t"
        "addc.cc.u32 t340, t332, %23;\n\t"
        "addc.u32 t341, 0x0, 0x0;\n\t"
        "sub.cc.u32 t342, t333, 0xfc632551;\n\t"
        "subc.cc.u32 t343, t334, 0xf3b9cac2;\n\t"
        "subc.cc.u32 t344, t335, 0xa7179e84;\n\t"
        "subc.cc.u32 t345, t336, 0xbce6faad;\n\t"
        "subc.cc.u32 t346, t337, 0xffffffff;\n\t"
        "subc.cc.u32 t347, t338, 0xffffffff;\n\t"
        "subc.cc.u32 t348, t339, 0x0;\n\t"
        "subc.cc.u32 t349, t340, 0xffffffff;\n\t"
        "subc.cc.u32 t350, t341, 0x0;\n\t"
        "subc.u32 t351, 0x0, 0x0;\n\t"
        "xor.b32 t352, t342, t333;\n\t"
        "and.b32 t353, t352, t351;\n\t"
        "xor.b32 t354, t353, t342;\n\t"
        "xor.b32 t355, t343, t334;\n\t"
        "and.b32 t356, t355, t351;\n\t"
        "xor.b32 t357, t356, t343;\n\t"
        "xor.b32 t358, t344, t335;\n\t"
        "and.b32 t359, t358, t351;\n\t"
        "xor.b32 t360, t359, t344;\n\t"
        "xor.b32 t361, t345, t336;\n\t"
        "and.b32 t362, t361, t351;\n\t"
        "xor.b32 t363, t362, t345;\n\t"
        "xor.b32 t364, t346, t337;\n\t"
        "and.b32 t365, t364, t351;\n\t"
        "xor.b32 t366, t365, t346;\n\t"
        "xor.b32 t367, t347, t338;\n\t"
        "and.b32 t368, t367, t351;\n\t"
        "xor.b32 t369, t368, t347;\n\t"
        "xor.b32 t370, t348, t339;\n\t"
        "and.b32 t371, t370, t351;\n\t"
        "xor.b32 t372, t371, t348;\n\t"
        "xor.b32 t373, t349, t340;\n\t"
        "and.b32 t374, t373, t351;\n\t"
        "xor.b32 t375, t374, t349;\n\t"
        "mov.u32 %0, t354;\n\t"
        "mov.u32 %1, t357;\n\t"
        "mov.u32 %2, t360;\n\t"
        "mov.u32 %3, t363;\n\t"
        "mov.u32 %4, t366;\n\t"
        "mov.u32 %5, t369;\n\t"
        "mov.u32 %6, t372;\n\t"
        "mov.u32 %7, t375;\n\t"
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
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31, t32, t33, t34, t35, t36, t37, t38, t39, t40, t41, t42, t43, t44, t45, t46, t47, t48, t49, t50, t51, t52, t53, t54, t55, t56, t57, t58, t59, t60, t61, t62, t63, t64, t65, t66, t67, t68, t69, t70, t71, t72, t73, t74, t75, t76, t77, t78, t79, t80, t81, t82, t83, t84, t85, t86, t87, t88, t89, t90, t91, t92, t93, t94, t95, t96, t97, t98, t99, t100, t101, t102, t103, t104, t105, t106, t107, t108, t109, t110, t111, t112, t113, t114, t115, t116, t117, t118, t119, t120, t121, t122, t123, t124, t125, t126, t127, t128, t129, t130, t131, t132, t133, t134, t135, t136, t137, t138, t139, t140, t141, t142, t143, t144, t145, t146, t147, t148, t149, t150, t151, t152, t153, t154, t155, t156, t157, t158, t159, t160, t161, t162, t163, t164, t165, t166, t167, t168, t169, t170, t171, t172, t173, t174, t175, t176, t177, t178, t179, t180, t181, t182, t183, t184, t185, t186, t187, t188, t189, t190, t191, t192, t193, t194, t195, t196, t197, t198, t199, t200, t201, t202, t203, t204, t205, t206, t207, t208, t209, t210, t211, t212, t213, t214, t215, t216, t217, t218, t219, t220, t221, t222, t223, t224, t225, t226, t227, t228, t229, t230, t231, t232, t233, t234, t235, t236, t237, t238, t239, t240, t241, t242, t243, t244, t245, t246, t247, t248, t249, t250, t251, t252, t253, t254, t255, t256, t257, t258, t259, t260, t261, t262, t263, t264, t265, t266, t267, t268, t269, t270, t271, t272, t273, t274, t275, t276, t277, t278, t279, t280, t281, t282, t283, t284, t285, t286, t287, t288, t289, t290, t291, t292, t293, t294, t295, t296, t297, t298, t299, t300, t301, t302, t303, t304, t305, t306, t307, t308, t309, t310, t311, t312, t313, t314, t315, t316, t317, t318, t319, t320, t321, t322, t323, t324, t325, t326, t327, t328, t329, t330, t331, t332, t333, t334, t335, t336, t337, t338, t339, t340, t341, t342, t343, t344, t345, t346, t347, t348, t349, t350, t351, t352, t353, t354, t355, t356, t357, t358, t359, t360, t361, t362, t363, t364, t365, t366, t367, t368, t369, t370, t371, t372, t373, t374, t375;
    uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;
    t0 = (uint32_t)((uint32_t)(0xbe79eea2u * b_i));
    t8 = (uint32_t)(((uint64_t)0xbe79eea2u * b_i) >> 32);
    t1 = (uint32_t)((uint32_t)(0x83244c95u * b_i));
    t9 = (uint32_t)(((uint64_t)0x83244c95u * b_i) >> 32);
    t2 = (uint32_t)((uint32_t)(0x49bd6fa6u * b_i));
    t10 = (uint32_t)(((uint64_t)0x49bd6fa6u * b_i) >> 32);
    t3 = (uint32_t)((uint32_t)(0x4699799cu * b_i));
    t11 = (uint32_t)(((uint64_t)0x4699799cu * b_i) >> 32);
    t4 = (uint32_t)((uint32_t)(0x2b6bec59u * b_i));
    t12 = (uint32_t)(((uint64_t)0x2b6bec59u * b_i) >> 32);
    t5 = (uint32_t)((uint32_t)(0x2845b239u * b_i));
    t13 = (uint32_t)(((uint64_t)0x2845b239u * b_i) >> 32);
    t6 = (uint32_t)((uint32_t)(0xf3d95620u * b_i));
    t14 = (uint32_t)(((uint64_t)0xf3d95620u * b_i) >> 32);
    t7 = (uint32_t)((uint32_t)(0x66e12d94u * b_i));
    t15 = (uint32_t)(((uint64_t)0x66e12d94u * b_i) >> 32);
    w_ = (uint64_t)t1 + t8; t16 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t2 + t9 + cf_; t17 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t3 + t10 + cf_; t18 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t4 + t11 + cf_; t19 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t5 + t12 + cf_; t20 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t6 + t13 + cf_; t21 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t7 + t14 + cf_; t22 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t15 + 0x0u + cf_; t23 = (uint32_t)w_;
    t24 = (uint32_t)((uint32_t)(t0 * 0xee00bc4fu));
    t25 = (uint32_t)(((uint64_t)t0 * 0xee00bc4fu) >> 32);
    t26 = (uint32_t)((uint32_t)(t17 * 0xee00bc4fu));
    t27 = (uint32_t)(((uint64_t)t17 * 0xee00bc4fu) >> 32);
    t28 = (uint32_t)((uint32_t)(t19 * 0xee00bc4fu));
    t29 = (uint32_t)(((uint64_t)t19 * 0xee00bc4fu) >> 32);
    t30 = (uint32_t)((uint32_t)(t21 * 0xee00bc4fu));
    t31 = (uint32_t)(((uint64_t)t21 * 0xee00bc4fu) >> 32);
    t34 = (uint32_t)((uint32_t)(t16 * 0xee00bc4fu));
    t35 = (uint32_t)(((uint64_t)t16 * 0xee00bc4fu) >> 32);
    t36 = (uint32_t)((uint32_t)(t18 * 0xee00bc4fu));
    t37 = (uint32_t)(((uint64_t)t18 * 0xee00bc4fu) >> 32);
    t38 = (uint32_t)((uint32_t)(t20 * 0xee00bc4fu));
    t39 = (uint32_t)(((uint64_t)t20 * 0xee00bc4fu) >> 32);
    w_ = (uint64_t)(uint32_t)(t16 * 0xccd1c8aau) + t26; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t16 * 0xccd1c8aau) >> 32) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t18 * 0xccd1c8aau) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t18 * 0xccd1c8aau) >> 32) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t20 * 0xccd1c8aau) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t20 * 0xccd1c8aau) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0xccd1c8aau) + t34; t34 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0xccd1c8aau) >> 32) + t35 + cf_; t35 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t17 * 0xccd1c8aau) + t36 + cf_; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t17 * 0xccd1c8aau) >> 32) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t19 * 0xccd1c8aau) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t19 * 0xccd1c8aau) >> 32) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t40 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x7d74d2e4u) + t26; t26 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0x7d74d2e4u) >> 32) + t27 + cf_; t27 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t17 * 0x7d74d2e4u) + t28 + cf_; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t17 * 0x7d74d2e4u) >> 32) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t19 * 0x7d74d2e4u) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t19 * 0x7d74d2e4u) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t16 * 0x7d74d2e4u) + t36; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t16 * 0x7d74d2e4u) >> 32) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t18 * 0x7d74d2e4u) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t18 * 0x7d74d2e4u) >> 32) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + 0x0u + cf_; t40 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t16 * 0x48c94408u) + t28; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t16 * 0x48c94408u) >> 32) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t18 * 0x48c94408u) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t18 * 0x48c94408u) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x48c94408u) + t36; t36 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0x48c94408u) >> 32) + t37 + cf_; t37 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t17 * 0x48c94408u) + t38 + cf_; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t17 * 0x48c94408u) >> 32) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + 0x0u + cf_; t40 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0xc588c6f6u) + t28; t28 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0xc588c6f6u) >> 32) + t29 + cf_; t29 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t17 * 0xc588c6f6u) + t30 + cf_; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t17 * 0xc588c6f6u) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t16 * 0xc588c6f6u) + t38; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t16 * 0xc588c6f6u) >> 32) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + 0x0u + cf_; t40 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t16 * 0x50fe77ecu) + t30; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t16 * 0x50fe77ecu) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x50fe77ecu) + t38; t38 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0x50fe77ecu) >> 32) + t39 + cf_; t39 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t40 + 0x0u + cf_; t40 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0xa9d6281cu) + t30; t30 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t0 * 0xa9d6281cu) >> 32) + t31 + cf_; t31 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t32 + 0x0u + cf_; t32 = (uint32_t)w_;
    w_ = (uint64_t)t25 + t34; t42 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t26 + t35 + cf_; t43 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t27 + t36 + cf_; t44 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t28 + t37 + cf_; t45 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t29 + t38 + cf_; t46 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t30 + t39 + cf_; t47 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t31 + t40 + cf_; t48 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t22 * 0xee00bc4fu) + t48; t49 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t21 * 0xccd1c8aau) + t49; t50 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t20 * 0x7d74d2e4u) + t50; t51 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t19 * 0x48c94408u) + t51; t52 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t18 * 0xc588c6f6u) + t52; t53 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t17 * 0x50fe77ecu) + t53; t54 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t16 * 0xa9d6281cu) + t54; t55 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t0 * 0x60d06633u) + t55; t56 = (uint32_t)w_;
    t57 = (uint32_t)((uint32_t)(t24 * 0xfc632551u));
    t58 = (uint32_t)(((uint64_t)t24 * 0xfc632551u) >> 32);
    t59 = (uint32_t)((uint32_t)(t43 * 0xfc632551u));
    t60 = (uint32_t)(((uint64_t)t43 * 0xfc632551u) >> 32);
    t61 = (uint32_t)((uint32_t)(t45 * 0xfc632551u));
    t62 = (uint32_t)(((uint64_t)t45 * 0xfc632551u) >> 32);
    t63 = (uint32_t)((uint32_t)(t47 * 0xfc632551u));
    t64 = (uint32_t)(((uint64_t)t47 * 0xfc632551u) >> 32);
    t74 = (uint32_t)((uint32_t)(t42 * 0xfc632551u));
    t75 = (uint32_t)(((uint64_t)t42 * 0xfc632551u) >> 32);
    t76 = (uint32_t)((uint32_t)(t44 * 0xfc632551u));
    t77 = (uint32_t)(((uint64_t)t44 * 0xfc632551u) >> 32);
    t78 = (uint32_t)((uint32_t)(t46 * 0xfc632551u));
    t79 = (uint32_t)(((uint64_t)t46 * 0xfc632551u) >> 32);
    t80 = (uint32_t)((uint32_t)(t56 * 0xfc632551u));
    t81 = (uint32_t)(((uint64_t)t56 * 0xfc632551u) >> 32);
    w_ = (uint64_t)(uint32_t)(t42 * 0xf3b9cac2u) + t59; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xf3b9cac2u) >> 32) + t60 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xf3b9cac2u) + t61 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xf3b9cac2u) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xf3b9cac2u) + t63 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xf3b9cac2u) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xf3b9cac2u) + 0x0u + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xf3b9cac2u) >> 32) + 0x0u + cf_; t66 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xf3b9cac2u) + t74; t74 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xf3b9cac2u) >> 32) + t75 + cf_; t75 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xf3b9cac2u) + t76 + cf_; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xf3b9cac2u) >> 32) + t77 + cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xf3b9cac2u) + t78 + cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xf3b9cac2u) >> 32) + t79 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xf3b9cac2u) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xf3b9cac2u) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t82 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xa7179e84u) + t59; t59 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xa7179e84u) >> 32) + t60 + cf_; t60 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xa7179e84u) + t61 + cf_; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xa7179e84u) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xa7179e84u) + t63 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xa7179e84u) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xa7179e84u) + t65 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xa7179e84u) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t67 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0xa7179e84u) + t76; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xa7179e84u) >> 32) + t77 + cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xa7179e84u) + t78 + cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xa7179e84u) >> 32) + t79 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xa7179e84u) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xa7179e84u) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xa7179e84u) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xa7179e84u) >> 32) + 0x0u + cf_; t83 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0xbce6faadu) + t61; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xbce6faadu) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xbce6faadu) + t63 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xbce6faadu) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xbce6faadu) + t65 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xbce6faadu) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xbce6faadu) + t67 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xbce6faadu) >> 32) + 0x0u + cf_; t68 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xbce6faadu) + t76; t76 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xbce6faadu) >> 32) + t77 + cf_; t77 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xbce6faadu) + t78 + cf_; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xbce6faadu) >> 32) + t79 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xbce6faadu) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xbce6faadu) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xbce6faadu) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xbce6faadu) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t84 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xffffffffu) + t61; t61 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xffffffffu) >> 32) + t62 + cf_; t62 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xffffffffu) + t63 + cf_; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xffffffffu) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xffffffffu) + t65 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xffffffffu) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xffffffffu) + t67 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xffffffffu) >> 32) + t68 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t69 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0xffffffffu) + t78; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xffffffffu) >> 32) + t79 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xffffffffu) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xffffffffu) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xffffffffu) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xffffffffu) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xffffffffu) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xffffffffu) >> 32) + 0x0u + cf_; t85 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0xffffffffu) + t63; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xffffffffu) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xffffffffu) + t65 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xffffffffu) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xffffffffu) + t67 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xffffffffu) >> 32) + t68 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xffffffffu) + t69 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xffffffffu) >> 32) + 0x0u + cf_; t70 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xffffffffu) + t78; t78 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xffffffffu) >> 32) + t79 + cf_; t79 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xffffffffu) + t80 + cf_; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xffffffffu) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xffffffffu) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xffffffffu) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xffffffffu) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xffffffffu) >> 32) + t85 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t86 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0x0u) + t63; t63 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0x0u) >> 32) + t64 + cf_; t64 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0x0u) + t65 + cf_; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0x0u) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0x0u) + t67 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0x0u) >> 32) + t68 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0x0u) + t69 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0x0u) >> 32) + t70 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t71 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0x0u) + t80; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0x0u) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0x0u) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0x0u) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0x0u) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0x0u) >> 32) + t85 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0x0u) + t86 + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0x0u) >> 32) + 0x0u + cf_; t87 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t42 * 0xffffffffu) + t65; t65 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t42 * 0xffffffffu) >> 32) + t66 + cf_; t66 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t44 * 0xffffffffu) + t67 + cf_; t67 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t44 * 0xffffffffu) >> 32) + t68 + cf_; t68 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t46 * 0xffffffffu) + t69 + cf_; t69 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t46 * 0xffffffffu) >> 32) + t70 + cf_; t70 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t56 * 0xffffffffu) + t71 + cf_; t71 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t56 * 0xffffffffu) >> 32) + 0x0u + cf_; t72 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t24 * 0xffffffffu) + t80; t80 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t24 * 0xffffffffu) >> 32) + t81 + cf_; t81 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t43 * 0xffffffffu) + t82 + cf_; t82 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t43 * 0xffffffffu) >> 32) + t83 + cf_; t83 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t45 * 0xffffffffu) + t84 + cf_; t84 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t45 * 0xffffffffu) >> 32) + t85 + cf_; t85 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t47 * 0xffffffffu) + t86 + cf_; t86 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t47 * 0xffffffffu) >> 32) + t87 + cf_; t87 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t88 = (uint32_t)w_;
    w_ = (uint64_t)t58 + t74; t89 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t59 + t75 + cf_; t90 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t60 + t76 + cf_; t91 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t61 + t77 + cf_; t92 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t62 + t78 + cf_; t93 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t63 + t79 + cf_; t94 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t64 + t80 + cf_; t95 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t65 + t81 + cf_; t96 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t66 + t82 + cf_; t97 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t67 + t83 + cf_; t98 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t68 + t84 + cf_; t99 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t69 + t85 + cf_; t100 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t70 + t86 + cf_; t101 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t71 + t87 + cf_; t102 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t72 + t88 + cf_; t103 = (uint32_t)w_;
    w_ = (uint64_t)t0 + t57; t104 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t16 + t89 + cf_; t105 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t17 + t90 + cf_; t106 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t18 + t91 + cf_; t107 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t19 + t92 + cf_; t108 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t20 + t93 + cf_; t109 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t21 + t94 + cf_; t110 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t22 + t95 + cf_; t111 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t23 + t96 + cf_; t112 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t97 + cf_; t113 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t98 + cf_; t114 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t99 + cf_; t115 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t100 + cf_; t116 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t101 + cf_; t117 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t102 + cf_; t118 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + t103 + cf_; t119 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t120 = (uint32_t)w_;
    w_ = (uint64_t)t112 - 0xfc632551u; t121 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t113 - 0xf3b9cac2u - cf_; t122 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t114 - 0xa7179e84u - cf_; t123 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t115 - 0xbce6faadu - cf_; t124 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t116 - 0xffffffffu - cf_; t125 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t117 - 0xffffffffu - cf_; t126 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t118 - 0x0u - cf_; t127 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t119 - 0xffffffffu - cf_; t128 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t120 - 0x0u - cf_; t129 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t130 = (uint32_t)w_;
    t131 = (uint32_t)(t121 ^ t112);
    t132 = (uint32_t)(t131 & t130);
    t133 = (uint32_t)(t132 ^ t121);
    t134 = (uint32_t)(t122 ^ t113);
    t135 = (uint32_t)(t134 & t130);
    t136 = (uint32_t)(t135 ^ t122);
    t137 = (uint32_t)(t123 ^ t114);
    t138 = (uint32_t)(t137 & t130);
    t139 = (uint32_t)(t138 ^ t123);
    t140 = (uint32_t)(t124 ^ t115);
    t141 = (uint32_t)(t140 & t130);
    t142 = (uint32_t)(t141 ^ t124);
    t143 = (uint32_t)(t125 ^ t116);
    t144 = (uint32_t)(t143 & t130);
    t145 = (uint32_t)(t144 ^ t125);
    t146 = (uint32_t)(t126 ^ t117);
    t147 = (uint32_t)(t146 & t130);
    t148 = (uint32_t)(t147 ^ t126);
    t149 = (uint32_t)(t127 ^ t118);
    t150 = (uint32_t)(t149 & t130);
    t151 = (uint32_t)(t150 ^ t127);
    t152 = (uint32_t)(t128 ^ t119);
    t153 = (uint32_t)(t152 & t130);
    t154 = (uint32_t)(t153 ^ t128);
    t155 = (uint32_t)((uint32_t)(a_0_i * t133));
    t156 = (uint32_t)(((uint64_t)a_0_i * t133) >> 32);
    t157 = (uint32_t)((uint32_t)(a_2_i * t133));
    t158 = (uint32_t)(((uint64_t)a_2_i * t133) >> 32);
    t159 = (uint32_t)((uint32_t)(a_4_i * t133));
    t160 = (uint32_t)(((uint64_t)a_4_i * t133) >> 32);
    t161 = (uint32_t)((uint32_t)(a_6_i * t133));
    t162 = (uint32_t)(((uint64_t)a_6_i * t133) >> 32);
    t172 = (uint32_t)((uint32_t)(a_1_i * t133));
    t173 = (uint32_t)(((uint64_t)a_1_i * t133) >> 32);
    t174 = (uint32_t)((uint32_t)(a_3_i * t133));
    t175 = (uint32_t)(((uint64_t)a_3_i * t133) >> 32);
    t176 = (uint32_t)((uint32_t)(a_5_i * t133));
    t177 = (uint32_t)(((uint64_t)a_5_i * t133) >> 32);
    t178 = (uint32_t)((uint32_t)(a_7_i * t133));
    t179 = (uint32_t)(((uint64_t)a_7_i * t133) >> 32);
    w_ = (uint64_t)(uint32_t)(a_1_i * t136) + t157; t157 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t136) >> 32) + t158 + cf_; t158 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t136) + t159 + cf_; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t136) >> 32) + t160 + cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t136) + t161 + cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t136) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t136) + 0x0u + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t136) >> 32) + 0x0u + cf_; t164 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t136) + t172; t172 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t136) >> 32) + t173 + cf_; t173 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t136) + t174 + cf_; t174 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t136) >> 32) + t175 + cf_; t175 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t136) + t176 + cf_; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t136) >> 32) + t177 + cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t136) + t178 + cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t136) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t180 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t139) + t157; t157 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t139) >> 32) + t158 + cf_; t158 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t139) + t159 + cf_; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t139) >> 32) + t160 + cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t139) + t161 + cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t139) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t139) + t163 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t139) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t165 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t139) + t174; t174 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t139) >> 32) + t175 + cf_; t175 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t139) + t176 + cf_; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t139) >> 32) + t177 + cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t139) + t178 + cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t139) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t139) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t139) >> 32) + 0x0u + cf_; t181 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t142) + t159; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t142) >> 32) + t160 + cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t142) + t161 + cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t142) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t142) + t163 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t142) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t142) + t165 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t142) >> 32) + 0x0u + cf_; t166 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t142) + t174; t174 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t142) >> 32) + t175 + cf_; t175 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t142) + t176 + cf_; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t142) >> 32) + t177 + cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t142) + t178 + cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t142) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t142) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t142) >> 32) + t181 + cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t182 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t145) + t159; t159 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t145) >> 32) + t160 + cf_; t160 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t145) + t161 + cf_; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t145) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t145) + t163 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t145) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t145) + t165 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t145) >> 32) + t166 + cf_; t166 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t167 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t145) + t176; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t145) >> 32) + t177 + cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t145) + t178 + cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t145) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t145) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t145) >> 32) + t181 + cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t145) + t182 + cf_; t182 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t145) >> 32) + 0x0u + cf_; t183 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t148) + t161; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t148) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t148) + t163 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t148) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t148) + t165 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t148) >> 32) + t166 + cf_; t166 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t148) + t167 + cf_; t167 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t148) >> 32) + 0x0u + cf_; t168 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t148) + t176; t176 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t148) >> 32) + t177 + cf_; t177 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t148) + t178 + cf_; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t148) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t148) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t148) >> 32) + t181 + cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t148) + t182 + cf_; t182 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t148) >> 32) + t183 + cf_; t183 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t184 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t151) + t161; t161 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t151) >> 32) + t162 + cf_; t162 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t151) + t163 + cf_; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t151) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t151) + t165 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t151) >> 32) + t166 + cf_; t166 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t151) + t167 + cf_; t167 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t151) >> 32) + t168 + cf_; t168 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t169 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t151) + t178; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t151) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t151) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t151) >> 32) + t181 + cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t151) + t182 + cf_; t182 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t151) >> 32) + t183 + cf_; t183 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t151) + t184 + cf_; t184 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t151) >> 32) + 0x0u + cf_; t185 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_1_i * t154) + t163; t163 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_1_i * t154) >> 32) + t164 + cf_; t164 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_3_i * t154) + t165 + cf_; t165 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_3_i * t154) >> 32) + t166 + cf_; t166 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_5_i * t154) + t167 + cf_; t167 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_5_i * t154) >> 32) + t168 + cf_; t168 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_7_i * t154) + t169 + cf_; t169 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_7_i * t154) >> 32) + 0x0u + cf_; t170 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(a_0_i * t154) + t178; t178 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_0_i * t154) >> 32) + t179 + cf_; t179 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_2_i * t154) + t180 + cf_; t180 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_2_i * t154) >> 32) + t181 + cf_; t181 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_4_i * t154) + t182 + cf_; t182 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_4_i * t154) >> 32) + t183 + cf_; t183 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(a_6_i * t154) + t184 + cf_; t184 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)a_6_i * t154) >> 32) + t185 + cf_; t185 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t186 = (uint32_t)w_;
    w_ = (uint64_t)t156 + t172; t187 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t157 + t173 + cf_; t188 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t158 + t174 + cf_; t189 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t159 + t175 + cf_; t190 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t160 + t176 + cf_; t191 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t161 + t177 + cf_; t192 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t162 + t178 + cf_; t193 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t163 + t179 + cf_; t194 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t164 + t180 + cf_; t195 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t165 + t181 + cf_; t196 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t166 + t182 + cf_; t197 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t167 + t183 + cf_; t198 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t168 + t184 + cf_; t199 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t169 + t185 + cf_; t200 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t170 + t186 + cf_; t201 = (uint32_t)w_;
    t202 = (uint32_t)((uint32_t)(t155 * 0xee00bc4fu));
    t203 = (uint32_t)(((uint64_t)t155 * 0xee00bc4fu) >> 32);
    t204 = (uint32_t)((uint32_t)(t188 * 0xee00bc4fu));
    t205 = (uint32_t)(((uint64_t)t188 * 0xee00bc4fu) >> 32);
    t206 = (uint32_t)((uint32_t)(t190 * 0xee00bc4fu));
    t207 = (uint32_t)(((uint64_t)t190 * 0xee00bc4fu) >> 32);
    t208 = (uint32_t)((uint32_t)(t192 * 0xee00bc4fu));
    t209 = (uint32_t)(((uint64_t)t192 * 0xee00bc4fu) >> 32);
    t212 = (uint32_t)((uint32_t)(t187 * 0xee00bc4fu));
    t213 = (uint32_t)(((uint64_t)t187 * 0xee00bc4fu) >> 32);
    t214 = (uint32_t)((uint32_t)(t189 * 0xee00bc4fu));
    t215 = (uint32_t)(((uint64_t)t189 * 0xee00bc4fu) >> 32);
    t216 = (uint32_t)((uint32_t)(t191 * 0xee00bc4fu));
    t217 = (uint32_t)(((uint64_t)t191 * 0xee00bc4fu) >> 32);
    w_ = (uint64_t)(uint32_t)(t187 * 0xccd1c8aau) + t204; t204 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t187 * 0xccd1c8aau) >> 32) + t205 + cf_; t205 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t189 * 0xccd1c8aau) + t206 + cf_; t206 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t189 * 0xccd1c8aau) >> 32) + t207 + cf_; t207 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t191 * 0xccd1c8aau) + t208 + cf_; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t191 * 0xccd1c8aau) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0xccd1c8aau) + t212; t212 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0xccd1c8aau) >> 32) + t213 + cf_; t213 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t188 * 0xccd1c8aau) + t214 + cf_; t214 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t188 * 0xccd1c8aau) >> 32) + t215 + cf_; t215 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t190 * 0xccd1c8aau) + t216 + cf_; t216 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t190 * 0xccd1c8aau) >> 32) + t217 + cf_; t217 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t218 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0x7d74d2e4u) + t204; t204 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0x7d74d2e4u) >> 32) + t205 + cf_; t205 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t188 * 0x7d74d2e4u) + t206 + cf_; t206 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t188 * 0x7d74d2e4u) >> 32) + t207 + cf_; t207 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t190 * 0x7d74d2e4u) + t208 + cf_; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t190 * 0x7d74d2e4u) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t210 + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t187 * 0x7d74d2e4u) + t214; t214 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t187 * 0x7d74d2e4u) >> 32) + t215 + cf_; t215 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t189 * 0x7d74d2e4u) + t216 + cf_; t216 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t189 * 0x7d74d2e4u) >> 32) + t217 + cf_; t217 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t218 + 0x0u + cf_; t218 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t187 * 0x48c94408u) + t206; t206 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t187 * 0x48c94408u) >> 32) + t207 + cf_; t207 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t189 * 0x48c94408u) + t208 + cf_; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t189 * 0x48c94408u) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t210 + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0x48c94408u) + t214; t214 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0x48c94408u) >> 32) + t215 + cf_; t215 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t188 * 0x48c94408u) + t216 + cf_; t216 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t188 * 0x48c94408u) >> 32) + t217 + cf_; t217 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t218 + 0x0u + cf_; t218 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0xc588c6f6u) + t206; t206 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0xc588c6f6u) >> 32) + t207 + cf_; t207 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t188 * 0xc588c6f6u) + t208 + cf_; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t188 * 0xc588c6f6u) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t210 + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t187 * 0xc588c6f6u) + t216; t216 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t187 * 0xc588c6f6u) >> 32) + t217 + cf_; t217 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t218 + 0x0u + cf_; t218 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t187 * 0x50fe77ecu) + t208; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t187 * 0x50fe77ecu) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t210 + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0x50fe77ecu) + t216; t216 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0x50fe77ecu) >> 32) + t217 + cf_; t217 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t218 + 0x0u + cf_; t218 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0xa9d6281cu) + t208; t208 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t155 * 0xa9d6281cu) >> 32) + t209 + cf_; t209 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t210 + 0x0u + cf_; t210 = (uint32_t)w_;
    w_ = (uint64_t)t203 + t212; t220 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t204 + t213 + cf_; t221 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t205 + t214 + cf_; t222 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t206 + t215 + cf_; t223 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t207 + t216 + cf_; t224 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t208 + t217 + cf_; t225 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t209 + t218 + cf_; t226 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t193 * 0xee00bc4fu) + t226; t227 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t192 * 0xccd1c8aau) + t227; t228 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t191 * 0x7d74d2e4u) + t228; t229 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t190 * 0x48c94408u) + t229; t230 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t189 * 0xc588c6f6u) + t230; t231 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t188 * 0x50fe77ecu) + t231; t232 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t187 * 0xa9d6281cu) + t232; t233 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t155 * 0x60d06633u) + t233; t234 = (uint32_t)w_;
    t235 = (uint32_t)((uint32_t)(t202 * 0xfc632551u));
    t236 = (uint32_t)(((uint64_t)t202 * 0xfc632551u) >> 32);
    t237 = (uint32_t)((uint32_t)(t221 * 0xfc632551u));
    t238 = (uint32_t)(((uint64_t)t221 * 0xfc632551u) >> 32);
    t239 = (uint32_t)((uint32_t)(t223 * 0xfc632551u));
    t240 = (uint32_t)(((uint64_t)t223 * 0xfc632551u) >> 32);
    t241 = (uint32_t)((uint32_t)(t225 * 0xfc632551u));
    t242 = (uint32_t)(((uint64_t)t225 * 0xfc632551u) >> 32);
    t252 = (uint32_t)((uint32_t)(t220 * 0xfc632551u));
    t253 = (uint32_t)(((uint64_t)t220 * 0xfc632551u) >> 32);
    t254 = (uint32_t)((uint32_t)(t222 * 0xfc632551u));
    t255 = (uint32_t)(((uint64_t)t222 * 0xfc632551u) >> 32);
    t256 = (uint32_t)((uint32_t)(t224 * 0xfc632551u));
    t257 = (uint32_t)(((uint64_t)t224 * 0xfc632551u) >> 32);
    t258 = (uint32_t)((uint32_t)(t234 * 0xfc632551u));
    t259 = (uint32_t)(((uint64_t)t234 * 0xfc632551u) >> 32);
    w_ = (uint64_t)(uint32_t)(t220 * 0xf3b9cac2u) + t237; t237 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xf3b9cac2u) >> 32) + t238 + cf_; t238 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xf3b9cac2u) + t239 + cf_; t239 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xf3b9cac2u) >> 32) + t240 + cf_; t240 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xf3b9cac2u) + t241 + cf_; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xf3b9cac2u) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xf3b9cac2u) + 0x0u + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xf3b9cac2u) >> 32) + 0x0u + cf_; t244 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xf3b9cac2u) + t252; t252 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xf3b9cac2u) >> 32) + t253 + cf_; t253 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xf3b9cac2u) + t254 + cf_; t254 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xf3b9cac2u) >> 32) + t255 + cf_; t255 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xf3b9cac2u) + t256 + cf_; t256 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xf3b9cac2u) >> 32) + t257 + cf_; t257 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xf3b9cac2u) + t258 + cf_; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xf3b9cac2u) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t260 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xa7179e84u) + t237; t237 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xa7179e84u) >> 32) + t238 + cf_; t238 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xa7179e84u) + t239 + cf_; t239 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xa7179e84u) >> 32) + t240 + cf_; t240 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xa7179e84u) + t241 + cf_; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xa7179e84u) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xa7179e84u) + t243 + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xa7179e84u) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t245 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0xa7179e84u) + t254; t254 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xa7179e84u) >> 32) + t255 + cf_; t255 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xa7179e84u) + t256 + cf_; t256 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xa7179e84u) >> 32) + t257 + cf_; t257 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xa7179e84u) + t258 + cf_; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xa7179e84u) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xa7179e84u) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xa7179e84u) >> 32) + 0x0u + cf_; t261 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0xbce6faadu) + t239; t239 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xbce6faadu) >> 32) + t240 + cf_; t240 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xbce6faadu) + t241 + cf_; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xbce6faadu) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xbce6faadu) + t243 + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xbce6faadu) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xbce6faadu) + t245 + cf_; t245 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xbce6faadu) >> 32) + 0x0u + cf_; t246 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xbce6faadu) + t254; t254 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xbce6faadu) >> 32) + t255 + cf_; t255 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xbce6faadu) + t256 + cf_; t256 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xbce6faadu) >> 32) + t257 + cf_; t257 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xbce6faadu) + t258 + cf_; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xbce6faadu) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xbce6faadu) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xbce6faadu) >> 32) + t261 + cf_; t261 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t262 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xffffffffu) + t239; t239 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xffffffffu) >> 32) + t240 + cf_; t240 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xffffffffu) + t241 + cf_; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xffffffffu) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xffffffffu) + t243 + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xffffffffu) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xffffffffu) + t245 + cf_; t245 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xffffffffu) >> 32) + t246 + cf_; t246 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t247 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0xffffffffu) + t256; t256 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xffffffffu) >> 32) + t257 + cf_; t257 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xffffffffu) + t258 + cf_; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xffffffffu) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xffffffffu) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xffffffffu) >> 32) + t261 + cf_; t261 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xffffffffu) + t262 + cf_; t262 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xffffffffu) >> 32) + 0x0u + cf_; t263 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0xffffffffu) + t241; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xffffffffu) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xffffffffu) + t243 + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xffffffffu) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xffffffffu) + t245 + cf_; t245 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xffffffffu) >> 32) + t246 + cf_; t246 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xffffffffu) + t247 + cf_; t247 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xffffffffu) >> 32) + 0x0u + cf_; t248 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xffffffffu) + t256; t256 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xffffffffu) >> 32) + t257 + cf_; t257 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xffffffffu) + t258 + cf_; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xffffffffu) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xffffffffu) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xffffffffu) >> 32) + t261 + cf_; t261 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xffffffffu) + t262 + cf_; t262 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xffffffffu) >> 32) + t263 + cf_; t263 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t264 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0x0u) + t241; t241 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0x0u) >> 32) + t242 + cf_; t242 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0x0u) + t243 + cf_; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0x0u) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0x0u) + t245 + cf_; t245 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0x0u) >> 32) + t246 + cf_; t246 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0x0u) + t247 + cf_; t247 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0x0u) >> 32) + t248 + cf_; t248 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t249 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0x0u) + t258; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0x0u) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0x0u) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0x0u) >> 32) + t261 + cf_; t261 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0x0u) + t262 + cf_; t262 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0x0u) >> 32) + t263 + cf_; t263 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0x0u) + t264 + cf_; t264 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0x0u) >> 32) + 0x0u + cf_; t265 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t220 * 0xffffffffu) + t243; t243 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t220 * 0xffffffffu) >> 32) + t244 + cf_; t244 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t222 * 0xffffffffu) + t245 + cf_; t245 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t222 * 0xffffffffu) >> 32) + t246 + cf_; t246 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t224 * 0xffffffffu) + t247 + cf_; t247 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t224 * 0xffffffffu) >> 32) + t248 + cf_; t248 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t234 * 0xffffffffu) + t249 + cf_; t249 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t234 * 0xffffffffu) >> 32) + 0x0u + cf_; t250 = (uint32_t)w_;
    w_ = (uint64_t)(uint32_t)(t202 * 0xffffffffu) + t258; t258 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t202 * 0xffffffffu) >> 32) + t259 + cf_; t259 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t221 * 0xffffffffu) + t260 + cf_; t260 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t221 * 0xffffffffu) >> 32) + t261 + cf_; t261 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t223 * 0xffffffffu) + t262 + cf_; t262 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t223 * 0xffffffffu) >> 32) + t263 + cf_; t263 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)(uint32_t)(t225 * 0xffffffffu) + t264 + cf_; t264 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (((uint64_t)t225 * 0xffffffffu) >> 32) + t265 + cf_; t265 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t266 = (uint32_t)w_;
    w_ = (uint64_t)t236 + t252; t267 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t237 + t253 + cf_; t268 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t238 + t254 + cf_; t269 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t239 + t255 + cf_; t270 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t240 + t256 + cf_; t271 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t241 + t257 + cf_; t272 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t242 + t258 + cf_; t273 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t243 + t259 + cf_; t274 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t244 + t260 + cf_; t275 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t245 + t261 + cf_; t276 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t246 + t262 + cf_; t277 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t247 + t263 + cf_; t278 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t248 + t264 + cf_; t279 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t249 + t265 + cf_; t280 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t250 + t266 + cf_; t281 = (uint32_t)w_;
    w_ = (uint64_t)t155 + t235; t282 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t187 + t267 + cf_; t283 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t188 + t268 + cf_; t284 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t189 + t269 + cf_; t285 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t190 + t270 + cf_; t286 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t191 + t271 + cf_; t287 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t192 + t272 + cf_; t288 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t193 + t273 + cf_; t289 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t194 + t274 + cf_; t290 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t195 + t275 + cf_; t291 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t196 + t276 + cf_; t292 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t197 + t277 + cf_; t293 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t198 + t278 + cf_; t294 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t199 + t279 + cf_; t295 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t200 + t280 + cf_; t296 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t201 + t281 + cf_; t297 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t298 = (uint32_t)w_;
    w_ = (uint64_t)t290 - 0xfc632551u; t299 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t291 - 0xf3b9cac2u - cf_; t300 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t292 - 0xa7179e84u - cf_; t301 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t293 - 0xbce6faadu - cf_; t302 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t294 - 0xffffffffu - cf_; t303 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t295 - 0xffffffffu - cf_; t304 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t296 - 0x0u - cf_; t305 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t297 - 0xffffffffu - cf_; t306 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t298 - 0x0u - cf_; t307 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t308 = (uint32_t)w_;
    t309 = (uint32_t)(t299 ^ t290);
    t310 = (uint32_t)(t309 & t308);
    t311 = (uint32_t)(t310 ^ t299);
    t312 = (uint32_t)(t300 ^ t291);
    t313 = (uint32_t)(t312 & t308);
    t314 = (uint32_t)(t313 ^ t300);
    t315 = (uint32_t)(t301 ^ t292);
    t316 = (uint32_t)(t315 & t308);
    t317 = (uint32_t)(t316 ^ t301);
    t318 = (uint32_t)(t302 ^ t293);
    t319 = (uint32_t)(t318 & t308);
    t320 = (uint32_t)(t319 ^ t302);
    t321 = (uint32_t)(t303 ^ t294);
    t322 = (uint32_t)(t321 & t308);
    t323 = (uint32_t)(t322 ^ t303);
    t324 = (uint32_t)(t304 ^ t295);
    t325 = (uint32_t)(t324 & t308);
    t326 = (uint32_t)(t325 ^ t304);
    t327 = (uint32_t)(t305 ^ t296);
    t328 = (uint32_t)(t327 & t308);
    t329 = (uint32_t)(t328 ^ t305);
    t330 = (uint32_t)(t306 ^ t297);
    t331 = (uint32_t)(t330 & t308);
    t332 = (uint32_t)(t331 ^ t306);
    w_ = (uint64_t)t311 + c_0_i; t333 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t314 + c_1_i + cf_; t334 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t317 + c_2_i + cf_; t335 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t320 + c_3_i + cf_; t336 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t323 + c_4_i + cf_; t337 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t326 + c_5_i + cf_; t338 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t329 + c_6_i + cf_; t339 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)t332 + c_7_i + cf_; t340 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 32);
    w_ = (uint64_t)0x0u + 0x0u + cf_; t341 = (uint32_t)w_;
    w_ = (uint64_t)t333 - 0xfc632551u; t342 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t334 - 0xf3b9cac2u - cf_; t343 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t335 - 0xa7179e84u - cf_; t344 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t336 - 0xbce6faadu - cf_; t345 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t337 - 0xffffffffu - cf_; t346 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t338 - 0xffffffffu - cf_; t347 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t339 - 0x0u - cf_; t348 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t340 - 0xffffffffu - cf_; t349 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)t341 - 0x0u - cf_; t350 = (uint32_t)w_; cf_ = (uint32_t)(w_ >> 63);
    w_ = (uint64_t)0x0u - 0x0u - cf_; t351 = (uint32_t)w_;
    t352 = (uint32_t)(t342 ^ t333);
    t353 = (uint32_t)(t352 & t351);
    t354 = (uint32_t)(t353 ^ t342);
    t355 = (uint32_t)(t343 ^ t334);
    t356 = (uint32_t)(t355 & t351);
    t357 = (uint32_t)(t356 ^ t343);
    t358 = (uint32_t)(t344 ^ t335);
    t359 = (uint32_t)(t358 & t351);
    t360 = (uint32_t)(t359 ^ t344);
    t361 = (uint32_t)(t345 ^ t336);
    t362 = (uint32_t)(t361 & t351);
    t363 = (uint32_t)(t362 ^ t345);
    t364 = (uint32_t)(t346 ^ t337);
    t365 = (uint32_t)(t364 & t351);
    t366 = (uint32_t)(t365 ^ t346);
    t367 = (uint32_t)(t347 ^ t338);
    t368 = (uint32_t)(t367 & t351);
    t369 = (uint32_t)(t368 ^ t347);
    t370 = (uint32_t)(t348 ^ t339);
    t371 = (uint32_t)(t370 & t351);
    t372 = (uint32_t)(t371 ^ t348);
    t373 = (uint32_t)(t349 ^ t340);
    t374 = (uint32_t)(t373 & t351);
    t375 = (uint32_t)(t374 ^ t349);
    r[0] = t354;
    r[1] = t357;
    r[2] = t360;
    r[3] = t363;
    r[4] = t366;
    r[5] = t369;
    r[6] = t372;
    r[7] = t375;
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
    sqr(t0, x);
    mul(t0, t0, x);
    sqr(t1, t0);
    sqr(t1, t1);
    mul(t1, t1, t0);
    sqr(t2, t1);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr(t2, t2);
    mul(t2, t2, t1);
    sqr(t3, t2);
    MAB_NOUNROLL
    for (int i = 1; i < 8; i++) sqr(t3, t3);
    mul(t3, t3, t2);
    sqr(t2, t3);
    MAB_NOUNROLL
    for (int i = 1; i < 16; i++) sqr(t2, t2);
    mul(t2, t2, t3);
    sqr(z, t2);
    MAB_NOUNROLL
    for (int i = 1; i < 64; i++) sqr(z, z);
    mul(z, z, t2);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 32; i++) sqr(z, z);
    mul(z, z, t2);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr(z, z);
    mul(z, z, t1);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr(z, z);
    mul(z, z, t1);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr(z, z);
    mul(z, z, t1);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 6; i++) sqr(z, z);
    mul(z, z, t1);
    sqr(z, z);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 6; i++) sqr(z, z);
    mul(z, z, t1);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 4; i++) sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr(z, z);
    mul(z, z, t1);
    sqr(z, z);
    sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 5; i++) sqr(z, z);
    mul(z, z, t0);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    MAB_NOUNROLL
    for (int i = 1; i < 3; i++) sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
    sqr(z, z);
    mul(z, z, x);
    sqr(z, z);
  }
};
