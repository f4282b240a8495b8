// C ABI instantiation for SECP256K1: generated field code (full-Montgomery fall-back plan) + kernels.
#include "gen/field_SECP256K1.cuh"
#define MAB_P SECP256K1
#define MAB_F F_SECP256K1
#define MAB_JIT_SRC "jit_src_SECP256K1.inc"
#include "mab_capi.inc"
