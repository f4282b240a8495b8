// C ABI instantiation for NIST256ORDER: generated field code (full-Montgomery fall-back plan) + kernels.
#include "gen/field_NIST256ORDER.cuh"
#define MAB_P NIST256ORDER
#define MAB_F F_NIST256ORDER
#define MAB_JIT_SRC "jit_src_NIST256ORDER.inc"
#include "mab_capi.inc"
