// C ABI instantiation for X448: generated field code + hand-written kernels.
#include "gen/field_X448.cuh"
#define MAB_P X448
#define MAB_F F_X448
#define MAB_HAS_CURVE 1
#define MAB_JIT_SRC "jit_src_X448.inc"
#include "mab_capi.inc"
