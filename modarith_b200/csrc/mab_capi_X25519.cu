// C ABI instantiation for X25519: generated field code + hand-written kernels.
#include "gen/field_X25519.cuh"
#define MAB_P X25519
#define MAB_F F_X25519
#define MAB_HAS_CURVE 1
#define MAB_HAS_EDWARDS 1
#define MAB_JIT_SRC "jit_src_X25519.inc"
#include "mab_capi.inc"
