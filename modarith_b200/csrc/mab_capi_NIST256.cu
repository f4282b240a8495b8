// C ABI instantiation for NIST256: generated field code + hand-written kernels.
#include "gen/field_NIST256.cuh"
#define MAB_P NIST256
#define MAB_F F_NIST256
#define MAB_HAS_WEIERSTRASS 1

#include "mab_capi.inc"
