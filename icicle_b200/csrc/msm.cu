// MSM entry points of the C ABI; dispatches to the per-curve translation units (msm_inst.cu, one object per curve).
// Curves that were not built into this library (make CURVES=...) report API_NOT_IMPLEMENTED.
#include "../../include/icicle_b200.h"
#include <cstring>
#include <cstdint>

#define DECL(ID)                                                                                                       \
  extern "C" int b200_msm_entry_##ID(const void*, const void*, int, const b200_msm_config*, void*) __attribute__((weak)); \
  extern "C" int b200_msm_precompute_entry_##ID(const void*, int, const b200_msm_config*, void*) __attribute__((weak)); \
  extern "C" int b200_msm_plan_c_entry_##ID(int, const b200_msm_config*) __attribute__((weak));                         \
  extern "C" int b200_msm_plan_levels_entry_##ID(int, const b200_msm_config*) __attribute__((weak));                    \
  extern "C" int b200_ec_sum_entry_##ID(const void*, int, const b200_vec_ops_config*, void*) __attribute__((weak));       \
  extern "C" int b200_ecntt_entry_##ID(const void*, int, int, const b200_ntt_config*, void*, const uint32_t*, const uint32_t*, int) __attribute__((weak));
DECL(0) DECL(1) DECL(2) DECL(3) DECL(4) DECL(5) DECL(6) DECL(8)

#define CASE(ID, TU)                                                                                                   \
  case ID:                                                                                                             \
    if (!b200_msm_entry_##TU) return B200_API_NOT_IMPLEMENTED;
#define ALL_CASES(CALL)                                                                                                \
  switch (curve) {                                                                                                     \
    CASE(0, 0) CALL(0);                                                                                                \
    CASE(1, 1) CALL(1);                                                                                                \
    CASE(2, 2) CALL(2);                                                                                                \
    CASE(3, 3) CALL(3);                                                                                                \
    CASE(4, 4) CALL(4);                                                                                                \
    CASE(5, 5) CALL(5);                                                                                                \
    CASE(6, 6) CALL(6);                                                                                                \
    CASE(7, 6) CALL(6); /* bw6-761 G2 shares the G1 instantiation (same base field, b-free formulas) */                \
    CASE(8, 8) CALL(8);                                                                                                \
  default:                                                                                                             \
    return B200_INVALID_ARGUMENT;                                                                                      \
  }

extern "C" {

__attribute__((visibility("default"))) void b200_msm_default_config(b200_msm_config* cfg)
{
  // default_msm_config(): icicle/include/icicle/msm.h:60-78
  memset(cfg, 0, sizeof(*cfg));
  cfg->precompute_factor = 1;
  cfg->batch_size = 1;
  cfg->are_points_shared_in_batch = 1;
}

__attribute__((visibility("default"))) int
b200_msm(int curve, const void* scalars, const void* bases, int msm_size, const b200_msm_config* cfg, void* results)
{
  if (!cfg || !scalars || !bases || !results) return B200_INVALID_POINTER;
#define CALL(TU) return b200_msm_entry_##TU(scalars, bases, msm_size, cfg, results)
  ALL_CASES(CALL)
#undef CALL
}

__attribute__((visibility("default"))) int
b200_msm_precompute_bases(int curve, const void* input_bases, int nof_bases, const b200_msm_config* cfg, void* output_bases)
{
  if (!cfg || !input_bases || !output_bases) return B200_INVALID_POINTER;
#define CALL(TU) return b200_msm_precompute_entry_##TU(input_bases, nof_bases, cfg, output_bases)
  ALL_CASES(CALL)
#undef CALL
}

__attribute__((visibility("default"))) int
b200_ec_sum(int curve, const void* points, int n, const b200_vec_ops_config* cfg, void* out)
{
  if (!cfg || !points || !out) return B200_INVALID_POINTER;
#define CALL(TU) return b200_ec_sum_entry_##TU(points, n, cfg, out)
  ALL_CASES(CALL)
#undef CALL
}

// scalar field whose NTT domain supplies the ECNTT twiddles (curve_config.h: scalar_t of each curve)
static int curve_scalar_field(int curve)
{
  switch (curve) {
  case B200_CURVE_BN254_G1: return B200_FIELD_BN254_FR;
  case B200_CURVE_BLS12_381_G1: return B200_FIELD_BLS12_381_FR;
  case B200_CURVE_BLS12_377_G1: return B200_FIELD_BLS12_377_FR;
  case B200_CURVE_BW6_761_G1: return B200_FIELD_BLS12_377_FQ;
  default: return -1;
  }
}
extern "C" int b200_internal_ntt_domain(int field, const uint32_t** tw, const uint32_t** aux, int* max_log); // ntt.cu

__attribute__((visibility("default"))) int
b200_ecntt(int curve, const void* input, int size, int dir, const b200_ntt_config* cfg, void* output)
{
  if (!cfg || !input || !output) return B200_INVALID_POINTER;
  const int field = curve_scalar_field(curve);
  if (field < 0) return B200_API_NOT_IMPLEMENTED;
  const uint32_t *tw = nullptr, *aux = nullptr;
  int dom_log = 0;
  int err = b200_internal_ntt_domain(field, &tw, &aux, &dom_log);
  if (err) return err;
#define CALL(TU) return b200_ecntt_entry_##TU(input, size, dir, cfg, output, tw, aux, dom_log)
  ALL_CASES(CALL)
#undef CALL
}

__attribute__((visibility("default"))) int b200_msm_choose_c(int curve, int msm_size, const b200_msm_config* cfg)
{
  if (!cfg) return -1;
#define CALL(TU) return b200_msm_plan_c_entry_##TU(msm_size, cfg)
  ALL_CASES(CALL)
#undef CALL
}

__attribute__((visibility("default"))) int b200_msm_pair_levels(int curve, int msm_size, const b200_msm_config* cfg)
{
  if (!cfg) return -1;
#define CALL(TU) return b200_msm_plan_levels_entry_##TU(msm_size, cfg)
  ALL_CASES(CALL)
#undef CALL
}

} // extern "C"
