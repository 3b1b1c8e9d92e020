// libicicle_backend_cuda_curve_<curve>.so : MSM (+G2), precompute-bases and curve Montgomery-conversion registrations.
// Hooks used: REGISTER_MSM_BACKEND / REGISTER_MSM_PRE_COMPUTE_BASES_BACKEND / REGISTER_MSM_G2_BACKEND /
// REGISTER_MSM_G2_PRE_COMPUTE_BASES_BACKEND (icicle/include/icicle/backend/msm_backend.h:21,38,57,74) and
// REGISTER_{AFFINE,PROJECTIVE}[_G2]_CONVERT_MONTGOMERY_BACKEND (icicle/include/icicle/curves/montgomery_conversion.h:27,45,...),
// REGISTER_ECNTT_BACKEND (icicle/include/icicle/backend/ecntt_backend.h:26-33) on curves whose scalar field has an NTT.
#include "shim_common.h"
#include "icicle/msm.h"
#include "icicle/vec_ops.h"
#include "icicle/backend/msm_backend.h"
#include "icicle/backend/msm_config.h"
#include "icicle/curves/curve_config.h"
#include "icicle/curves/montgomery_conversion.h"
#ifdef ECNTT
  #include "icicle/ntt.h"
  #include "icicle/backend/ntt_config.h"
  #include "icicle/backend/ecntt_backend.h"
#endif

using namespace icicle;
using namespace curve_config;
using namespace b200_shim;

namespace {

  b200_msm_config to_c(const MSMConfig& c)
  {
    b200_msm_config o;
    b200_msm_default_config(&o);
    o.stream = c.stream;
    o.precompute_factor = c.precompute_factor;
    o.c = c.c;
    o.bitsize = c.bitsize;
    o.batch_size = c.batch_size;
    o.are_points_shared_in_batch = c.are_points_shared_in_batch;
    o.are_scalars_on_device = c.are_scalars_on_device;
    o.are_scalars_montgomery_form = c.are_scalars_montgomery_form;
    o.are_points_on_device = c.are_points_on_device;
    o.are_points_montgomery_form = c.are_points_montgomery_form;
    o.are_results_on_device = c.are_results_on_device;
    o.is_async = c.is_async;
    o.ext_large_bucket_factor = ext_int(c.ext, CudaBackendConfig::CUDA_MSM_LARGE_BUCKET_FACTOR, 0);
    o.ext_nof_chunks = ext_int(c.ext, CudaBackendConfig::CUDA_MSM_NOF_CHUNKS, 0);
    o.ext_is_big_triangle = ext_int(c.ext, CudaBackendConfig::CUDA_MSM_IS_BIG_TRIANGLE, 0);
    return o;
  }
  b200_vec_ops_config to_c(const VecOpsConfig& c)
  {
    b200_vec_ops_config o;
    b200_vec_ops_default_config(&o);
    o.stream = c.stream;
    o.is_a_on_device = c.is_a_on_device;
    o.is_result_on_device = c.is_result_on_device;
    o.is_async = c.is_async;
    return o;
  }

  template <int CURVE, class A, class P>
  eIcicleError msm_t(const Device&, const scalar_t* scalars, const A* bases, int msm_size, const MSMConfig& config, P* results)
  {
    b200_msm_config c = to_c(config);
    // opt-in ConfigExtension key "multi_gpu" = number of devices: host-resident inputs/outputs are sharded over
    // that many GPUs, one host thread each (the reference's own multi-device model, multi-device.md:32-36)
    const int multi = ext_int(config.ext, "multi_gpu", 0);
    if (multi > 1 && !c.are_scalars_on_device && !c.are_points_on_device && !c.are_results_on_device)
      return to_err(b200_msm_multi_gpu(CURVE, scalars, bases, msm_size, &c, results, multi, nullptr));
    return to_err(b200_msm(CURVE, scalars, bases, msm_size, &c, results));
  }
  template <int CURVE, class A>
  eIcicleError precompute_t(const Device&, const A* in, int n, const MSMConfig& config, A* out)
  {
    b200_msm_config c = to_c(config);
    // the reference's precompute takes the *output location* from are_points_on_device of the same config for the
    // CPU backend (cpu_msm.hpp:454-481 writes host memory); wrappers set are_results_on_device for device outputs.
    return to_err(b200_msm_precompute_bases(CURVE, in, n, &c, out));
  }
  template <int CURVE, class A>
  eIcicleError affine_mont_t(const Device&, const A* in, size_t n, bool is_into, const VecOpsConfig& config, A* out)
  {
    b200_vec_ops_config c = to_c(config);
    return to_err(b200_affine_convert_montgomery(CURVE, in, n, is_into, &c, out));
  }
  template <int CURVE, class P>
  eIcicleError projective_mont_t(const Device&, const P* in, size_t n, bool is_into, const VecOpsConfig& config, P* out)
  {
    b200_vec_ops_config c = to_c(config);
    return to_err(b200_projective_convert_montgomery(CURVE, in, n, is_into, &c, out));
  }

#ifdef ECNTT
  // ECNttFieldImpl (ecntt_backend.h:15-22): projective_t elements, scalar_t twiddles from the scalar field's NTT domain
  template <int CURVE>
  eIcicleError ecntt_t(const Device&, const projective_t* in, int size, NTTDir dir, const NTTConfig<scalar_t>& config, projective_t* out)
  {
    b200_ntt_config c;
    b200_ntt_default_config(&c);
    c.stream = config.stream;
    c.coset_gen = &config.coset_gen;
    c.batch_size = config.batch_size;
    c.columns_batch = config.columns_batch;
    c.are_inputs_on_device = config.are_inputs_on_device;
    c.are_outputs_on_device = config.are_outputs_on_device;
    c.is_async = config.is_async;
    c.ordering = static_cast<int>(config.ordering);
    return to_err(b200_ecntt(CURVE, in, size, dir == NTTDir::kForward ? B200_NTT_FORWARD : B200_NTT_INVERSE, &c, out));
  }
#endif

  constexpr int G1 = g1_curve_id();
  static_assert(G1 >= 0, "this curve has no B200 backend");

} // namespace

REGISTER_MSM_BACKEND(B200_DEVICE_TYPE, (msm_t<G1, affine_t, projective_t>));
REGISTER_MSM_PRE_COMPUTE_BASES_BACKEND(B200_DEVICE_TYPE, (precompute_t<G1, affine_t>));
REGISTER_AFFINE_CONVERT_MONTGOMERY_BACKEND(B200_DEVICE_TYPE, (affine_mont_t<G1, affine_t>));
REGISTER_PROJECTIVE_CONVERT_MONTGOMERY_BACKEND(B200_DEVICE_TYPE, (projective_mont_t<G1, projective_t>));
#ifdef ECNTT
REGISTER_ECNTT_BACKEND(B200_DEVICE_TYPE, (ecntt_t<G1>));
#endif
#ifdef G2_ENABLED
REGISTER_MSM_G2_BACKEND(B200_DEVICE_TYPE, (msm_t<G1 + 1, g2_affine_t, g2_projective_t>));
REGISTER_MSM_G2_PRE_COMPUTE_BASES_BACKEND(B200_DEVICE_TYPE, (precompute_t<G1 + 1, g2_affine_t>));
REGISTER_AFFINE_G2_CONVERT_MONTGOMERY_BACKEND(B200_DEVICE_TYPE, (affine_mont_t<G1 + 1, g2_affine_t>));
REGISTER_PROJECTIVE_G2_CONVERT_MONTGOMERY_BACKEND(B200_DEVICE_TYPE, (projective_mont_t<G1 + 1, g2_projective_t>));
#endif
