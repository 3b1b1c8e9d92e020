// One translation unit per curve group (compiled with -DB200_MSM_CURVE=<b200_curve_t value>) so the heavy templates
// build in parallel.  The entry points are looked up by icicle_b200/csrc/msm.cu.
#include "msm_impl.cuh"
#include <type_traits>
#include "ecntt.cuh"

#define B200_CAT2(a, b) a##b
#define B200_CAT(a, b) B200_CAT2(a, b)

namespace b200 { namespace msm {
template <int ID> struct CurveOf;
template <> struct CurveOf<B200_CURVE_BN254_G1> { typedef CurveT<params::bn254_fr, Fp<params::bn254_fq>> type; };
template <> struct CurveOf<B200_CURVE_BN254_G2> { typedef CurveT<params::bn254_fr, Fp2<params::bn254_fq>> type; };
template <> struct CurveOf<B200_CURVE_BLS12_381_G1> { typedef CurveT<params::bls12_381_fr, Fp<params::bls12_381_fq>> type; };
template <> struct CurveOf<B200_CURVE_BLS12_381_G2> { typedef CurveT<params::bls12_381_fr, Fp2<params::bls12_381_fq>> type; };
template <> struct CurveOf<B200_CURVE_BLS12_377_G1> { typedef CurveT<params::bls12_377_fr, Fp<params::bls12_377_fq>> type; };
template <> struct CurveOf<B200_CURVE_BLS12_377_G2> { typedef CurveT<params::bls12_377_fr, Fp2<params::bls12_377_fq>> type; };
// bw6-761: G1 and G2 live over the same base field (curves/params/bw6_761.h:15-18) and the XYZZ formulas do not
// involve the curve constant b, so one instantiation serves both.
template <> struct CurveOf<B200_CURVE_BW6_761_G1> { typedef CurveT<params::bls12_377_fq, Fp<params::bw6_761_fq>> type; };
template <> struct CurveOf<B200_CURVE_GRUMPKIN> { typedef CurveT<params::bn254_fq, Fp<params::bn254_fr>> type; };
}} // namespace b200::msm

using namespace b200;
using namespace b200::msm;
typedef CurveOf<B200_MSM_CURVE>::type ThisCurve;

// the reference registers ECNTT for G1 only (ecntt_backend.h:15-22: projective_t); the Fq2 curves do not instantiate it
template <class C, bool G1 = std::is_same<typename C::Base, typename base_fp<typename C::Base>::type>::value>
struct EcnttEntry {
  static int run(const void* in, int size, int dir, const b200_ntt_config* cfg, void* out, const uint32_t* tw, const uint32_t* aux, int dom_log)
  {
    return ecntt::ecntt_impl<C>(in, size, dir, cfg, out, tw, aux, dom_log);
  }
};
template <class C>
struct EcnttEntry<C, false> {
  static int run(const void*, int, int, const b200_ntt_config*, void*, const uint32_t*, const uint32_t*, int) { return B200_API_NOT_IMPLEMENTED; }
};

extern "C" {
__attribute__((visibility("hidden"))) int B200_CAT(b200_msm_entry_, B200_MSM_CURVE)(
  const void* scalars, const void* bases, int msm_size, const b200_msm_config* cfg, void* results)
{
  return msm_impl<ThisCurve>(scalars, bases, msm_size, cfg, results);
}
__attribute__((visibility("hidden"))) int B200_CAT(b200_msm_precompute_entry_, B200_MSM_CURVE)(
  const void* in, int n, const b200_msm_config* cfg, void* out)
{
  return precompute_impl<ThisCurve>(in, n, cfg, out);
}
__attribute__((visibility("hidden"))) int B200_CAT(b200_ec_sum_entry_, B200_MSM_CURVE)(
  const void* points, int n, const b200_vec_ops_config* cfg, void* out)
{
  return ec_sum_impl<ThisCurve>(points, n, cfg, out);
}
__attribute__((visibility("hidden"))) int B200_CAT(b200_ecntt_entry_, B200_MSM_CURVE)(
  const void* input, int size, int dir, const b200_ntt_config* cfg, void* output, const uint32_t* tw, const uint32_t* aux, int dom_log)
{
  return EcnttEntry<ThisCurve>::run(input, size, dir, cfg, output, tw, aux, dom_log);
}
#if B200_MSM_CURVE == 0
// planning query (host only, curve independent): the chunk sizes the host-pointer pipeline would use for an msm of `msm_size` points
__attribute__((visibility("default"))) int b200_msm_pipeline_schedule(int msm_size, uint32_t* sizes, int max_chunks)
{
  if (msm_size <= 0 || !sizes || max_chunks <= 0) return -1;
  return (int)pipeline_schedule((uint32_t)msm_size, sizes, (uint32_t)std::min(max_chunks, 32));
}
#endif
__attribute__((visibility("hidden"))) int B200_CAT(b200_msm_plan_c_entry_, B200_MSM_CURVE)(int msm_size, const b200_msm_config* cfg)
{
  return make_plan<ThisCurve>(msm_size, cfg).c;
}
__attribute__((visibility("hidden"))) int B200_CAT(b200_msm_plan_levels_entry_, B200_MSM_CURVE)(int msm_size, const b200_msm_config* cfg)
{
  const MsmPlan pl = make_plan<ThisCurve>(msm_size, cfg);
  const int batch = cfg->batch_size > 0 ? cfg->batch_size : 1;
  const bool shared = cfg->are_points_shared_in_batch || batch == 1;
  const int chunk = batch_chunk((uint32_t)msm_size, pl, batch, shared, cfg, (size_t)ThisCurve::Base::BYTES);
  const uint64_t max_ent = (uint64_t)msm_size * pl.nwin * chunk;
  const uint64_t max_buckets = ((uint64_t)pl.nbm * chunk) << (pl.c - 1);
  return choose_pair_levels(max_ent, max_buckets);
}
}
