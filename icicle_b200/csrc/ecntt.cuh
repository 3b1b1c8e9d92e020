// ECNTT: the NTT over curve points (out[k] = sum_i w^(ik) * P_i, scalar-field twiddles acting on G1 points) for sm_100a.
//
// Replaces (reference): ECNttFieldImpl (icicle/include/icicle/backend/ecntt_backend.h:15-22), frontend
// icicle/src/ecntt.cpp:5-18, CPU implementation = ntt_cpu::cpu_ntt<scalar_t, projective_t>
// (icicle/backend/cpu/src/curve/cpu_ecntt.cpp:12-19 -> backend/cpu/include/ntt_cpu.h:69-232), i.e. exactly the field NTT's
// definition with `E * S` = point x scalar: forward out[k] = sum_i (g^i * P_i) w^(ik); inverse out[i] = g^-i N^-1 sum_k P_k w^(-ik);
// kRN/kRR read bit-reversed input, kNR/kRR write bit-reversed output; kNM/kMN are handled like kNR/kRN (the "M" order is
// backend-private); batch_size / columns_batch as in NTTConfig.  Outputs are the same GROUP ELEMENTS as the
// reference's (its tests compare with projective equality, tests/test_curve_api.cpp:293-400), not the same representatives.
//
// Schedule: one butterfly = one scalar multiplication (~1.5 * |r| group operations of ~14 Montgomery products each), so
// the transform is integer-multiply bound by three orders of magnitude over its memory traffic and needs no tiling:
//   k_ecntt_load   projective (standard form) -> XYZZ (Montgomery), optional bit-reversed gather, optional coset g^i * P_i
//   k_ecntt_stage  log2(N) decimation-in-frequency stages, thread per butterfly: (a, b) -> (a + b, w * (a - b)); twiddles
//                  come from the scalar field's NTT domain (ntt.cu), converted out of Montgomery form and consumed by a
//                  left-to-right double-and-add; w = 1 butterflies skip the multiplication
//   k_ecntt_store  XYZZ -> projective standard form in the requested order, with N^-1 (and g^-i) for the inverse
#pragma once
#include "msm_impl.cuh"

namespace b200 { namespace ecntt {

using msm::load_xyzz;
using msm::store_xyzz;

struct Params {
  const uint32_t* tw;  // scalar-field domain: tw[i] = root^i (Montgomery form), i < 2^dom_log
  const uint32_t* aux; // aux[k] = 2^-k (Montgomery form)
  uint32_t dom_log, n_log, batch;
  uint64_t bstride, estride; // element (b, i) lives at b*bstride + i*estride
  int inverse;
};

// k * p by left-to-right double-and-add; k in standard form
template <class S, class F>
__device__ XYZZ<F> ec_smul(const XYZZ<F>& p, const S& k)
{
  XYZZ<F> acc = XYZZ<F>::inf();
  if (p.is_inf()) return acc;
  int top = -1;
#pragma unroll 1
  for (int i = S::N - 1; i >= 0; i--) {
    if (k.v[i]) {
      top = i * 32 + 31 - __clz(k.v[i]);
      break;
    }
  }
#pragma unroll 1
  for (int b = top; b >= 0; b--) {
    acc = acc.dbl();
    if ((k.v[b >> 5] >> (b & 31)) & 1u) acc.add(p);
  }
  return acc;
}

template <class S>
__device__ S pow_u64(S base, uint64_t e) // Montgomery in, Montgomery out
{
  S r = S::one();
  while (e) {
    if (e & 1ull) r = r * base;
    base = base * base;
    e >>= 1;
  }
  return r;
}

// gm[0] = g (forward) or g^-1 (inverse), Montgomery form
template <class S>
__global__ void k_ecntt_coset_setup(const uint32_t* g_std, int inverse, uint32_t* gm)
{
  S g = load_fp<S>(g_std).to_mont();
  if (inverse) g = msm::inv_fp(g);
  store_fp<S>(gm, g);
}

__device__ __forceinline__ void split_index(uint64_t t, const Params& p, bool columns, uint64_t& b, uint64_t& i)
{
  const uint64_t n = 1ull << p.n_log;
  if (columns) {
    b = t % p.batch;
    i = t / p.batch;
  } else {
    b = t / n;
    i = t % n;
  }
}

template <class S, class F>
__global__ void __launch_bounds__(msm::MSM_THREADS) k_ecntt_load(
  const uint32_t* __restrict__ in, uint32_t* __restrict__ work, Params p, bool columns, bool gather_in, const uint32_t* __restrict__ gm)
{
  constexpr int PW = 3 * F::N, XW = 4 * F::N;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ((uint64_t)p.batch << p.n_log)) return;
  uint64_t b, i;
  split_index(t, p, columns, b, i);
  const uint64_t isrc = (gather_in && p.n_log) ? (__brevll(i) >> (64 - p.n_log)) : i; // logical element i sits at rev(i)
  const uint32_t* src = in + (b * p.bstride + isrc * p.estride) * PW;
  Projective<F> pr = {load_el<F>(src).to_mont(), load_el<F>(src + F::N).to_mont(), load_el<F>(src + 2 * F::N).to_mont()};
  XYZZ<F> q = XYZZ<F>::from_projective(pr);
  if (gm && !p.inverse && i > 0) { // forward coset: P_i <- g^i * P_i (ntt_cpu.h:73)
    const S e = pow_u64(load_fp<S>(gm), i).from_mont();
    q = ec_smul<S, F>(q, e);
  }
  store_xyzz<F>(work + (b * p.bstride + i * p.estride) * XW, q);
}

// decimation-in-frequency stage t (pairs at distance 2^t), natural positions in, bit-reversed frequencies out after stage 0
template <class S, class F>
__global__ void __launch_bounds__(msm::MSM_THREADS) k_ecntt_stage(uint32_t* __restrict__ work, Params p, bool columns, uint32_t t)
{
  constexpr int XW = 4 * F::N;
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t half_n = 1ull << (p.n_log - 1);
  if (g >= half_n * p.batch) return;
  uint64_t b, j;
  if (columns) {
    b = g % p.batch;
    j = g / p.batch;
  } else {
    b = g / half_n;
    j = g % half_n;
  }
  const uint64_t lo = j & ((1ull << t) - 1);
  const uint64_t i0 = ((j >> t) << (t + 1)) | lo, i1 = i0 | (1ull << t);
  uint32_t* pa = work + (b * p.bstride + i0 * p.estride) * XW;
  uint32_t* pb = work + (b * p.bstride + i1 * p.estride) * XW;
  XYZZ<F> a = load_xyzz<F>(pa), c = load_xyzz<F>(pb);
  {
    XYZZ<F> sum = a;
    sum.add(c);
    store_xyzz<F>(pa, sum);
  }
  XYZZ<F> dif = a;
  dif.add(c.neg());
  if (lo != 0) {
    // w_{2^(t+1)}^lo = root^(lo << (dom_log - t - 1)); inverse transform: the conjugate power (ntt_task.h:1220-1222)
    uint64_t ex = lo << (p.dom_log - t - 1);
    if (p.inverse) ex = ((1ull << p.dom_log) - ex) & ((1ull << p.dom_log) - 1);
    const S w = load_fp<S>(p.tw + ex * S::N).from_mont();
    dif = ec_smul<S, F>(dif, w);
  }
  store_xyzz<F>(pb, dif);
}

template <class S, class F>
__global__ void __launch_bounds__(msm::MSM_THREADS) k_ecntt_store(
  const uint32_t* __restrict__ work, uint32_t* __restrict__ out, Params p, bool columns, bool scatter_out, const uint32_t* __restrict__ gm)
{
  constexpr int PW = 3 * F::N, XW = 4 * F::N;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ((uint64_t)p.batch << p.n_log)) return;
  uint64_t b, pos;
  split_index(t, p, columns, b, pos);
  const uint64_t k = p.n_log ? (__brevll(pos) >> (64 - p.n_log)) : 0; // position pos of the work array holds frequency k
  XYZZ<F> q = load_xyzz<F>(work + (b * p.bstride + pos * p.estride) * XW);
  if (p.inverse && p.n_log) { // N^-1 (ntt_task.h:1231-1235) and the inverse coset g^-k (ntt_cpu.h:226)
    S s = load_fp<S>(p.aux + (size_t)p.n_log * S::N);
    if (gm && k > 0) s = s * pow_u64(load_fp<S>(gm), k);
    q = ec_smul<S, F>(q, s.from_mont());
  }
  const Projective<F> pr = q.to_projective();
  uint32_t* o = out + (b * p.bstride + (scatter_out ? k : pos) * p.estride) * PW;
  store_el(o, pr.x.from_mont());
  store_el(o + F::N, pr.y.from_mont());
  store_el(o + 2 * F::N, pr.z.from_mont());
}

template <class C>
int ecntt_impl(const void* input, int size, int dir, const b200_ntt_config* cfg, void* output, const uint32_t* tw, const uint32_t* aux, int dom_log)
{
  typedef typename C::Scalar S;
  typedef typename C::Base F;
  constexpr int PW = 3 * F::N, XW = 4 * F::N;
  cudaStream_t s = (cudaStream_t)cfg->stream;
  if (size <= 0 || (size & (size - 1))) return B200_INVALID_ARGUMENT; // cpu_ntt_main.h:38
  int n_log = 0;
  while ((1 << n_log) < size) n_log++;
  if (!tw || n_log > dom_log) {
    fprintf(stderr, "[icicle_b200] ecntt: the scalar field's NTT domain is not initialised or smaller than 2^%d\n", n_log);
    return B200_INVALID_ARGUMENT; // cpu_ntt_main.h:39-41
  }
  const uint32_t batch = cfg->batch_size > 0 ? cfg->batch_size : 1;
  const uint64_t total = (uint64_t)size * batch;
  const size_t bytes = total * PW * 4;
  const int ord = cfg->ordering;
  const bool gather_in = (ord == B200_RN || ord == B200_RR || ord == B200_MN);
  const bool scatter_out = (ord == B200_NN || ord == B200_RN || ord == B200_MN);
  const bool columns = cfg->columns_batch != 0;

  Scratch sin, sout, swork, sg_in, sg;
  const void* din;
  void* dout;
  int err;
  if ((err = stage_in(din, input, bytes, cfg->are_inputs_on_device, s, sin))) return err;
  if ((err = stage_out(dout, output, bytes, cfg->are_outputs_on_device, s, sout))) return err;
  if ((err = swork.alloc(total * XW * 4, s))) return err;

  Params p;
  p.tw = tw;
  p.aux = aux;
  p.dom_log = (uint32_t)dom_log;
  p.n_log = (uint32_t)n_log;
  p.batch = batch;
  p.bstride = columns ? 1 : (uint64_t)size;
  p.estride = columns ? batch : 1;
  p.inverse = (dir == B200_NTT_INVERSE);

  const uint32_t* gm = nullptr;
  if (cfg->coset_gen) {
    const uint32_t* g = (const uint32_t*)cfg->coset_gen;
    bool is_one = (g[0] == 1);
    for (int i = 1; i < S::N; i++) is_one = is_one && (g[i] == 0);
    if (!is_one) {
      if ((err = sg_in.alloc(S::BYTES, s))) return err;
      if ((err = sg.alloc(S::BYTES, s))) return err;
      B200_CUDA_TRY(cudaMemcpyAsync(sg_in.p, cfg->coset_gen, S::BYTES, cudaMemcpyHostToDevice, s), B200_COPY_FAILED);
      k_ecntt_coset_setup<S><<<1, 1, 0, s>>>(sg_in.as<uint32_t>(), p.inverse, sg.as<uint32_t>()); B200_LAUNCHED(1);
      gm = sg.as<uint32_t>();
    }
  }
  const unsigned T = msm::MSM_THREADS;
  const unsigned g_all = (unsigned)((total + T - 1) / T);
  k_ecntt_load<S, F><<<g_all, T, 0, s>>>((const uint32_t*)din, swork.as<uint32_t>(), p, columns, gather_in, gm); B200_LAUNCHED(1);
  B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
  if (n_log > 0) {
    const unsigned g_half = (unsigned)((total / 2 + T - 1) / T);
    for (int t = n_log - 1; t >= 0; t--) {
      k_ecntt_stage<S, F><<<g_half, T, 0, s>>>(swork.as<uint32_t>(), p, columns, (uint32_t)t); B200_LAUNCHED(1);
    }
    B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
  }
  k_ecntt_store<S, F><<<g_all, T, 0, s>>>(swork.as<uint32_t>(), (uint32_t*)dout, p, columns, scatter_out, gm); B200_LAUNCHED(1);
  B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
  return finish_out(output, dout, bytes, cfg->are_outputs_on_device, cfg->is_async, s);
}

}} // namespace b200::ecntt
