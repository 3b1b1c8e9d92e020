/*
 * icicle_b200 -- C ABI of the B200 (sm_100a) MSM / NTT / vec-ops engine.
 *
 * This is the drop-in boundary below ICICLE's backend-registration layer: every entry point here is what one of the
 * reference's per-device backend hooks binds to (the C++ registration shims under icicle_b200/shim/ are one-line
 * adapters).  Plain pointers and sizes only; no C++ / torch types.  All file:line citations are
 * relative to the reference tree (ingonyama-zk/icicle @ 625532a6).
 *
 * Data conventions (identical to the reference):
 *   - field element  = N little-endian uint32 limbs, canonical value in [0,p)      icicle/include/icicle/math/storage.h:36-48
 *   - affine point   = {x, y}, zero is (0,0)                                        icicle/include/icicle/curves/affine.h:11-39
 *   - projective     = homogeneous {X, Y, Z}, zero is (0,1,0)                       icicle/include/icicle/curves/projective.h:23-31
 *   - G2 coordinates = {real, imaginary} pairs of base-field elements               icicle/include/icicle/fields/complex_extension.h
 *   - "Montgomery form" means x*R mod p with R = 2^(32*N)                           icicle/include/icicle/fields/params_gen.h:35-50
 * Return value: 0 on success, otherwise the numeric value of the reference's eIcicleError
 * (icicle/include/icicle/errors.h:13-29).
 */
#ifndef ICICLE_B200_H
#define ICICLE_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes: numeric values of eIcicleError (errors.h:13-29) ---- */
enum {
  B200_SUCCESS = 0,
  B200_INVALID_DEVICE = 1,
  B200_OUT_OF_MEMORY = 2,
  B200_INVALID_POINTER = 3,
  B200_ALLOCATION_FAILED = 4,
  B200_DEALLOCATION_FAILED = 5,
  B200_COPY_FAILED = 6,
  B200_SYNCHRONIZATION_FAILED = 7,
  B200_STREAM_CREATION_FAILED = 8,
  B200_STREAM_DESTRUCTION_FAILED = 9,
  B200_API_NOT_IMPLEMENTED = 10,
  B200_INVALID_ARGUMENT = 11,
  B200_UNKNOWN_ERROR = 14
};

/* ---- fields (scalar / coefficient types).  limbs: bn254/bls/stark252 = 8, *_fq 381/377 = 12, bw6 = 24, bears = 1 ---- */
typedef enum {
  B200_FIELD_BN254_FR = 0,     /* bn254::scalar_t,     fields/snark_fields/bn254_scalar.h   (also grumpkin base field) */
  B200_FIELD_BN254_FQ = 1,     /* bn254 base field,    fields/snark_fields/bn254_base.h     (also grumpkin scalar field) */
  B200_FIELD_BLS12_381_FR = 2, /* fields/snark_fields/bls12_381_scalar.h */
  B200_FIELD_BLS12_381_FQ = 3, /* fields/snark_fields/bls12_381_base.h */
  B200_FIELD_BLS12_377_FR = 4, /* fields/snark_fields/bls12_377_scalar.h */
  B200_FIELD_BLS12_377_FQ = 5, /* fields/snark_fields/bls12_377_base.h     (also bw6_761 scalar field) */
  B200_FIELD_BW6_761_FQ = 6,   /* fields/snark_fields/bw6_761_base.h */
  B200_FIELD_STARK252 = 7,     /* fields/stark_fields/stark252.h */
  B200_FIELD_BABYBEAR = 8,     /* fields/stark_fields/babybear.h */
  B200_FIELD_KOALABEAR = 9,    /* fields/stark_fields/koalabear.h */
  B200_FIELD_M31 = 10,         /* fields/stark_fields/m31.h: vec-ops only (no NTT upstream); Montgomery form == standard form (m31.h:232-234) */
  B200_FIELD_GOLDILOCKS = 11,  /* fields/stark_fields/goldilocks.h: p = 2^64 - 2^32 + 1, 2 limbs, NTT + vec-ops */
  B200_FIELD_BABYBEAR_EXT4 = 12,  /* babybear::extension_t  = QuarticExtensionField (fields/quartic_extension.h), 4 limbs: vec-ops; */
  B200_FIELD_KOALABEAR_EXT4 = 13, /* koalabear::extension_t   its NTT is b200_ntt_extension on the BASE field id */
  B200_FIELD_COUNT
} b200_field_t;

/* ---- curve groups for MSM ---- */
typedef enum {
  B200_CURVE_BN254_G1 = 0,     /* curves/params/bn254.h */
  B200_CURVE_BN254_G2 = 1,
  B200_CURVE_BLS12_381_G1 = 2, /* curves/params/bls12_381.h */
  B200_CURVE_BLS12_381_G2 = 3,
  B200_CURVE_BLS12_377_G1 = 4, /* curves/params/bls12_377.h */
  B200_CURVE_BLS12_377_G2 = 5,
  B200_CURVE_BW6_761_G1 = 6,   /* curves/params/bw6_761.h (G2 is over the same base field) */
  B200_CURVE_BW6_761_G2 = 7,
  B200_CURVE_GRUMPKIN = 8,     /* curves/params/grumpkin.h */
  B200_CURVE_COUNT
} b200_curve_t;

/* ------------------------------------------------------------------------------------------------------------------
 * Device runtime -- what our DeviceAPI subclass forwards to (icicle/include/icicle/device_api.h:44-182; model:
 * icicle/backend/cpu/src/cpu_device_api.cpp).  `stream` is an opaque cudaStream_t (icicleStreamHandle, device_api.h:25).
 * ---------------------------------------------------------------------------------------------------------------- */
B200_API int b200_get_device_count(int* count);                    /* DeviceAPI::get_device_count        device_api.h:52  */
B200_API int b200_set_device(int device_id);                       /* DeviceAPI::set_device              device_api.h:46  */
B200_API int b200_malloc(void** ptr, size_t bytes);                /* DeviceAPI::allocate_memory         device_api.h:66  */
B200_API int b200_malloc_async(void** ptr, size_t bytes, void* stream);
B200_API int b200_free(void* ptr);                                 /* DeviceAPI::free_memory             device_api.h:84  */
B200_API int b200_free_async(void* ptr, void* stream);
B200_API int b200_get_available_memory(size_t* total, size_t* free_bytes); /* DeviceAPI::get_available_memory device_api.h:101 */
B200_API int b200_memset(void* ptr, int value, size_t bytes);      /* DeviceAPI::memset                  device_api.h:111 */
B200_API int b200_memset_async(void* ptr, int value, size_t bytes, void* stream);
B200_API int b200_copy_to_device(void* dst, const void* src, size_t bytes, void* stream, int is_async); /* copy / copy_async h2d  device_api.h:131-150 */
B200_API int b200_copy_to_host(void* dst, const void* src, size_t bytes, void* stream, int is_async);   /* d2h */
B200_API int b200_copy_device_to_device(void* dst, const void* src, size_t bytes, void* stream, int is_async); /* d2d */
B200_API int b200_synchronize(void* stream);                       /* DeviceAPI::synchronize (stream==NULL: whole device) device_api.h:158 */
B200_API int b200_create_stream(void** stream);                    /* DeviceAPI::create_stream           device_api.h:166 */
B200_API int b200_destroy_stream(void* stream);                    /* DeviceAPI::destroy_stream          device_api.h:173 */
/* pinned host staging (used by the e2e path and by bench.py; not part of the reference API) */
B200_API int b200_host_alloc_pinned(void** ptr, size_t bytes);
B200_API int b200_host_free_pinned(void* ptr);
/* size in bytes of one element of `field` / one affine or projective point of `curve` */
B200_API int b200_field_bytes(int field);
B200_API int b200_curve_scalar_field(int curve);
B200_API int b200_curve_affine_bytes(int curve);
B200_API int b200_curve_projective_bytes(int curve);

/* ------------------------------------------------------------------------------------------------------------------
 * MSM -- replaces MsmImpl / MsmPreComputeImpl (icicle/include/icicle/backend/msm_backend.h:11-17,29-34 and the G2
 * twins :47-53,65-70), i.e. cpu_msm / cpu_msm_precompute_bases (icicle/backend/cpu/src/curve/cpu_msm.hpp:430-481).
 * Field-for-field mirror of icicle::MSMConfig (icicle/include/icicle/msm.h:21-53) plus the backend extension keys the
 * closed CUDA backend reads from ConfigExtension (icicle/include/icicle/backend/msm_config.h:10-17), flattened.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
  void* stream;
  int precompute_factor;
  int c;                       /* 0 = choose automatically */
  int bitsize;                 /* 0 = scalar field bit size; otherwise scalars are taken mod 2^bitsize (cpu_msm.hpp:203,289) */
  int batch_size;
  uint8_t are_points_shared_in_batch;
  uint8_t are_scalars_on_device;
  uint8_t are_scalars_montgomery_form;
  uint8_t are_points_on_device;
  uint8_t are_points_montgomery_form;
  uint8_t are_results_on_device;
  uint8_t is_async;
  uint8_t reserved;
  int ext_large_bucket_factor; /* accepted, unused: work is split by fixed-size slices, not per bucket */
  int ext_nof_chunks;          /* 0 = auto: number of batch chunks processed at a time */
  int ext_is_big_triangle;     /* accepted, unused */
} b200_msm_config;

B200_API void b200_msm_default_config(b200_msm_config* cfg);      /* default_msm_config(), msm.h:60-78 */
/* results[b] = sum_i scalars[b*n+i] * bases[(shared ? 0 : b*n) + i]   (precompute: bases[pf*i + j]) */
B200_API int b200_msm(int curve, const void* scalars, const void* bases, int msm_size, const b200_msm_config* cfg, void* results);
/* out[pf*i + j] = 2^(j*shift) * in[i], affine; `shift` depends on (c, bitsize, pf) exactly as b200_msm expects. */
B200_API int b200_msm_precompute_bases(int curve, const void* input_bases, int nof_bases, const b200_msm_config* cfg, void* output_bases);
/* window size b200_msm would pick for this problem (exposed for the bench sweep and for tests) */
B200_API int b200_msm_choose_c(int curve, int msm_size, const b200_msm_config* cfg);
/* number of batched-affine pair levels the MSM schedule runs before the XYZZ bucket accumulation (0 = XYZZ only); our own
 * planning query, no reference counterpart (the reference has no such stage: cpu_msm.hpp:259-314 adds point by point) */
B200_API int b200_msm_pair_levels(int curve, int msm_size, const b200_msm_config* cfg);
/* chunk sizes (points) of the host-pointer copy/compute pipeline for an msm of msm_size points (host scalars and points,
 * batch 1, >= 2^23 points): returns the number of chunks written to sizes[0 .. max_chunks); our own planning query */
B200_API int b200_msm_pipeline_schedule(int msm_size, uint32_t* sizes, int max_chunks);

/* ------------------------------------------------------------------------------------------------------------------
 * NTT -- replaces NttImpl / NttInitDomainImpl / NttReleaseDomainImpl / NttGetRouFromDomainImpl
 * (icicle/include/icicle/backend/ntt_backend.h:13-19,52-53,68,81), i.e. cpu_ntt & CpuNttDomain
 * (icicle/backend/cpu/include/cpu_ntt_main.h:35-47, cpu_ntt_domain.h:63-110,613-654).
 * Mirror of icicle::NTTConfig<S> (icicle/include/icicle/ntt.h:52-64); coset_gen is passed by pointer because its size
 * depends on the field (NULL = one = no coset).
 * ---------------------------------------------------------------------------------------------------------------- */
enum { B200_NTT_FORWARD = 0, B200_NTT_INVERSE = 1 };                                   /* NTTDir,  ntt.h:23-26 */
enum { B200_NN = 0, B200_NR = 1, B200_RN = 2, B200_RR = 3, B200_NM = 4, B200_MN = 5 }; /* Ordering, ntt.h:37-44 */
enum { B200_NTT_ALG_AUTO = 0, B200_NTT_ALG_RADIX2 = 1, B200_NTT_ALG_MIXED_RADIX = 2 }; /* backend/ntt_config.h:7-18 */

typedef struct {
  void* stream;
  const void* coset_gen;       /* standard form, one field element; NULL = no coset */
  int batch_size;
  uint8_t columns_batch;
  uint8_t are_inputs_on_device;
  uint8_t are_outputs_on_device;
  uint8_t is_async;
  int ordering;
  int ext_ntt_algorithm;       /* CUDA_NTT_ALGORITHM extension key */
  int ext_fast_twiddles;       /* CUDA_NTT_FAST_TWIDDLES_MODE: accepted, unused */
} b200_ntt_config;

B200_API void b200_ntt_default_config(b200_ntt_config* cfg);
/* primitive_root must generate a subgroup of order 2^k; k becomes the domain's max_log_size (cpu_ntt_domain.h:78-94).
 * Idempotent if a domain already exists for (field, current device) (cpu_ntt_domain.h:69). */
B200_API int b200_ntt_init_domain(int field, const void* primitive_root, void* stream);
B200_API int b200_ntt_release_domain(int field);
B200_API int b200_ntt_get_root_of_unity_from_domain(int field, uint64_t logn, void* rou_out);
B200_API int b200_ntt(int field, const void* input, int size, int dir, const b200_ntt_config* cfg, void* output);
/* Extension-field NTT -- replaces NttExtFieldImpl (icicle/include/icicle/backend/ntt_backend.h:32-48, dispatcher
 * icicle/src/ntt.cpp:86-103; CPU: cpu_ntt<scalar_t, extension_t>, icicle/backend/cpu/src/field/cpu_ntt.cpp): `size` quartic
 * extension elements (4 base-field coefficients each, 16 B) per transform, BASE-field twiddles / coset generator / domain
 * (the scalar domain is reused); batch_size, columns_batch, ordering as for b200_ntt.  BabyBear and KoalaBear. */
B200_API int b200_ntt_extension(int field, const void* input, int size, int dir, const b200_ntt_config* cfg, void* output);
/* ECNTT -- replaces ECNttFieldImpl (icicle/include/icicle/backend/ecntt_backend.h:15-22, frontend icicle/src/ecntt.cpp:5-18;
 * CPU: ntt_cpu::cpu_ntt<scalar_t, projective_t>, icicle/backend/cpu/src/curve/cpu_ecntt.cpp:12-19): the NTT of `size` G1
 * points (homogeneous projective, standard form, 3*|Fq| bytes each) with the curve's SCALAR-field twiddles, i.e.
 * out[k] = sum_i w^(ik) * (g^i * P_i); the scalar field's domain must have been initialised with b200_ntt_init_domain.
 * `curve` is a G1 b200_curve_t of bn254 / bls12_381 / bls12_377 / bw6_761 (the reference's ECNTT feature list,
 * icicle/cmake/features.cmake:15-18).  Results are the reference's group elements (not its representatives). */
B200_API int b200_ecntt(int curve, const void* input, int size, int dir, const b200_ntt_config* cfg, void* output);

/* ------------------------------------------------------------------------------------------------------------------
 * vec-ops around the path -- replace the per-op hooks of icicle/include/icicle/backend/vec_ops_backend.h:11-83,85-270,
 * i.e. icicle/backend/cpu/src/field/cpu_vec_ops.cpp:354-633 and cpu_mont_conversion.cpp:11-27.
 * Mirror of icicle::VecOpsConfig (icicle/include/icicle/vec_ops.h:19-44).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
  void* stream;
  uint8_t is_a_on_device;
  uint8_t is_b_on_device;
  uint8_t is_result_on_device;
  uint8_t is_async;
  int batch_size;
  uint8_t columns_batch;
  uint8_t reserved[3];
} b200_vec_ops_config;

typedef enum {
  B200_VEC_ADD = 0,        /* vector_add          vec_ops_backend.h:85  */
  B200_VEC_SUB = 1,        /* vector_sub          */
  B200_VEC_MUL = 2,        /* vector_mul          */
  B200_VEC_ACCUMULATE = 3, /* vector_accumulate: a[i] += b[i], result pointer ignored */
  B200_SCALAR_ADD_VEC = 4, /* scalar_add_vec: out = a[batch] + b */
  B200_SCALAR_SUB_VEC = 5, /* scalar_sub_vec: out = a[batch] - b */
  B200_SCALAR_MUL_VEC = 6  /* scalar_mul_vec: out = a[batch] * b */
} b200_vec_op_t;

B200_API void b200_vec_ops_default_config(b200_vec_ops_config* cfg);
/* element-wise op over size*batch_size elements (scalar_* ops: `a` holds one scalar per batch, cpu_vec_ops.cpp:325-341) */
B200_API int b200_vec_op(int field, int op, const void* a, const void* b, uint64_t size, const b200_vec_ops_config* cfg, void* out);
/* extension_vector_mixed_mul (vec_ops_backend.h:284-290,346; cpu_vec_ops.cpp): out[i] = a[i] * b[i] with a[] in the quartic extension
 * `ext_field` (B200_FIELD_*_EXT4) and b[] in its base field.  Every other extension vec-op (REGISTER_*_EXT_FIELD_BACKEND,
 * vec_ops_backend.h:297-494) is the ordinary entry point called with the extension's field id. */
B200_API int b200_ext_mixed_mul(int ext_field, const void* a, const void* b, uint64_t size, const b200_vec_ops_config* cfg, void* out);
/* vector_inv / vector_div: out = a^-1, out = a / b element-wise; inverse(0) = 0 like the reference (modular_arithmetic.h:621-623)
 * REGISTER_VECTOR_INV_BACKEND / REGISTER_VECTOR_DIV_BACKEND (vec_ops_backend.h:107,136); cpu_vec_ops.cpp:386-403 */
B200_API int b200_vector_inv(int field, const void* a, uint64_t size, const b200_vec_ops_config* cfg, void* out);
B200_API int b200_vector_div(int field, const void* a, const void* b, uint64_t size, const b200_vec_ops_config* cfg, void* out);
/* vector_sum / vector_product: one output element per batch (VectorReduceOpImpl, vec_ops_backend.h:22-23,156,166; cpu_vec_ops.cpp:428-490) */
B200_API int b200_vector_sum(int field, const void* a, uint64_t size, const b200_vec_ops_config* cfg, void* out);
B200_API int b200_vector_product(int field, const void* a, uint64_t size, const b200_vec_ops_config* cfg, void* out);
/* the three vec-ops the reference's device-agnostic Polynomial backend needs on top of the above:
 * highest_non_zero_idx (vec_ops_backend.h:54-55,236; cpu_vec_ops.cpp:600-633) -- out_idx[batch], -1 for the zero vector
 * poly_eval  (vec_ops_backend.h:64-71,246; cpu_vec_ops.cpp:676-705) -- Horner, coefficient batches x one domain
 * poly_division (vec_ops_backend.h:73-83,256; cpu_vec_ops.cpp:708-777) -- school-book long division, q and r out */
B200_API int b200_highest_non_zero_idx(int field, const void* a, uint64_t size, const b200_vec_ops_config* cfg, int64_t* out_idx);
B200_API int b200_poly_eval(int field, const void* coeffs, uint64_t coeffs_size, const void* domain, uint64_t domain_size,
                            const b200_vec_ops_config* cfg, void* evals);
B200_API int b200_poly_division(int field, const void* numerator, uint64_t numerator_size, const void* denominator, uint64_t denominator_size,
                                const b200_vec_ops_config* cfg, void* q_out, uint64_t q_size, void* r_out, uint64_t r_size);
/* convert_montgomery (vec_ops_backend.h ConvertMontgomery; cpu_vec_ops.cpp) */
B200_API int b200_convert_montgomery(int field, const void* in, uint64_t size, int is_into, const b200_vec_ops_config* cfg, void* out);
/* bit_reverse (cpu_vec_ops.cpp:535-575): out[i] = in[bitrev(i)], size must be a power of two */
B200_API int b200_bit_reverse(int field, const void* in, uint64_t size, const b200_vec_ops_config* cfg, void* out);
/* matrix_transpose (vec_ops_backend.h:204-212; cpu_matrix_ops.cpp): out[c*rows + r] = in[r*cols + c] */
B200_API int b200_matrix_transpose(int field, const void* in, uint32_t rows, uint32_t cols, const b200_vec_ops_config* cfg, void* out);
/* slice (cpu_vec_ops.cpp:577-596): out[i] = in[offset + i*stride] */
B200_API int b200_slice(int field, const void* in, uint64_t offset, uint64_t stride, uint64_t size_in, uint64_t size_out,
               const b200_vec_ops_config* cfg, void* out);
/* curve Montgomery conversion (icicle/include/icicle/curves/montgomery_conversion.h:22-51; cpu_mont_conversion.cpp:11-27) */
B200_API int b200_affine_convert_montgomery(int curve, const void* in, uint64_t n, int is_into, const b200_vec_ops_config* cfg, void* out);
B200_API int b200_projective_convert_montgomery(int curve, const void* in, uint64_t n, int is_into, const b200_vec_ops_config* cfg, void* out);

/* out = sum of n homogeneous projective points (standard form).  New capability: the combine step of a point-sharded
 * multi-GPU MSM after the NCCL all-gather of per-GPU partial results (the reference has no inter-device reduction;
 * ncclReduce cannot add group elements).  Uses the is_a_on_device / is_result_on_device / stream fields of cfg. */
B200_API int b200_ec_sum(int curve, const void* points, int n, const b200_vec_ops_config* cfg, void* out);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU orchestration.  The reference API is one device per call and prescribes "one host thread per
 * device" (docs/docs/start/architecture/multi-device.md:32-36,76; wrappers/rust/icicle-core/src/msm/tests.rs:26-40); these
 * entry points do exactly that inside the backend: HOST-resident inputs/outputs, one host thread per device, the ordinary
 * single-device path on each shard.  Batches shard by batch index (no exchange); one large MSM shards by point range and
 * the per-device partial results are summed with b200_ec_sum (the only exchange: n_devices * |projective| bytes).
 * device_ids == NULL means devices 0 .. n_devices-1; n_devices <= 0 means all visible devices.  The registration shims
 * call them when the caller's ConfigExtension carries the opt-in key "multi_gpu" (number of devices).
 * ---------------------------------------------------------------------------------------------------------------- */
B200_API int b200_msm_multi_gpu(int curve, const void* scalars, const void* bases, int msm_size, const b200_msm_config* cfg, void* results,
                                int n_devices, const int* device_ids);
/* batched NTT: batch rows are partitioned; the twiddle domain of the CURRENT device is replicated on the others */
B200_API int b200_ntt_multi_gpu(int field, const void* input, int size, int dir, const b200_ntt_config* cfg, void* output,
                                int n_devices, const int* device_ids);
/* ONE transform spanning several GPUs (the reference stops at one device, multi-device.md:28-36).
 * N = 2^(a_log+b_log) points viewed as an A x B row-major matrix of the natural-order array; rank r of n_ranks (a power of two
 * dividing A and B) holds the COLUMN SLAB [A][B/n_ranks] on its device.  phase1 (in place): A-point NTTs down the columns +
 * the w_N^(col*k) factors; the slab is then n_ranks contiguous blocks of A/n_ranks rows and block s must be delivered to rank
 * s (all-to-all: NCCL all_to_all_single between processes, peer copies between host threads), each rank receiving its blocks
 * in source-rank order.  phase2: B-point NTTs along the rows + local transpose -> out_slab = column slab [B][A/n_ranks] of the
 * B x A view of the natural-order RESULT.  Both directions (the inverse of a forward result takes a_log/b_log swapped).
 * The domain of the current device must cover N.  b200_ntt_multi_gpu() runs exactly this for a single large host-resident
 * transform (batch 1), with host-side scatter / gather of the slabs. */
B200_API int b200_ntt_dist_phase1(int field, void* slab, int a_log, int b_log, int n_ranks, int rank, int dir, void* stream);
B200_API int b200_ntt_dist_phase2(int field, const void* received, void* out_slab, int a_log, int b_log, int n_ranks, int rank, int dir, void* stream);
/* the contiguous split both deployments use (threads here, one process per GPU in bench.py): part `index` of `parts` */
B200_API void b200_shard_range(uint64_t total, int parts, int index, uint64_t* begin, uint64_t* count);

/* ---- instrumentation (not part of the reference API; used by bench.py and the profiling scripts) ---- */
/* number of kernels of THIS library launched so far in the process (library kernels such as cub's are not counted) */
B200_API long long b200_get_launch_count(void);
/* when on, MSM / NTT calls record CUDA events per stage on the launching stream and synchronise at the end of the call */
B200_API void b200_set_profiling(int on);
/* stage timings of the last profiled call: names_out = "what,stage0,stage1,..."; returns the number of stages */
B200_API int b200_get_last_profile(char* names_out, int names_cap, float* ms_out, int max_stages);

/* developer / test knobs (msm_pair_levels, msm_pipeline_min, msm_pipeline_chunks, msm_no_pipeline, msm_chunk_target,
 * msm_no_wide_loads, msm_staging_mb, msm_sort, ntt_geom, ntt31_off, ntt_columns_strided, ntt_maxr, ntt_tiles, ntt_maxs,
 * ntt31_tma_off, copier_threads): initialised ONCE from the environment (B200_<NAME>) when the library loads -- the hot path never calls
 * getenv() -- and changed afterwards only here; value < 0 = unset (built-in policy).  Returns INVALID_ARGUMENT for an unknown name.
 * One extra name, "l2_fetch_granularity" (32 / 64 / 128), is an explicit OPT-IN to a device-wide CUDA limit
 * (cudaLimitMaxL2FetchGranularity of the current device): 32 cuts the MSM's DRAM traffic by a third at equal run time. */
B200_API int b200_set_tuning(const char* name, int value);
B200_API int b200_get_tuning(const char* name);
/* The library's temporaries come from a PRIVATE stream-ordered pool per device (the process-wide default pool and device
 * limits are never modified); freed scratch is retained there between calls (bounded by B200_SCRATCH_RETAIN_MB at load).
 * b200_trim_scratch() synchronises the current device and returns everything above keep_bytes to the driver. */
B200_API int b200_trim_scratch(size_t keep_bytes);

/* library version / build info string (static storage) */
B200_API const char* b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ICICLE_B200_H */
