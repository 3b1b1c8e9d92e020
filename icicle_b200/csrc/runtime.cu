// Device runtime entry points of the C ABI: what our DeviceAPI subclass (icicle_b200/shim/device_api_shim.cpp) forwards
// to.  Semantics follow the reference's DeviceAPI contract (icicle/include/icicle/device_api.h:44-182) with the CUDA
// mapping the reference itself uses for its one open CUDA device (icicle/backend/cuda_pqc/src/cuda_pqc_device_api.cu).
#include "common.cuh"
#include <cstring>

using namespace b200;

#include <atomic>
#include <mutex>
#include <string>
namespace {
  std::atomic<long long> g_launches{0};
  std::atomic<int> g_profiling{0};
  std::mutex g_prof_mu;
  std::string g_prof_what;
  int g_prof_n = 0;
  char g_prof_names[StageTimer::MAX_STAGES][32];
  float g_prof_ms[StageTimer::MAX_STAGES];
} // namespace

namespace b200 {
  void StageTimer::begin(cudaStream_t stream)
  {
    on = g_profiling.load() != 0;
    s = stream;
    n = 0;
    if (!on) return;
    cudaEventCreate(&ev[0]);
    cudaEventRecord(ev[0], s);
  }
  void StageTimer::mark(const char* name)
  {
    if (!on || n >= MAX_STAGES) return;
    names[n] = name;
    cudaEventCreate(&ev[n + 1]);
    cudaEventRecord(ev[n + 1], s);
    n++;
  }
  void StageTimer::finish(const char* what)
  {
    if (!on) return;
    cudaEventSynchronize(ev[n]);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_what = what;
    g_prof_n = n;
    for (int i = 0; i < n; i++) {
      cudaEventElapsedTime(&g_prof_ms[i], ev[i], ev[i + 1]);
      strncpy(g_prof_names[i], names[i], 31);
      g_prof_names[i][31] = 0;
    }
    for (int i = 0; i <= n; i++) cudaEventDestroy(ev[i]);
  }
} // namespace b200

extern "C" {

void b200_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long b200_get_launch_count(void) { return g_launches.load(); }
void b200_set_profiling(int on) { g_profiling.store(on); }
int b200_get_last_profile(char* names_out, int names_cap, float* ms_out, int max_stages)
{
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::string joined = g_prof_what;
  int k = g_prof_n < max_stages ? g_prof_n : max_stages;
  for (int i = 0; i < k; i++) {
    joined += std::string(",") + g_prof_names[i];
    ms_out[i] = g_prof_ms[i];
  }
  if (names_out && names_cap > 0) {
    strncpy(names_out, joined.c_str(), names_cap - 1);
    names_out[names_cap - 1] = 0;
  }
  return k;
}

const char* b200_version(void) { return "icicle_b200 0.1 (sm_100a; MSM+NTT+vec-ops; C ABI v1)"; }

int b200_get_device_count(int* count)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    *count = 0;
    return B200_INVALID_DEVICE;
  }
  *count = n;
  return B200_SUCCESS;
}

int b200_set_device(int device_id)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device_id < 0 || device_id >= n) {
    (void)cudaGetLastError();
    return B200_INVALID_DEVICE;
  }
  B200_CUDA_TRY(cudaSetDevice(device_id), B200_INVALID_DEVICE);
  // keep freed stream-ordered scratch in the pool: MSM/NTT temporaries are re-used call after call
  static thread_local unsigned long long configured_mask = 0;
  if (device_id < 64 && !(configured_mask & (1ull << device_id))) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
      uint64_t thresh = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    configured_mask |= (1ull << device_id);
  }
  return B200_SUCCESS;
}

int b200_malloc(void** ptr, size_t bytes)
{
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    *ptr = nullptr;
    return map_alloc_error(e);
  }
  return B200_SUCCESS;
}
int b200_malloc_async(void** ptr, size_t bytes, void* stream)
{
  cudaError_t e = cudaMallocAsync(ptr, bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    *ptr = nullptr;
    return map_alloc_error(e);
  }
  return B200_SUCCESS;
}
int b200_free(void* ptr)
{
  B200_CUDA_TRY(cudaFree(ptr), B200_DEALLOCATION_FAILED);
  return B200_SUCCESS;
}
int b200_free_async(void* ptr, void* stream)
{
  B200_CUDA_TRY(cudaFreeAsync(ptr, (cudaStream_t)stream), B200_DEALLOCATION_FAILED);
  return B200_SUCCESS;
}
int b200_get_available_memory(size_t* total, size_t* free_bytes)
{
  B200_CUDA_TRY(cudaMemGetInfo(free_bytes, total), B200_UNKNOWN_ERROR);
  return B200_SUCCESS;
}
int b200_memset(void* ptr, int value, size_t bytes)
{
  B200_CUDA_TRY(cudaMemset(ptr, value, bytes), B200_UNKNOWN_ERROR);
  return B200_SUCCESS;
}
int b200_memset_async(void* ptr, int value, size_t bytes, void* stream)
{
  B200_CUDA_TRY(cudaMemsetAsync(ptr, value, bytes, (cudaStream_t)stream), B200_UNKNOWN_ERROR);
  return B200_SUCCESS;
}
static int copy_impl(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, void* stream, int is_async)
{
  if (is_async) {
    B200_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, kind, (cudaStream_t)stream), B200_COPY_FAILED);
  } else {
    B200_CUDA_TRY(cudaMemcpy(dst, src, bytes, kind), B200_COPY_FAILED);
  }
  return B200_SUCCESS;
}
int b200_copy_to_device(void* dst, const void* src, size_t bytes, void* stream, int is_async)
{
  return copy_impl(dst, src, bytes, cudaMemcpyHostToDevice, stream, is_async);
}
int b200_copy_to_host(void* dst, const void* src, size_t bytes, void* stream, int is_async)
{
  return copy_impl(dst, src, bytes, cudaMemcpyDeviceToHost, stream, is_async);
}
int b200_copy_device_to_device(void* dst, const void* src, size_t bytes, void* stream, int is_async)
{
  return copy_impl(dst, src, bytes, cudaMemcpyDeviceToDevice, stream, is_async);
}
int b200_synchronize(void* stream)
{
  if (stream) {
    B200_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream), B200_SYNCHRONIZATION_FAILED);
  } else {
    B200_CUDA_TRY(cudaDeviceSynchronize(), B200_SYNCHRONIZATION_FAILED);
  }
  return B200_SUCCESS;
}
int b200_create_stream(void** stream)
{
  cudaStream_t s;
  B200_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), B200_STREAM_CREATION_FAILED);
  *stream = (void*)s;
  return B200_SUCCESS;
}
int b200_destroy_stream(void* stream)
{
  B200_CUDA_TRY(cudaStreamDestroy((cudaStream_t)stream), B200_STREAM_DESTRUCTION_FAILED);
  return B200_SUCCESS;
}
int b200_host_alloc_pinned(void** ptr, size_t bytes)
{
  cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    *ptr = nullptr;
    return B200_ALLOCATION_FAILED;
  }
  return B200_SUCCESS;
}
int b200_host_free_pinned(void* ptr)
{
  B200_CUDA_TRY(cudaFreeHost(ptr), B200_DEALLOCATION_FAILED);
  return B200_SUCCESS;
}

int b200_field_bytes(int field) { return 4 * field_limbs(field); }

int b200_curve_scalar_field(int curve)
{
  switch (curve) {
  case B200_CURVE_BN254_G1: case B200_CURVE_BN254_G2: return B200_FIELD_BN254_FR;
  case B200_CURVE_BLS12_381_G1: case B200_CURVE_BLS12_381_G2: return B200_FIELD_BLS12_381_FR;
  case B200_CURVE_BLS12_377_G1: case B200_CURVE_BLS12_377_G2: return B200_FIELD_BLS12_377_FR;
  case B200_CURVE_BW6_761_G1: case B200_CURVE_BW6_761_G2: return B200_FIELD_BLS12_377_FQ;
  case B200_CURVE_GRUMPKIN: return B200_FIELD_BN254_FQ;
  default: return -1;
  }
}
static int curve_coord_bytes(int curve)
{
  switch (curve) {
  case B200_CURVE_BN254_G1: case B200_CURVE_GRUMPKIN: return 32;
  case B200_CURVE_BN254_G2: return 64;
  case B200_CURVE_BLS12_381_G1: case B200_CURVE_BLS12_377_G1: return 48;
  case B200_CURVE_BLS12_381_G2: case B200_CURVE_BLS12_377_G2: return 96;
  case B200_CURVE_BW6_761_G1: case B200_CURVE_BW6_761_G2: return 96;
  default: return 0;
  }
}
int b200_curve_affine_bytes(int curve) { return 2 * curve_coord_bytes(curve); }
int b200_curve_projective_bytes(int curve) { return 3 * curve_coord_bytes(curve); }

} // extern "C"
