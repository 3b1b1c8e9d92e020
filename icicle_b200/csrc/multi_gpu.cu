// Multi-GPU orchestration of the MSM / NTT path.
//
// The reference API is single-device per call: the active device is thread-local and "to utilise several GPUs, dedicate one
// host thread to each device" (docs/docs/start/architecture/multi-device.md:32-36,76; the Rust tests drive every device id
// from its own thread, wrappers/rust/icicle-core/src/msm/tests.rs:26-40).  Sharding therefore lives ABOVE the registered
// single-device implementation: these entry points take HOST-resident inputs / outputs, spawn one host thread per device and
// call the ordinary single-GPU path (b200_msm / b200_ntt) on each shard:
//   * batched MSM / batched NTT: the batch index is partitioned -- no data-path exchange at all;
//   * one large MSM: the POINT RANGE is partitioned (each GPU reads only its 1/G of the inputs); the G partial results
//     (one projective point each) are summed on the first device (b200_ec_sum) -- the only exchange is G * |projective| bytes;
//   * batches smaller than the device count: each MSM's point range is split over G / batch devices.
// The registration shims reach this through the opt-in ConfigExtension key "multi_gpu" (= number of devices), so unmodified
// ICICLE callers shard by setting one key; the one-process-per-GPU deployment (torchrun + NCCL, bench.py) uses the same
// shard arithmetic (b200_shard_range) and replaces the host-side combine by an NCCL all-gather + b200_ec_sum.
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

using namespace b200;

extern "C" int b200_internal_ntt_domain_root(int field, void* root_out, int* max_log); // ntt.cu

namespace {

  // contiguous split of `total` units over `parts`: part i gets [begin, begin + count)
  inline void shard_range(uint64_t total, int parts, int i, uint64_t* begin, uint64_t* count)
  {
    const uint64_t base = total / (uint64_t)parts, extra = total % (uint64_t)parts;
    *begin = (uint64_t)i * base + std::min<uint64_t>((uint64_t)i, extra);
    *count = base + ((uint64_t)i < extra ? 1 : 0);
  }

  int resolve_devices(int n_devices, const int* device_ids, std::vector<int>& out)
  {
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess) {
      (void)cudaGetLastError();
      return B200_INVALID_DEVICE;
    }
    if (n_devices <= 0) n_devices = have;
    if (n_devices > have && !device_ids) n_devices = have; // "use up to n devices": clamp to what the box has
    out.resize(n_devices);
    for (int i = 0; i < n_devices; i++) {
      out[i] = device_ids ? device_ids[i] : i;
      if (out[i] < 0 || out[i] >= have) return B200_INVALID_DEVICE;
    }
    return B200_SUCCESS;
  }

  struct DeviceGuard { // restores the calling thread's device
    int saved = 0;
    DeviceGuard() { cudaGetDevice(&saved); }
    ~DeviceGuard() { cudaSetDevice(saved); }
  };

  // one unit of MSM work: `count` consecutive MSMs of the batch starting at `batch0`, restricted to points [p0, p0 + pn)
  struct MsmUnit {
    int worker; // index of the host thread / device slot that runs it (device ids may repeat)
    int batch0, count;
    uint64_t p0, pn;
    int slot; // index into the partial-result array (batch == 1 style units), -1 = writes straight to results
  };

  // reusable barrier for the device threads of one call (C++17: no std::barrier)
  class Barrier
  {
    std::mutex mu;
    std::condition_variable cv;
    int count, waiting = 0, generation = 0;

  public:
    explicit Barrier(int n) : count(n) {}
    void wait()
    {
      std::unique_lock<std::mutex> lk(mu);
      const int gen = generation;
      if (++waiting == count) {
        waiting = 0;
        generation++;
        cv.notify_all();
      } else {
        cv.wait(lk, [&] { return gen != generation; });
      }
    }
  };

  bool coset_is_one(const void* g, size_t bytes)
  {
    if (!g) return true;
    const uint32_t* w = (const uint32_t*)g;
    if (w[0] != 1) return false;
    for (size_t i = 1; i < bytes / 4; i++)
      if (w[i]) return false;
    return true;
  }

  // ONE host-resident transform of 2^n_log points over G devices: scatter column slabs, phase 1, all-to-all by peer copies over
  // NVLink, phase 2, gather (see b200_ntt_dist_phase1/2 in ntt.cu).  Natural order in, natural order out.
  int ntt_distributed(int field, const void* input, int n_log, int dir, void* output, const std::vector<int>& devs, const uint32_t* root)
  {
    const int G = (int)devs.size();
    const int a_log = (n_log + 1) / 2, b_log = n_log - a_log;
    const size_t E = (size_t)b200_field_bytes(field);
    const size_t A = (size_t)1 << a_log, B = (size_t)1 << b_log;
    const size_t slab_elems = (A * B) / (size_t)G, blk = slab_elems / (size_t)G; // block = (A/G) x (B/G) elements
    std::vector<int> rc(G, B200_SUCCESS);
    std::vector<uint8_t*> recv(G, nullptr);
    Barrier bar(G);
    std::vector<std::thread> threads;
    for (int d = 0; d < G; d++) {
      threads.emplace_back([&, d] {
        auto fail = [&](int code) { if (rc[d] == B200_SUCCESS) rc[d] = code; };
        cudaStream_t st = nullptr;
        uint8_t *slab = nullptr, *out = nullptr;
        if (cudaSetDevice(devs[d]) != cudaSuccess) fail(B200_INVALID_DEVICE);
        if (rc[d] == B200_SUCCESS && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) fail(B200_STREAM_CREATION_FAILED);
        if (rc[d] == B200_SUCCESS) {
          for (int o = 0; o < G; o++) // NVLink peer access (already-enabled is fine; same device is skipped)
            if (devs[o] != devs[d]) {
              int can = 0;
              cudaDeviceCanAccessPeer(&can, devs[d], devs[o]);
              if (can && cudaDeviceEnablePeerAccess(devs[o], 0) != cudaSuccess) (void)cudaGetLastError();
            }
          if (b200_ntt_init_domain(field, root, st) != B200_SUCCESS) fail(B200_INVALID_ARGUMENT);
          if (cudaMalloc(&slab, slab_elems * E) != cudaSuccess || cudaMalloc(&recv[d], slab_elems * E) != cudaSuccess || cudaMalloc(&out, slab_elems * E) != cudaSuccess) {
            (void)cudaGetLastError();
            fail(B200_OUT_OF_MEMORY);
          }
        }
        // scatter: column slab d of the A x B view of the host array
        if (rc[d] == B200_SUCCESS &&
            cudaMemcpy2DAsync(slab, (B / G) * E, (const uint8_t*)input + (size_t)d * (B / G) * E, B * E, (B / G) * E, A, cudaMemcpyHostToDevice, st) != cudaSuccess)
          fail(B200_COPY_FAILED);
        if (rc[d] == B200_SUCCESS) {
          const int e = b200_ntt_dist_phase1(field, slab, a_log, b_log, G, d, dir, st);
          if (e) fail(e);
        }
        bar.wait(); // every recv buffer exists
        bool all_ok = true;
        for (int o = 0; o < G; o++) all_ok = all_ok && rc[o] == B200_SUCCESS;
        if (all_ok) {
          for (int o = 0; o < G; o++) { // my block o -> rank o's receive slot d
            if (cudaMemcpyPeerAsync(recv[o] + (size_t)d * blk * E, devs[o], slab + (size_t)o * blk * E, devs[d], blk * E, st) != cudaSuccess) fail(B200_COPY_FAILED);
          }
          if (cudaStreamSynchronize(st) != cudaSuccess) fail(B200_SYNCHRONIZATION_FAILED);
        }
        bar.wait(); // every block has landed
        all_ok = true;
        for (int o = 0; o < G; o++) all_ok = all_ok && rc[o] == B200_SUCCESS;
        if (all_ok) {
          int e = b200_ntt_dist_phase2(field, recv[d], out, a_log, b_log, G, d, dir, st);
          if (e) fail(e);
          // gather: column slab d of the B x A view of the natural-order result
          if (rc[d] == B200_SUCCESS &&
              cudaMemcpy2DAsync((uint8_t*)output + (size_t)d * (A / G) * E, A * E, out, (A / G) * E, (A / G) * E, B, cudaMemcpyDeviceToHost, st) != cudaSuccess)
            fail(B200_COPY_FAILED);
        }
        if (st) {
          if (cudaStreamSynchronize(st) != cudaSuccess) fail(B200_SYNCHRONIZATION_FAILED);
          cudaStreamDestroy(st);
        }
        bar.wait(); // nobody frees a buffer a peer may still read
        cudaFree(slab);
        cudaFree(recv[d]);
        cudaFree(out);
      });
    }
    for (auto& t : threads) t.join();
    for (int d = 0; d < G; d++)
      if (rc[d] != B200_SUCCESS) return rc[d];
    return B200_SUCCESS;
  }

} // namespace

extern "C" {

void b200_shard_range(uint64_t total, int parts, int index, uint64_t* begin, uint64_t* count)
{
  if (parts <= 0 || index < 0 || index >= parts) {
    *begin = 0;
    *count = 0;
    return;
  }
  shard_range(total, parts, index, begin, count);
}

int b200_msm_multi_gpu(int curve, const void* scalars, const void* bases, int msm_size, const b200_msm_config* cfg, void* results,
                       int n_devices, const int* device_ids)
{
  if (!cfg || !scalars || !bases || !results) return B200_INVALID_POINTER;
  if (msm_size <= 0) return B200_INVALID_ARGUMENT;
  // sharding reads slices of host arrays from several devices; device-resident data belongs to ONE device already
  if (cfg->are_scalars_on_device || cfg->are_points_on_device || cfg->are_results_on_device) return B200_INVALID_ARGUMENT;
  std::vector<int> devs;
  int err = resolve_devices(n_devices, device_ids, devs);
  if (err) return err;
  const int G = (int)devs.size();
  const int batch = cfg->batch_size > 0 ? cfg->batch_size : 1;
  const int field = b200_curve_scalar_field(curve);
  if (field < 0) return B200_INVALID_ARGUMENT;
  const size_t sbytes = (size_t)b200_field_bytes(field), abytes = (size_t)b200_curve_affine_bytes(curve), pbytes = (size_t)b200_curve_projective_bytes(curve);
  const int pf = cfg->precompute_factor > 0 ? cfg->precompute_factor : 1;
  const bool shared = cfg->are_points_shared_in_batch || batch == 1;
  const uint64_t n = (uint64_t)msm_size;
  if (G == 1) {
    DeviceGuard guard;
    if (cudaSetDevice(devs[0]) != cudaSuccess) return B200_INVALID_DEVICE;
    return b200_msm(curve, scalars, bases, msm_size, cfg, results);
  }

  b200_msm_config base = *cfg;
  base.stream = nullptr; // the caller's stream belongs to the caller's device; every worker creates its own
  base.is_async = 0;
  // precomputed bases are laid out for the window size of the WHOLE msm (b200_msm_precompute_bases): shards must use it too
  if (pf > 1 && base.c == 0) base.c = b200_msm_choose_c(curve, msm_size, cfg);

  // ---- work units ----------------------------------------------------------------------------------------------------------
  std::vector<MsmUnit> units;
  const int ranges = (batch >= G) ? 1 : std::max(1, G / batch); // point ranges per MSM
  std::vector<uint8_t> partial;
  if (ranges == 1) {
    for (int d = 0; d < G; d++) {
      uint64_t b0, bc;
      shard_range((uint64_t)batch, G, d, &b0, &bc);
      if (bc) units.push_back({d, (int)b0, (int)bc, 0, n, -1});
    }
  } else {
    partial.resize((size_t)batch * ranges * pbytes);
    int d = 0;
    for (int b = 0; b < batch; b++) {
      for (int r = 0; r < ranges; r++) {
        uint64_t p0, pn;
        shard_range(n, ranges, r, &p0, &pn);
        if (pn) units.push_back({d % G, b, 1, p0, pn, b * ranges + r});
        else memset(partial.data() + ((size_t)b * ranges + r) * pbytes, 0, pbytes); // (0,0,0) is skipped by the combine (Z == 0)
        d++;
      }
    }
  }

  // ---- one host thread per device ---------------------------------------------------------------------------------------
  std::vector<int> rc(G, B200_SUCCESS);
  std::vector<std::thread> threads;
  for (int d = 0; d < G; d++) {
    threads.emplace_back([&, d] {
      if (cudaSetDevice(devs[d]) != cudaSuccess) {
        rc[d] = B200_INVALID_DEVICE;
        return;
      }
      cudaStream_t st = nullptr;
      if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
        rc[d] = B200_STREAM_CREATION_FAILED;
        return;
      }
      for (const MsmUnit& u : units) {
        if (u.worker != d || rc[d] != B200_SUCCESS) continue;
        b200_msm_config c = base;
        c.stream = st;
        c.batch_size = u.count;
        const uint8_t* sc = (const uint8_t*)scalars + ((size_t)u.batch0 * n + u.p0) * sbytes;
        const uint8_t* bs = (const uint8_t*)bases + ((shared ? 0 : (size_t)u.batch0 * n * pf) + (size_t)u.p0 * pf) * abytes;
        uint8_t* out = (u.slot < 0) ? (uint8_t*)results + (size_t)u.batch0 * pbytes : partial.data() + (size_t)u.slot * pbytes;
        rc[d] = b200_msm(curve, sc, bs, (int)u.pn, &c, out);
      }
      cudaStreamSynchronize(st);
      cudaStreamDestroy(st);
    });
  }
  for (auto& t : threads) t.join();
  for (int d = 0; d < G; d++)
    if (rc[d] != B200_SUCCESS) return rc[d];

  // ---- combine the point-range partials (the only exchange: ranges * |projective| bytes per MSM) -----------------------------
  if (ranges > 1) {
    DeviceGuard guard;
    if (cudaSetDevice(devs[0]) != cudaSuccess) return B200_INVALID_DEVICE;
    b200_vec_ops_config vc;
    b200_vec_ops_default_config(&vc);
    for (int b = 0; b < batch; b++) {
      err = b200_ec_sum(curve, partial.data() + (size_t)b * ranges * pbytes, ranges, &vc, (uint8_t*)results + (size_t)b * pbytes);
      if (err) return err;
    }
  }
  return B200_SUCCESS;
}

int b200_ntt_multi_gpu(int field, const void* input, int size, int dir, const b200_ntt_config* cfg, void* output, int n_devices, const int* device_ids)
{
  if (!cfg || !input || !output) return B200_INVALID_POINTER;
  if (cfg->are_inputs_on_device || cfg->are_outputs_on_device) return B200_INVALID_ARGUMENT;
  std::vector<int> devs;
  int err = resolve_devices(n_devices, device_ids, devs);
  if (err) return err;
  const int batch = cfg->batch_size > 0 ? cfg->batch_size : 1;
  // ---- ONE large transform: span the devices (4-step with a single all-to-all over NVLink) -------------------------------------
  {
    int n_log = 0;
    while ((1 << n_log) < size) n_log++;
    int Gd = 1;
    while (Gd * 2 <= (int)devs.size()) Gd *= 2; // a power of two of devices
    const size_t ebytes0 = (size_t)b200_field_bytes(field);
    if (batch == 1 && Gd > 1 && size > 0 && (size & (size - 1)) == 0 && n_log >= 16 && (n_log / 2) >= 8 && !cfg->columns_batch && cfg->ordering == B200_NN &&
        coset_is_one(cfg->coset_gen, ebytes0) && (1 << (n_log / 2)) >= Gd) {
      uint32_t root0[32];
      int max_log0 = 0;
      if ((err = b200_internal_ntt_domain_root(field, root0, &max_log0))) return err;
      if (n_log > max_log0) return B200_INVALID_ARGUMENT;
      devs.resize(Gd);
      DeviceGuard guard;
      return ntt_distributed(field, input, n_log, dir, output, devs, root0);
    }
  }
  int G = std::min<int>((int)devs.size(), batch);
  // a column batch shards by column groups, i.e. strided slices of the host matrix: handled on one device (the transposes
  // dominate); a single transform spanning devices is b200_ntt_distributed_* (4-step)
  if (cfg->columns_batch) G = 1;
  if (G <= 1) {
    DeviceGuard guard;
    if (cudaSetDevice(devs[0]) != cudaSuccess) return B200_INVALID_DEVICE;
    return b200_ntt(field, input, size, dir, cfg, output);
  }
  // every device needs the twiddle domain: replicate the caller's (current device) domain
  uint32_t root[32];
  int max_log = 0;
  if ((err = b200_internal_ntt_domain_root(field, root, &max_log))) return err;
  const size_t ebytes = (size_t)b200_field_bytes(field);
  std::vector<int> rc(G, B200_SUCCESS);
  std::vector<std::thread> threads;
  for (int d = 0; d < G; d++) {
    threads.emplace_back([&, d] {
      if (cudaSetDevice(devs[d]) != cudaSuccess) {
        rc[d] = B200_INVALID_DEVICE;
        return;
      }
      cudaStream_t st = nullptr;
      if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
        rc[d] = B200_STREAM_CREATION_FAILED;
        return;
      }
      if ((rc[d] = b200_ntt_init_domain(field, root, st)) == B200_SUCCESS) { // idempotent per (field, device)
        uint64_t b0, bc;
        shard_range((uint64_t)batch, G, d, &b0, &bc);
        b200_ntt_config c = *cfg;
        c.stream = st;
        c.is_async = 0;
        c.batch_size = (int)bc;
        const size_t off = (size_t)b0 * (size_t)size * ebytes;
        if (bc) rc[d] = b200_ntt(field, (const uint8_t*)input + off, size, dir, &c, (uint8_t*)output + off);
      }
      cudaStreamSynchronize(st);
      cudaStreamDestroy(st);
    });
  }
  for (auto& t : threads) t.join();
  for (int d = 0; d < G; d++)
    if (rc[d] != B200_SUCCESS) return rc[d];
  return B200_SUCCESS;
}

} // extern "C"
