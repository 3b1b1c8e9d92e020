// Host <-> device movement for PAGEABLE caller memory (what Rust / Go / C++ callers of the reference API normally pass).
//
// cudaMemcpyAsync from / to pageable memory is staged by the driver through one internal bounce buffer on the calling thread:
// it blocks the caller and runs at a fraction of the PCIe rate (~10 GB/s measured on this box against ~55 GB/s pinned).
// Here a few host copier threads move the data through a ring of pinned slots (memcpy user <-> slot, cudaMemcpyAsync
// slot <-> device on private non-blocking streams), which keeps PCIe busy from ordinary host vectors.  Used by
//   * the chunked MSM pipeline (msm_impl.cuh msm_chunked): copies of chunk i+1 run under the kernels of chunk i;
//   * stage_in / finish_out (common.cuh) for any pageable buffer of >= 32 MiB (NTT / vec-ops host calls).
// Pinned sources take plain cudaMemcpyAsync.  Streams, events and pinned slots are cached per (host thread, device).
#pragma once
#include <cuda_runtime.h>
#include "../../include/icicle_b200.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace b200 {

int tune_copier_threads(); // common.cuh: tune(T_COPIER_THREADS)
// host threads moving pageable memory through the pinned ring: enough to keep PCIe (~55 GB/s) busy with ~6-8 GB/s memcpys each
// CPUs this process may really use: the cgroup v2 quota (containers often show every core of the box but grant a fraction:
// the GPU boxes of this pool expose 128 cores with cpu.max = 16 CPUs), else the affinity mask / core count
inline int usable_cpus()
{
  static const int cached = [] {
    int n = (int)std::thread::hardware_concurrency();
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
      long long quota = 0, period = 0;
      char first[32] = {0};
      if (fscanf(f, "%31s %lld", first, &period) == 2 && strcmp(first, "max") != 0 && period > 0) {
        quota = atoll(first);
        if (quota > 0) n = (int)std::max<long long>(1, std::min<long long>(n, quota / period));
      }
      fclose(f);
    }
    return n > 0 ? n : 8;
  }();
  return cached;
}
// host threads moving pageable memory through the pinned ring: enough to keep PCIe (~55 GB/s) busy with ~8-12 GB/s memcpys each
// (8 is the measured optimum on the 128-core / 16-CPU-quota box, profiles/r2_e2e_copier_threads.txt), but never more than this
// rank's share of the CPUs the container grants (one process per GPU: LOCAL_WORLD_SIZE ranks share them)
inline int copier_thread_count()
{
  const int forced = tune_copier_threads();
  if (forced > 0) return std::min(forced, 32);
  static const int ranks = [] {
    const char* e = getenv("LOCAL_WORLD_SIZE"); // read once per process, not on the hot path
    const int r = e ? atoi(e) : 1;
    return r > 0 ? r : 1;
  }();
  const int share = usable_cpus() / ranks - 2; // leave room for the launching thread and the caller's own work
  return std::max(2, std::min(8, share));
}

struct CopierCtx {
  cudaStream_t st = nullptr;
  void* slot[2] = {nullptr, nullptr};
  cudaEvent_t slot_free[2] = {nullptr, nullptr};
  bool used[2] = {false, false}; // slot_free[k] has been recorded at least once (possibly by an EARLIER call still in flight)
};
constexpr size_t COPIER_SLOT_BYTES = 4u << 20;

struct DeviceRes { // per (host thread, device): streams / events / copier contexts reused call after call
  DeviceRes() = default;
  DeviceRes(const DeviceRes&) = delete;
  DeviceRes& operator=(const DeviceRes&) = delete;
  // thread_local: released when the host thread ends (the multi-GPU orchestrator spawns one thread per device and call); errors
  // from a runtime that is already shutting down are ignored
  ~DeviceRes()
  {
    for (CopierCtx* c : copiers) {
      for (int k = 0; k < 2; k++) {
        if (c->slot[k]) (void)cudaFreeHost(c->slot[k]);
        if (c->slot_free[k]) (void)cudaEventDestroy(c->slot_free[k]);
      }
      if (c->st) (void)cudaStreamDestroy(c->st);
      delete c;
    }
    for (cudaEvent_t e : events) (void)cudaEventDestroy(e);
    if (blocking_ev) (void)cudaEventDestroy(blocking_ev);
    if (copy_stream) (void)cudaStreamDestroy(copy_stream);
    (void)cudaGetLastError();
  }
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t blocking_ev = nullptr; // cudaEventBlockingSync: wait for a stream without burning a CPU
  std::vector<cudaEvent_t> events;
  std::vector<CopierCtx*> copiers;
  cudaEvent_t event(size_t i)
  {
    while (events.size() <= i) {
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      events.push_back(e);
    }
    return events[i];
  }
  CopierCtx* copier(size_t i)
  {
    while (copiers.size() <= i) {
      CopierCtx* c = new CopierCtx;
      bool ok = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) == cudaSuccess;
      for (int k = 0; k < 2 && ok; k++) {
        // BlockingSync: a copier thread waiting for its slot sleeps instead of spinning -- several ranks share a CPU quota
        ok = cudaHostAlloc(&c->slot[k], COPIER_SLOT_BYTES, cudaHostAllocDefault) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->slot_free[k], cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess;
      }
      if (!ok) {
        (void)cudaGetLastError();
        delete c;
        return nullptr;
      }
      copiers.push_back(c);
    }
    return copiers[i];
  }
};
struct DeviceRes;
inline DeviceRes* device_res();
// wait for everything enqueued on `s` WITHOUT spinning (long host-pointer calls: the calling thread would otherwise burn a CPU for
// the whole MSM while several ranks share the container's CPU quota); ~20-50 us wake-up latency, so only used for long calls
inline cudaError_t stream_sync_blocking(cudaStream_t s);

inline DeviceRes* device_res()
{
  static thread_local DeviceRes res[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  DeviceRes* r = &res[dev];
  if (!r->copy_stream && cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  return r;
}

enum HostKind : int { HK_DEVICE = 0, HK_PINNED = 1, HK_PAGEABLE = 2 };
inline HostKind host_kind(const void* p, bool flag_on_device)
{
  if (flag_on_device) return HK_DEVICE;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return HK_PAGEABLE;
  }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return HK_DEVICE;
  return a.type == cudaMemoryTypeHost ? HK_PINNED : HK_PAGEABLE;
}

// One array (scalars or points) of a chunked call: where chunk i's slice lives on the device.
struct ChunkArray {
  const uint8_t* host = nullptr; // source if it has to be copied
  uint8_t* dev = nullptr;        // device base (staging area, or the caller's device buffer)
  size_t elem_bytes = 0;         // bytes per point index
  HostKind kind = HK_DEVICE;
};

// Pageable sources: T host threads move the slices through pinned slots.  Thread t owns pieces t, t+T, ... (ordered by chunk)
// and publishes, per chunk, an event recorded after its last piece of that chunk plus a host-side progress counter, so the
// calling thread only makes the compute stream wait on events that have really been recorded.
struct PageableCopy {
  struct Piece { const uint8_t* src; uint8_t* dst; size_t bytes; uint32_t chunk; };
  std::vector<Piece> pieces;
  std::vector<std::thread> threads;
  std::vector<CopierCtx*> ctx;
  std::vector<cudaEvent_t> chunk_ev; // [thread * nchunks + chunk]
  std::vector<uint8_t> has_ev;       // same indexing: thread t had pieces in chunk c
  std::unique_ptr<std::atomic<int>[]> progress; // per thread: chunks fully issued
  std::atomic<int> failed{0};
  uint32_t nchunks = 0;
  int dev = 0;

  void add(const ChunkArray& a, const uint32_t* coff, const uint32_t* csize, uint32_t nch)
  {
    for (uint32_t c = 0; c < nch; c++) {
      size_t off = (size_t)coff[c] * a.elem_bytes, left = (size_t)csize[c] * a.elem_bytes;
      while (left) {
        const size_t b = std::min(left, COPIER_SLOT_BYTES);
        pieces.push_back({a.host + off, a.dev + off, b, c});
        off += b;
        left -= b;
      }
    }
  }
  void run_thread(int t, int T)
  {
    cudaSetDevice(dev);
    CopierCtx* c = ctx[t];
    int k = 0;
    uint32_t cur_chunk = 0;
    bool any_in_chunk = false;
    auto close_chunks_until = [&](uint32_t upto) { // chunks [cur_chunk, upto) are complete for this thread
      while (cur_chunk < upto) {
        if (any_in_chunk) {
          cudaEventRecord(chunk_ev[(size_t)t * nchunks + cur_chunk], c->st);
          has_ev[(size_t)t * nchunks + cur_chunk] = 1;
        }
        any_in_chunk = false;
        cur_chunk++;
        progress[t].store((int)cur_chunk, std::memory_order_release);
      }
    };
    for (size_t i = (size_t)t; i < pieces.size(); i += (size_t)T) {
      const Piece& pc = pieces[i];
      close_chunks_until(pc.chunk);
      const int sl = k & 1;
      // the slot may still be the source of a copy in flight -- from this call or from the previous one (ring_h2d returns
      // as soon as its copies are issued)
      if (c->used[sl] && cudaEventSynchronize(c->slot_free[sl]) != cudaSuccess) failed.store(1);
      memcpy(c->slot[sl], pc.src, pc.bytes);
      if (cudaMemcpyAsync(pc.dst, c->slot[sl], pc.bytes, cudaMemcpyHostToDevice, c->st) != cudaSuccess) failed.store(1);
      cudaEventRecord(c->slot_free[sl], c->st);
      c->used[sl] = true;
      any_in_chunk = true;
      k++;
    }
    close_chunks_until(nchunks);
  }
  int start(DeviceRes* res, uint32_t nch, cudaEvent_t ready)
  {
    nchunks = nch;
    cudaGetDevice(&dev);
    // the pieces must be ordered by chunk across both arrays so that chunk i completes early: stable sort by chunk
    std::stable_sort(pieces.begin(), pieces.end(), [](const Piece& x, const Piece& y) { return x.chunk < y.chunk; });
    int T = copier_thread_count();
    T = (int)std::min<size_t>((size_t)T, std::max<size_t>(1, pieces.size()));
    for (int t = 0; t < T; t++) {
      CopierCtx* c = res->copier((size_t)t);
      if (!c) return B200_ALLOCATION_FAILED;
      ctx.push_back(c);
      cudaStreamWaitEvent(c->st, ready, 0); // the staging buffers exist and earlier work on the caller's stream is done
    }
    chunk_ev.resize((size_t)T * nch);
    has_ev.assign((size_t)T * nch, 0);
    for (size_t i = 0; i < chunk_ev.size(); i++) chunk_ev[i] = res->event(64 + i);
    progress.reset(new std::atomic<int>[T]);
    for (int t = 0; t < T; t++) progress[t].store(0);
    for (int t = 0; t < T; t++) threads.emplace_back([this, t, T] { run_thread(t, T); });
    return B200_SUCCESS;
  }
  // make stream s wait for chunk c's copies (blocks the HOST until every copier thread has issued them)
  void wait_chunk(uint32_t c, cudaStream_t s)
  {
    const int T = (int)threads.size();
    for (int t = 0; t < T; t++) {
      while (progress[t].load(std::memory_order_acquire) <= (int)c) std::this_thread::sleep_for(std::chrono::microseconds(30)); // no busy spin
      if (has_ev[(size_t)t * nchunks + c]) cudaStreamWaitEvent(s, chunk_ev[(size_t)t * nchunks + c], 0);
    }
  }
  void join()
  {
    for (auto& th : threads) th.join();
    threads.clear();
  }
  ~PageableCopy() { join(); }
};


inline cudaError_t stream_sync_blocking(cudaStream_t s)
{
  DeviceRes* r = device_res();
  if (!r) return cudaStreamSynchronize(s);
  if (!r->blocking_ev && cudaEventCreateWithFlags(&r->blocking_ev, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) {
    (void)cudaGetLastError();
    r->blocking_ev = nullptr;
    return cudaStreamSynchronize(s);
  }
  cudaError_t e = cudaEventRecord(r->blocking_ev, s);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(r->blocking_ev);
}

constexpr size_t RING_MIN_BYTES = 32u << 20;

inline int ring_threads(size_t bytes)
{
  int T = copier_thread_count();
  return (int)std::max<size_t>(1, std::min<size_t>((size_t)T, bytes / COPIER_SLOT_BYTES));
}

// dst_dev[0, bytes) <- src_host (pageable).  Returns after every piece has been ISSUED; stream `s` is made to wait for the
// copies (and the copies wait for what was enqueued on `s` before: stream-ordered allocation of dst_dev included).
inline int ring_h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t s)
{
  DeviceRes* res = device_res();
  if (!res) return B200_UNKNOWN_ERROR;
  cudaEvent_t ready = res->event(0);
  cudaEventRecord(ready, s);
  PageableCopy pc;
  for (size_t off = 0; off < bytes; off += COPIER_SLOT_BYTES)
    pc.pieces.push_back({(const uint8_t*)src_host + off, (uint8_t*)dst_dev + off, std::min(COPIER_SLOT_BYTES, bytes - off), 0});
  int err = pc.start(res, 1, ready);
  if (err) return err;
  pc.wait_chunk(0, s);
  pc.join();
  return pc.failed.load() ? B200_COPY_FAILED : B200_SUCCESS;
}

// dst_host (pageable) <- src_dev[0, bytes) after everything enqueued on `s`.  BLOCKS until the data is in dst_host.
inline int ring_d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t s)
{
  DeviceRes* res = device_res();
  if (!res) return B200_UNKNOWN_ERROR;
  cudaEvent_t ready = res->event(0);
  cudaEventRecord(ready, s);
  const int T = ring_threads(bytes);
  std::vector<CopierCtx*> ctx;
  for (int t = 0; t < T; t++) {
    CopierCtx* c = res->copier((size_t)t);
    if (!c) return B200_ALLOCATION_FAILED;
    cudaStreamWaitEvent(c->st, ready, 0);
    ctx.push_back(c);
  }
  const size_t npieces = (bytes + COPIER_SLOT_BYTES - 1) / COPIER_SLOT_BYTES;
  std::atomic<int> failed{0};
  int dev = 0;
  cudaGetDevice(&dev);
  auto worker = [&](int t) {
    cudaSetDevice(dev);
    CopierCtx* c = ctx[t];
    // two slots: the device -> slot copy of piece k+1 runs while piece k is memcpy'd to the caller's buffer
    size_t mine[2] = {0, 0};
    size_t bsz[2] = {0, 0};
    bool full[2] = {false, false};
    int k = 0;
    auto drain = [&](int sl) {
      if (!full[sl]) return;
      if (cudaEventSynchronize(c->slot_free[sl]) != cudaSuccess) failed.store(1);
      memcpy((uint8_t*)dst_host + mine[sl], c->slot[sl], bsz[sl]);
      full[sl] = false;
    };
    for (int sl = 0; sl < 2; sl++) // an earlier ring_h2d may still be reading the slots
      if (c->used[sl] && cudaEventSynchronize(c->slot_free[sl]) != cudaSuccess) failed.store(1);
    for (size_t i = (size_t)t; i < npieces; i += (size_t)T) {
      const int sl = k & 1;
      drain(sl);
      const size_t off = i * COPIER_SLOT_BYTES, b = std::min(COPIER_SLOT_BYTES, bytes - off);
      if (cudaMemcpyAsync(c->slot[sl], (const uint8_t*)src_dev + off, b, cudaMemcpyDeviceToHost, c->st) != cudaSuccess) failed.store(1);
      cudaEventRecord(c->slot_free[sl], c->st);
      c->used[sl] = true;
      mine[sl] = off; bsz[sl] = b; full[sl] = true;
      drain(sl ^ 1);
      k++;
    }
    drain(0);
    drain(1);
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back(worker, t);
  worker(0);
  for (auto& x : th) x.join();
  return failed.load() ? B200_COPY_FAILED : B200_SUCCESS;
}

} // namespace b200
