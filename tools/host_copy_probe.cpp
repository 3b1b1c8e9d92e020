// Developer probe (GPU box): how fast can host threads memcpy pageable memory into 4 MiB slots, for 1..32 threads, and what
// CPU set / cgroup quota does the container give us?  (profiles/r2_host_copy_probe.txt explains the N>1 e2e numbers.)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sched.h>
#include <thread>
#include <vector>
int main()
{
  cpu_set_t set;
  CPU_ZERO(&set);
  sched_getaffinity(0, sizeof(set), &set);
  printf("hardware_concurrency %u, affinity cpus %d\n", std::thread::hardware_concurrency(), CPU_COUNT(&set));
  if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
    char b[128] = {0};
    if (fgets(b, 127, f)) printf("cgroup cpu.max: %s", b);
    fclose(f);
  }
  const size_t total = (size_t)4 << 30;
  char* src = (char*)malloc(total);
  memset(src, 1, total);
  for (int T : {1, 2, 4, 8, 16, 32}) {
    std::vector<std::thread> th;
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < T; t++) {
      th.emplace_back([=] {
        char* slot = (char*)aligned_alloc(4096, 4 << 20);
        memset(slot, 0, 4 << 20);
        const size_t per = total / T;
        for (size_t off = 0; off < per; off += (4 << 20)) { memcpy(slot, src + (size_t)t * per + off, 4 << 20); asm volatile("" : : "r"(slot) : "memory"); }
        free(slot);
      });
    }
    for (auto& x : th) x.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("threads %2d: %.1f GB/s\n", T, total / s / 1e9);
  }
  return 0;
}
