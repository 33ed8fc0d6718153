// Host-side int64 -> int32 saturating narrow of an action buffer: how fast can the box's cores shrink the PCIe payload?
// g++ -O3 -pthread -o tools/probes/probe_narrow.bin tools/probes/probe_narrow.cpp
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void narrow(const int64_t* in, int32_t* out, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    const uint64_t v = (uint64_t)in[i];
    out[i] = v < 0x80000000ull ? (int32_t)v : -1;
  }
}

int main() {
  const size_t n = (size_t)1 << 29;                       // 4 GiB in, 2 GiB out
  int64_t* in = (int64_t*)malloc(n * 8);
  int32_t* out = (int32_t*)malloc(n * 4);
  for (size_t i = 0; i < n; ++i) in[i] = (int64_t)(i * 2654435761u % 600000);
  memset(out, 0, n * 4);
  printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
  for (int nt : {1, 4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t)
        th.emplace_back([=] { const size_t a = n * t / nt, b = n * (t + 1) / nt; narrow(in + a, out + a, b - a); });
      for (auto& x : th) x.join();
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (rep) printf("threads %2d: %.1f ms, %.1f GB/s of int64 input (%.1f GB/s in+out)\n", nt, s * 1e3, n * 8 / s / 1e9, n * 12 / s / 1e9);
    }
  }
  return 0;
}
