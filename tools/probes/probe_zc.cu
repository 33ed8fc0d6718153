// Zero-copy read rate from pinned host memory: 16-byte loads vs cp.async.bulk tiles (does the TMA path get
// larger PCIe read completions than SM loads?). Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -o build/probe_zc tools/probes/probe_zc.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_ldg(const uint4* __restrict__ p, size_t n, unsigned long long* out) {
  unsigned long long acc = 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 7 * stride < n; i += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x1234567ull) *out = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int TILE, int STAGES>
__global__ void k_bulk(const unsigned char* __restrict__ p, size_t bytes, unsigned long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar[STAGES];
  const size_t ntiles = bytes / TILE;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned long long acc = 0;
  size_t t = blockIdx.x;
  uint32_t phase = 0;
  int issued = 0;
  auto issue = [&](size_t tile, int s) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(TILE) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + (size_t)s * TILE)),
                 "l"(p + tile * TILE), "r"(TILE), "r"(smem_u32(&bar[s])) : "memory");
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES && t + (size_t)s * gridDim.x < ntiles; ++s) issue(t + (size_t)s * gridDim.x, s);
  int s = 0;
  for (; t < ntiles; t += gridDim.x) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar[s])), "r"(phase) : "memory");
    const uint4* v = reinterpret_cast<const uint4*>(sm + (size_t)s * TILE);
    for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x) { const uint4 x = v[i]; acc += x.x ^ x.y ^ x.z ^ x.w; }
    __syncthreads();
    const size_t nxt = t + (size_t)STAGES * gridDim.x;
    if (threadIdx.x == 0 && nxt < ntiles) issue(nxt, s);
    if (++s == STAGES) { s = 0; phase ^= 1u; }
  }
  (void)issued;
  if (acc == 0x1234567ull) *out = acc;
}

int main() {
  const size_t bytes = (size_t)2 << 30;
  unsigned char* h;
  CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
  for (size_t i = 0; i < bytes; i += 4096) h[i] = (unsigned char)i;
  unsigned char* d;
  CK(cudaMalloc(&d, bytes));
  unsigned long long* out;
  CK(cudaMalloc(&out, 8));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  auto report = [&](const char* name) {
    CK(cudaEventSynchronize(b));
    cudaEventElapsedTime(&ms, a, b);
    printf("%-40s %8.2f GB/s\n", name, bytes / ms / 1e6);
  };
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(a); CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice)); cudaEventRecord(b); report("cudaMemcpyAsync H2D");
    for (int blocks : {148, 296, 592, 1184}) {
      cudaEventRecord(a); k_ldg<<<blocks, 512>>>(reinterpret_cast<const uint4*>(h), bytes / 16, out); cudaEventRecord(b);
      char nm[64]; snprintf(nm, 64, "LDG.128 x8, %d CTAs x 512", blocks); report(nm);
    }
    {
      CK(cudaFuncSetAttribute(k_bulk<16384, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      CK(cudaFuncSetAttribute(k_bulk<4096, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
      CK(cudaFuncSetAttribute(k_bulk<32768, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
      for (int blocks : {148, 296}) {
        char nm[64];
        cudaEventRecord(a); k_bulk<4096, 8><<<blocks, 128, 32768>>>(h, bytes, out); cudaEventRecord(b); snprintf(nm, 64, "bulk 4 KB x8 stages, %d CTAs", blocks); report(nm);
        cudaEventRecord(a); k_bulk<16384, 4><<<blocks, 128, 65536>>>(h, bytes, out); cudaEventRecord(b); snprintf(nm, 64, "bulk 16 KB x4 stages, %d CTAs", blocks); report(nm);
      }
      cudaEventRecord(a); k_bulk<32768, 4><<<148, 128, 131072>>>(h, bytes, out); cudaEventRecord(b); report("bulk 32 KB x4 stages, 148 CTAs");
    }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
