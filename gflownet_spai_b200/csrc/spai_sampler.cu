// C-ABI host side of the whole-trajectory sampler kernels (k4g_gumbel.cuh, k4p_steps.cuh);
// declarations in include/spai_b200.h.
#include <algorithm>

#include "k4g_gumbel.cuh"
#include "k4p_steps.cuh"
#include "spai_internal.cuh"

using namespace spai;

namespace {

struct Guard {
  int prev = -1;
  explicit Guard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
  ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

// stream-ordered scratch from the device's default pool (kept across calls: no allocation cost after the first)
struct Pool {
  cudaStream_t st;
  void* p[8];
  int cnt = 0;
  explicit Pool(cudaStream_t s, int device) : st(s) {
    static bool tuned[64] = {};
    if (device >= 0 && device < 64 && !tuned[device]) {
      cudaMemPool_t mp;
      if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      tuned[device] = true;
    }
  }
  template <typename T> int get(T** out, int64_t count) {
    void* q = nullptr;
    const cudaError_t e = cudaMallocAsync(&q, (size_t)std::max<int64_t>(count, 1) * sizeof(T), st);
    if (e != cudaSuccess || cnt >= 8) {
      set_error("sampler: cudaMallocAsync failed: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return SPAI_ERR_NOMEM;
    }
    p[cnt++] = q;
    *out = reinterpret_cast<T*>(q);
    return SPAI_OK;
  }
  ~Pool() { for (int i = 0; i < cnt; ++i) cudaFreeAsync(p[i], st); }
};

int sm_count(int device) {
  int sm = 148;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
  return sm;
}

}  // namespace

extern "C" {

int spai_sample_taken_dev(int device, const float* logits, int64_t A, int64_t B, uint64_t seed, int64_t sample0,
                          uint32_t* taken, int64_t words_ld, int32_t* length, float* keys_out, int64_t keys_ld,
                          void* stream) {
  if (!logits || !taken || !length || A <= 0 || A >= ((int64_t)1 << 31) || B < 0 || B >= ((int64_t)1 << 31) ||
      sample0 < 0 || words_ld < (A + 31) / 32 || (keys_out && keys_ld < A)) {
    set_error("spai_sample_taken_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  if (B == 0) return SPAI_OK;
  Guard g(device);
  k4g_count_kernel<<<(unsigned)B, K4G_COUNT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, A, seed, sample0, taken, words_ld, length, keys_out, keys_ld);
  SPAI_CUDA(cudaGetLastError());
  return SPAI_OK;
}

int spai_sample_order_dev(int device, const float* logits, int64_t A, int64_t B, uint64_t seed, int64_t sample0,
                          const int32_t* length, void* actions, int id_bytes, int64_t ld, void* stream) {
  if (!logits || !length || !actions || A <= 0 || A >= ((int64_t)1 << 31) || B < 0 || sample0 < 0 || ld < 1 ||
      (id_bytes != 4 && id_bytes != 8)) {
    set_error("spai_sample_order_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  if (B == 0) return SPAI_OK;
  Guard g(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int nb = 64;
  while (nb < K4G_MAX_BUCKETS && (int64_t)nb * 12 < ld) nb <<= 1;
  const int smem = k4g_smem_bytes(nb);
  static bool attr_set = false;
  if (!attr_set) {
    SPAI_CUDA(cudaFuncSetAttribute(k4g_order_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   k4g_smem_bytes(K4G_MAX_BUCKETS)));
    SPAI_CUDA(cudaFuncSetAttribute(k4g_order_kernel<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   k4g_smem_bytes(K4G_MAX_BUCKETS)));
    attr_set = true;
  }
  const int grid = (int)std::min<int64_t>(B, sm_count(device));
  const int64_t scratch_ld = std::min<int64_t>(ld, A);
  Pool pool(st, device);
  float *ntab = nullptr, *mx = nullptr;
  int* flags = nullptr;
  unsigned long long* scratch = nullptr;
  SPAI_TRY(pool.get(&ntab, K4G_TABLE));
  SPAI_TRY(pool.get(&mx, 1));
  SPAI_TRY(pool.get(&flags, 2));
  SPAI_TRY(pool.get(&scratch, (int64_t)grid * scratch_ld));
  SPAI_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int), st));
  k4g_max_kernel<<<1, 1024, 0, st>>>(logits, A, mx);
  k4g_ntable_kernel<<<K4G_TABLE, 256, 0, st>>>(logits, A, mx, ntab);
  if (id_bytes == 4)
    k4g_order_kernel<int32_t><<<grid, K4G_THREADS, smem, st>>>(logits, A, seed, sample0, B, length, ntab, mx, nb, scratch,
                                                              scratch_ld, reinterpret_cast<int32_t*>(actions), ld,
                                                              flags, flags + 1);
  else
    k4g_order_kernel<long long><<<grid, K4G_THREADS, smem, st>>>(logits, A, seed, sample0, B, length, ntab, mx, nb, scratch,
                                                                scratch_ld, reinterpret_cast<long long*>(actions), ld,
                                                                flags, flags + 1);
  SPAI_CUDA(cudaGetLastError());
  int herr[2] = {0, 0};
  SPAI_CUDA(cudaMemcpyAsync(herr, flags, sizeof(herr), cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));            // the status is host-visible: the only sync of the sampler
  if (herr[1] == 1) {
    set_error("spai_sample_order_dev: a trajectory is longer than ld (%lld)", (long long)ld);
    return SPAI_ERR_INVALID;
  }
  if (herr[1]) {
    set_error("spai_sample_order_dev: internal error %d (histogram and length disagree: `length` does not come "
              "from spai_sample_taken_dev with the same logits / seed / sample0)", herr[1]);
    return SPAI_ERR_INVALID;
  }
  return SPAI_OK;
}

int spai_sample_steps_dev(int device, const float* logits, int64_t A, int64_t B, uint32_t* taken, int64_t words_ld,
                          uint8_t* done, const float* uniforms, uint64_t seed, int64_t sample0, int64_t step0,
                          int64_t nsteps, void* actions, int id_bytes, float* probs, int64_t ld, int32_t* steps_taken,
                          void* stream) {
  if (!logits || !taken || !done || !actions || A <= 0 || A >= ((int64_t)1 << 31) || B < 0 || nsteps < 0 || sample0 < 0 ||
      step0 < 0 || words_ld < (A + 31) / 32 || ld < step0 + nsteps || (id_bytes != 4 && id_bytes != 8)) {
    set_error("spai_sample_steps_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  if (B == 0 || nsteps == 0) return SPAI_OK;
  Guard g(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // ~sqrt(A) ids per block (a multiple of 128: 16-byte loads per lane), at most K4P_MAX_NBLK blocks (a multiple of 32, zero-padded)
  int64_t blk = 128;
  while (blk * blk < A) blk += 128;
  blk = std::max<int64_t>(blk, round_up(ceil_div(A, (int64_t)K4P_MAX_NBLK), 128));
  const int nblk = (int)round_up(ceil_div(A, blk), 32);
  const int smem = K4P_WARPS * nblk * (int)sizeof(float);
  const unsigned grid = (unsigned)ceil_div(B, (int64_t)K4P_WARPS);
  if (id_bytes == 4)
    k4p_steps_kernel<int32_t><<<grid, K4P_WARPS * 32, smem, st>>>(logits, A, taken, words_ld, done, uniforms, seed, sample0,
                                                                 step0, B, nsteps, (int)blk, nblk,
                                                                 reinterpret_cast<int32_t*>(actions), probs, ld, steps_taken);
  else
    k4p_steps_kernel<long long><<<grid, K4P_WARPS * 32, smem, st>>>(logits, A, taken, words_ld, done, uniforms, seed, sample0,
                                                                   step0, B, nsteps, (int)blk, nblk,
                                                                   reinterpret_cast<long long*>(actions), probs, ld,
                                                                   steps_taken);
  SPAI_CUDA(cudaGetLastError());
  return SPAI_OK;
}

}  // extern "C"
