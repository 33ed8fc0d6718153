// K4g — whole trajectories of the masked-categorical environment step in one pass
// (exponential race / Gumbel-top-k), no B x A key tensor and no library sort.
//
// The reference draws one id per step from softmax(logits with the taken ids at -inf) until the
// terminal id comes up (policy.py:64-73, gflownet/gflownet.py:135-179, log.py:67-87). The graph and
// the weights do not change inside sample_states, so the pre-mask logits are ONE vector l[0..A) per
// epoch, and that process is the exponential race: id i "arrives" at time t_i = E_i / exp(l_i),
// E_i ~ Exp(1) independent; ids come out in order of arrival, and the trajectory stops at the
// terminal's arrival. In logs, with a_i = log E_i - l_i:
//     lq_i = a_i - a_terminal        (taken  <=>  lq_i < 0;  drawn in ascending (lq, id) order)
// E_i comes from Philox4x32-10 keyed by the call's seed, counter = (id / 4, global sample index), so
// a key is a pure function of (seed, sample, id): nothing is stored, every pass recomputes it, and a
// shard of a batch equals the same rows of the full batch.
//
// k4g_count_kernel   one CTA per sample: taken-bitmask (warp-assembled words) + length. This is all
//                    the reward kernels need (spai_reward_from_taken_dev).
// k4g_order_kernel   persistent CTAs, one sample at a time: bucket the taken ids by a monotone map of
//                    lq into NB <= 32768 buckets whose EXPECTED occupancy is equal (histogram in shared
//                    memory -> scan -> scatter (key, id) to an L2-resident scratch row), then finish
//                    each bucket by rank counting, one thread pair per bucket on shared-memory tiles.
//                    The map is the expected-arrivals function N(t) = sum_i (1 - exp(-w_i t)) of THIS policy
//                    (k4g_ntable_kernel) tabulated per sample on 4096 cells uniform in lq, piecewise linear
//                    and monotone by construction (max-scanned table, fmaf of a non-negative difference,
//                    clamped to the cell's ends), so bucket order == key order exactly and occupancy stays
//                    ~uniform for any policy.
#pragma once

#include "spai_internal.cuh"

namespace spai {

constexpr int K4G_TABLE = 832;                 // log2 t in [-40, 168), 4 points per octave
constexpr float K4G_LOG2T_MIN = -40.f;
constexpr float K4G_PER_OCTAVE = 4.f;
constexpr int K4G_CELLS = 4096;                // cells of the per-sample bucket map, uniform in lq over [-K4G_LQ_SPAN, 0)
constexpr float K4G_LQ_SPAN = 24.f;            // exp(-24) = 4e-11 of the terminal's arrival time: everything earlier -> bucket 0
constexpr int K4G_COUNT_THREADS = 512;
constexpr int K4G_THREADS = 1024;
constexpr int K4G_TILE_BUCKETS = K4G_THREADS / 2;   // one thread pair per bucket
constexpr int K4G_TILE_CAP = 8192;             // (key, id) pairs of a tile in shared memory (64 KB)
constexpr int K4G_BIG = 128;                   // buckets above this are ranked by the whole CTA
constexpr int K4G_BIG_LIST = 128;
constexpr int K4G_MAX_BUCKETS = 32768;

__host__ __device__ constexpr int k4g_smem_bytes(int nb) {
  return nb * 4 + (K4G_CELLS + 4) * 4 + 256 + K4G_BIG_LIST * 4 + K4G_TILE_CAP * 8;
}

// Philox4x32-10 (Salmon et al., SC'11): round keys key + r * (0x9E3779B9, 0xBB67AE85)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// log of an Exp(1) variate from 32 random bits: u = (x + 1/2) / 2^32, E = -log u. Relative error of E
// <= 6e-6 everywhere, also for u -> 1 where E -> 0 (the earliest arrivals): there 1 - u = d is what fp32
// holds exactly, E = -log(1 - d) = d + d^2/2 + ... (degree 6 below d = 1/16, truncation 2^-27 relative),
// and one MUFU.LG2 serves both other ranges (its 2^-21.4 absolute error is relative to E >= 0.0645).
// Branch-free: the warp never runs two log paths back to back.
__device__ __forceinline__ float k4g_log_exp1(uint32_t x) {
  const bool top = (x & 0x80000000u) != 0u;
  const float d = ((float)(~x) + 0.5f) * 2.3283064365386963e-10f;      // 1 - u
  const float u = ((float)x + 0.5f) * 2.3283064365386963e-10f;
  float e = -__logf(top ? 1.0f - d : u);
  const float poly = d * fmaf(d, fmaf(d, fmaf(d, fmaf(d, fmaf(d, 1.f / 6.f, 0.2f), 0.25f), 1.f / 3.f), 0.5f), 1.0f);
  if (top && d < 0.0625f) e = poly;
  return __logf(e);
}

__device__ __forceinline__ uint4 k4g_bits(int64_t group, int64_t sample, uint64_t seed) {
  return philox4x32_10(make_uint4((uint32_t)group, (uint32_t)((uint64_t)group >> 32), (uint32_t)sample,
                                  (uint32_t)((uint64_t)sample >> 32)),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

__device__ __forceinline__ float k4g_arrival(const float* __restrict__ logits, int64_t id, int64_t sample, uint64_t seed) {
  const uint4 r = k4g_bits(id >> 2, sample, seed);
  const uint32_t x = (id & 3) == 0 ? r.x : (id & 3) == 1 ? r.y : (id & 3) == 2 ? r.z : r.w;
  return k4g_log_exp1(x) - __ldg(logits + id);
}

// the four keys lq of group g (ids 4g .. 4g+3); ids >= A - 1 (terminal, out of range) get +inf
__device__ __forceinline__ void k4g_group_keys(const float* __restrict__ logits, int64_t A, int64_t g, int64_t sample,
                                               uint64_t seed, float a_term, bool vec_ok, float (&lq)[4]) {
  const uint4 r = k4g_bits(g, sample, seed);
  const int64_t id0 = g << 2;
  float l[4];
  if (vec_ok && id0 + 3 < A) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(logits + id0));
    l[0] = v.x; l[1] = v.y; l[2] = v.z; l[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) l[j] = (id0 + j < A) ? __ldg(logits + id0 + j) : 0.f;
  }
  const uint32_t x[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    lq[j] = (id0 + j < A - 1) ? (k4g_log_exp1(x[j]) - l[j]) - a_term : INFINITY;
}

// ---------------------------------------------------------------------------------------------
// taken-bitmask + length (+ optional export of every key, tests only)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K4G_COUNT_THREADS)
k4g_count_kernel(const float* __restrict__ logits, int64_t A, uint64_t seed, int64_t sample0,
                 uint32_t* __restrict__ taken, int64_t words_ld, int32_t* __restrict__ length,
                 float* __restrict__ lq_out, int64_t lq_ld) {
  __shared__ int s_cnt[K4G_COUNT_THREADS / 32];
  const int64_t b = blockIdx.x, sample = sample0 + b;
  const int lane = threadIdx.x & 31;
  const float a_term = k4g_arrival(logits, A - 1, sample, seed);
  const bool vec_ok = (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
  const int64_t groups = (A + 3) >> 2, words = (A + 31) >> 5;
  int cnt = 0;
  for (int64_t g0 = 0; g0 < groups; g0 += K4G_COUNT_THREADS) {      // warp-uniform trip count
    const int64_t g = g0 + threadIdx.x;
    uint32_t nib = 0;
    if (g < groups) {
      float lq[4];
      k4g_group_keys(logits, A, g, sample, seed, a_term, vec_ok, lq);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t id = (g << 2) + j;
        if (lq[j] < 0.f || id == A - 1) nib |= 1u << j;              // the terminal's own bit is set (it is drawn last)
        if (lq_out && id < A) lq_out[b * lq_ld + id] = (id == A - 1) ? 0.f : lq[j];
      }
    }
    cnt += __popc(nib);
    uint32_t v = nib << (4 * (lane & 7));
    v |= __shfl_xor_sync(0xffffffffu, v, 1);
    v |= __shfl_xor_sync(0xffffffffu, v, 2);
    v |= __shfl_xor_sync(0xffffffffu, v, 4);
    if ((lane & 7) == 0 && (g >> 3) < words) taken[b * words_ld + (g >> 3)] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < K4G_COUNT_THREADS / 32; ++i) t += s_cnt[i];
    length[b] = t;                                                   // ids drawn, terminal included
  }
}

// ---------------------------------------------------------------------------------------------
// policy-wide tables: max logit, expected arrivals N(t) on a log2 grid
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k4g_max_kernel(const float* __restrict__ logits, int64_t A, float* __restrict__ out) {
  __shared__ float s[32];
  float m = -INFINITY;
  for (int64_t i = threadIdx.x; i < A; i += 1024) m = fmaxf(m, logits[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = s[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) out[0] = m;
  }
}

// ntab[k] = sum_{i < A-1} (1 - exp(-w_i t_k)),  w_i = exp(l_i - max l),  log2 t_k = -40 + k/4
__global__ void __launch_bounds__(256) k4g_ntable_kernel(const float* __restrict__ logits, int64_t A,
                                                         const float* __restrict__ mxp, float* __restrict__ ntab) {
  __shared__ double s[8];
  const float lt = K4G_LOG2T_MIN + (float)blockIdx.x / K4G_PER_OCTAVE;
  const float mx = mxp[0];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < A - 1; i += 256) {
    const float x = exp2f((logits[i] - mx) * 1.4426950408889634f + lt);     // w_i * t_k (inf is fine)
    acc += (double)(-expm1f(-x));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += s[i];
    ntab[blockIdx.x] = (float)t;
  }
}

// N at log2 t (piecewise linear in log2 t; N ~ t below the grid, constant above)
__device__ __forceinline__ float k4g_n_at(const float* __restrict__ ntab, float lt) {
  const float x = (lt - K4G_LOG2T_MIN) * K4G_PER_OCTAVE;
  if (!(x > 0.f)) return ntab[0] * exp2f(lt - K4G_LOG2T_MIN);
  if (x >= (float)(K4G_TABLE - 1)) return ntab[K4G_TABLE - 1];
  const int k = (int)x;
  const float f = x - (float)k;
  return ntab[k] + f * (ntab[k + 1] - ntab[k]);
}
struct K4gMap {
  const float* vb;       // [K4G_CELLS + 1] bucket coordinate at the cell boundaries (non-decreasing, vb[CELLS] = nb)
  int nb;
  __device__ __forceinline__ int bucket(float lq) const {          // lq < 0
    const float x = (lq + K4G_LQ_SPAN) * ((float)K4G_CELLS / K4G_LQ_SPAN);
    if (!(x > 0.f)) return 0;
    int c = (int)x;
    c = c < K4G_CELLS ? c : K4G_CELLS - 1;                          // x == CELLS only by rounding of lq -> -0
    const float v0 = vb[c], v1 = vb[c + 1];
    const float v = fminf(fmaf(v1 - v0, x - (float)c, v0), v1);     // monotone in lq inside the cell and across cells
    const int bk = (int)v;
    return bk < nb ? bk : nb - 1;
  }
};

// keys are negative floats: ~bits ascends with lq; (key32 << 32 | id) orders by (lq, id)
__device__ __forceinline__ unsigned long long k4g_pack(float lq, uint32_t id) {
  return ((unsigned long long)(~__float_as_uint(lq)) << 32) | id;
}

template <typename OutT>
__global__ void __launch_bounds__(K4G_THREADS, 1)
k4g_order_kernel(const float* __restrict__ logits, int64_t A, uint64_t seed, int64_t sample0, int64_t B,
                 const int32_t* __restrict__ length, const float* __restrict__ ntab, const float* __restrict__ mxp,
                 int nb, unsigned long long* __restrict__ scratch, int64_t scratch_ld,
                 OutT* __restrict__ actions, int64_t ld, int* __restrict__ work, int* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char k4g_smem[];
  uint32_t* off = reinterpret_cast<uint32_t*>(k4g_smem);
  float* vb = reinterpret_cast<float*>(off + nb);                        // [K4G_CELLS + 1] (+3 pad)
  uint32_t* misc = reinterpret_cast<uint32_t*>(vb + K4G_CELLS + 4);      // 64 words: scan partials, counters
  uint32_t* big = misc + 64;
  unsigned long long* tile = reinterpret_cast<unsigned long long*>(big + K4G_BIG_LIST);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
  const int64_t groups = (A + 3) >> 2;
  unsigned long long* my = scratch + (int64_t)blockIdx.x * scratch_ld;
  const float mx = mxp[0];
  K4gMap map{vb, nb};

  for (;;) {
    __syncthreads();
    if (tid == 0) misc[40] = (uint32_t)atomicAdd(work, 1);
    __syncthreads();
    const int64_t b = (int64_t)misc[40];
    if (b >= B) break;
    const int64_t sample = sample0 + b;
    const int64_t T = (int64_t)length[b] - 1;                            // non-terminal ids drawn
    OutT* out = actions + b * ld;
    if (T + 1 > ld || T > scratch_ld || T < 0) {
      if (tid == 0) atomicExch(err, 1);
      continue;
    }
    const float a_term = k4g_arrival(logits, A - 1, sample, seed);

    // ---- per-sample map: vb[c] = nb * N(t at cell boundary c) / N(t_terminal), max-scanned
    for (int i = tid; i < nb; i += K4G_THREADS) off[i] = 0u;
    {
      const float lt_term = (a_term + mx) * 1.4426950408889634f;
      const float n_term = k4g_n_at(ntab, lt_term);
      const float scale = (n_term > 0.f && n_term < 3.0e38f) ? (float)nb / n_term : 0.f;
      float carry = 0.f;
      for (int base = 0; base <= K4G_CELLS; base += K4G_THREADS) {
        const int c = base + tid;
        float v = 0.f;
        if (c < K4G_CELLS) {
          const float lqc = (float)c * (K4G_LQ_SPAN / (float)K4G_CELLS) - K4G_LQ_SPAN;
          v = fminf(k4g_n_at(ntab, fmaf(lqc, 1.4426950408889634f, lt_term)) * scale, (float)nb);
          if (!(v == v)) v = 0.f;
        } else if (c == K4G_CELLS) v = (float)nb;
        float inc = v;                                                   // inclusive max-scan with a carry
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) inc = fmaxf(inc, __shfl_up_sync(0xffffffffu, inc, o, 32) * (lane >= o ? 1.f : 0.f));
        if (lane == 31) reinterpret_cast<float*>(misc)[warp] = inc;
        __syncthreads();
        if (warp == 0) {
          float w = reinterpret_cast<float*>(misc)[lane];
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) w = fmaxf(w, __shfl_up_sync(0xffffffffu, w, o, 32) * (lane >= o ? 1.f : 0.f));
          reinterpret_cast<float*>(misc)[lane] = w;
        }
        __syncthreads();
        const float before = fmaxf(carry, warp ? reinterpret_cast<float*>(misc)[warp - 1] : 0.f);
        if (c <= K4G_CELLS) vb[c] = fmaxf(before, inc);
        carry = fmaxf(carry, reinterpret_cast<float*>(misc)[31]);
        __syncthreads();
      }
    }
    __syncthreads();

    // ---- pass 1: bucket histogram
    for (int64_t g = tid; g < groups; g += K4G_THREADS) {
      float lq[4];
      k4g_group_keys(logits, A, g, sample, seed, a_term, vec_ok, lq);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (lq[j] < 0.f) atomicAdd(&off[map.bucket(lq[j])], 1u);
    }
    __syncthreads();

    // ---- exclusive scan of the histogram (coalesced rounds with a carry)
    uint32_t carry = 0;
    for (int base = 0; base < nb; base += K4G_THREADS) {
      const int i = base + tid;
      const uint32_t v = (i < nb) ? off[i] : 0u;
      uint32_t inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) misc[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        uint32_t w = misc[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
          if (lane >= o) w += t;
        }
        misc[lane] = w;                                                  // inclusive over warps
      }
      __syncthreads();
      const uint32_t before = carry + (warp ? misc[warp - 1] : 0u);
      if (i < nb) off[i] = before + inc - v;
      carry += misc[31];
      __syncthreads();
    }
    if (carry != (uint32_t)T) {                                          // the two passes disagree: cannot happen
      if (tid == 0) atomicExch(err, 2);
      continue;
    }

    // ---- pass 2: scatter (key, id); off[j] ends up as the END of bucket j
    for (int64_t g = tid; g < groups; g += K4G_THREADS) {
      float lq[4];
      k4g_group_keys(logits, A, g, sample, seed, a_term, vec_ok, lq);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (lq[j] < 0.f) {
          const uint32_t pos = atomicAdd(&off[map.bucket(lq[j])], 1u);
          my[pos] = k4g_pack(lq[j], (uint32_t)((g << 2) + j));
        }
    }
    __syncthreads();

    // ---- finish every bucket by rank counting
    int b0 = 0;
    if (tid == 0) misc[41] = 0u;
    while (b0 < nb) {
      const uint32_t e0 = b0 ? off[b0 - 1] : 0u;
      if (e0 >= (uint32_t)T) break;
      int b1 = b0 + K4G_TILE_BUCKETS < nb ? b0 + K4G_TILE_BUCKETS : nb;
      if (off[b1 - 1] - e0 > (uint32_t)K4G_TILE_CAP) {                   // shrink to the buckets that fit
        int lo = b0, hi = b1;                                            // first bucket whose end overflows the tile
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (off[mid] - e0 > (uint32_t)K4G_TILE_CAP) hi = mid; else lo = mid + 1;
        }
        b1 = lo;
      }
      if (b1 == b0) {                                                    // one bucket larger than a tile: rank from L2
        const uint32_t m = off[b0] - e0;
        for (uint32_t i = tid; i < m; i += K4G_THREADS) {
          const unsigned long long ki = my[e0 + i];
          uint32_t rank = 0;
          for (uint32_t k = 0; k < m; ++k) rank += my[e0 + k] < ki;
          out[e0 + rank] = (OutT)(uint32_t)ki;
        }
        b0 += 1;
        continue;
      }
      const uint32_t ne = off[b1 - 1] - e0;
      __syncthreads();                                                   // previous tile consumed
      for (uint32_t i = tid; i < ne; i += K4G_THREADS) tile[i] = my[e0 + i];
      __syncthreads();
      {
        const int j = b0 + (tid >> 1);
        if (j < b1) {
          const uint32_t sj = (j ? off[j - 1] : 0u), m = off[j] - sj;
          const unsigned long long* tb = tile + (sj - e0);
          if (m <= (uint32_t)K4G_BIG) {
            for (uint32_t i = tid & 1; i < m; i += 2) {
              const unsigned long long ki = tb[i];
              uint32_t rank = 0;
              for (uint32_t k = 0; k < m; ++k) rank += tb[k] < ki;
              out[sj + rank] = (OutT)(uint32_t)ki;
            }
          } else if ((tid & 1) == 0) {
            const uint32_t slot = atomicAdd(&misc[41], 1u);
            big[slot] = (uint32_t)j;                                     // <= TILE_CAP / BIG = 80 per tile
          }
        }
      }
      __syncthreads();
      const uint32_t nbig = misc[41];
      for (uint32_t q = 0; q < nbig; ++q) {
        const int j = (int)big[q];
        const uint32_t sj = (j ? off[j - 1] : 0u), m = off[j] - sj;
        const unsigned long long* tb = tile + (sj - e0);
        for (uint32_t i = tid; i < m; i += K4G_THREADS) {
          const unsigned long long ki = tb[i];
          uint32_t rank = 0;
          for (uint32_t k = 0; k < m; ++k) rank += tb[k] < ki;
          out[sj + rank] = (OutT)(uint32_t)ki;
        }
      }
      __syncthreads();
      if (tid == 0) misc[41] = 0u;
      b0 = b1;
    }
    // ---- the terminal id, then -1 padding
    if (tid == 0) out[T] = (OutT)(A - 1);
    for (int64_t p = T + 1 + tid; p < ld; p += K4G_THREADS) out[p] = (OutT)-1;
  }
}

}  // namespace spai
