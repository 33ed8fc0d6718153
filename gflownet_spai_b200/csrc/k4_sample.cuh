// K4 — masked categorical "next nonzero to drop" step of the environment.
//
// Replaces, for B samples at once: the -inf masking of taken ids and softmax of
// policy.py:64-73, the stack / renormalise of gflownet/gflownet.py:88-119, the
// Categorical draw (:148), the chosen-probability gather and -1 bookkeeping of
// gflownet/log.py:67-87 and the terminal test of gflownet.py:177-179.
//
// One block per sample. Three coalesced passes over the A logits (L2-resident:
// the logits vector is shared by all samples of an epoch): running max over
// untaken ids, sum of exp, then an inverse-CDF search for u * sum done warp by
// warp (each warp owns a contiguous range; its total comes from pass 2; the
// owning warp walks its range 32 ids at a time with a shuffle scan). The drawn
// id's bit is set in the sample's taken-mask, so the mask doubles as the state.
#pragma once

#include "spai_internal.cuh"

namespace spai {

constexpr int K4_THREADS = 512;

__device__ __forceinline__ float k4_wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float k4_wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(K4_THREADS)
k4_sample_kernel(const float* __restrict__ logits, int64_t logits_ld, int64_t A,
                 uint32_t* __restrict__ taken, int64_t words_ld,
                 const float* __restrict__ uniforms, uint8_t* __restrict__ done,
                 int64_t* __restrict__ action, float* __restrict__ prob) {
  constexpr int NWARP = K4_THREADS / 32;
  __shared__ float s_red[NWARP];
  __shared__ float s_wsum[NWARP];
  __shared__ float s_bcast[2];
  __shared__ long long s_pick;

  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (done[b]) {
    if (tid == 0) { action[b] = -1; prob[b] = 1.0f; }
    return;
  }
  const float* lg = logits + b * logits_ld;
  uint32_t* tk = taken + b * words_ld;
  // every warp owns the contiguous id range [warp*span, (warp+1)*span), span % 32 == 0
  const int64_t span = ((A + NWARP - 1) / NWARP + 31) / 32 * 32;
  const int64_t lo = warp * span;
  const int64_t hi = (lo + span < A) ? lo + span : A;

  // pass 1: max over untaken ids
  float mx = -INFINITY;
  for (int64_t base = lo; base < hi; base += 32) {
    const int64_t id = base + lane;
    const uint32_t word = tk[base >> 5];                     // warp-uniform
    if (id < hi && !((word >> lane) & 1u)) mx = fmaxf(mx, lg[id]);
  }
  mx = k4_wmax(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    float v = (lane < NWARP) ? s_red[lane] : -INFINITY;
    v = k4_wmax(v);
    if (lane == 0) s_bcast[0] = v;
  }
  __syncthreads();
  mx = s_bcast[0];

  // pass 2: per-warp sums of exp(l - max)
  float sum = 0.f;
  for (int64_t base = lo; base < hi; base += 32) {
    const int64_t id = base + lane;
    const uint32_t word = tk[base >> 5];
    if (id < hi && !((word >> lane) & 1u)) sum += __expf(lg[id] - mx);
  }
  sum = k4_wsum(sum);
  if (lane == 0) s_wsum[warp] = sum;
  if (tid == 0) s_pick = -1;
  __syncthreads();
  float total = 0.f;
#pragma unroll
  for (int w = 0; w < NWARP; ++w) total += s_wsum[w];
  const float target = uniforms[b] * total;

  // locate the owning warp: first warp whose inclusive prefix exceeds target
  float before = 0.f;
  int owner = -1, last_nonempty = -1;
#pragma unroll
  for (int w = 0; w < NWARP; ++w) {
    const float ws = s_wsum[w];
    if (ws > 0.f) last_nonempty = w;
    if (owner < 0) {
      if (before + ws > target && ws > 0.f) owner = w; else before += ws;
    }
  }
  if (owner < 0) { owner = last_nonempty; before = total - s_wsum[owner < 0 ? 0 : owner]; }

  // pass 3: the owning warp walks its range
  if (warp == owner) {
    float run = before;
    long long pick = -1, last_valid = -1;
    float pick_e = 0.f, last_e = 0.f;
    for (int64_t base = lo; base < hi && pick < 0; base += 32) {
      const int64_t id = base + lane;
      const uint32_t word = tk[base >> 5];
      const bool ok = id < hi && !((word >> lane) & 1u);
      const float e = ok ? __expf(lg[id] - mx) : 0.f;
      float inc = e;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const bool cross = ok && (run + inc > target);
      const unsigned cb = __ballot_sync(0xffffffffu, cross);
      const unsigned vb = __ballot_sync(0xffffffffu, ok);
      if (cb) {
        const int src = __ffs(cb) - 1;
        pick = base + src;
        pick_e = __shfl_sync(0xffffffffu, e, src);
      } else if (vb) {
        const int src = 31 - __clz(vb);
        last_valid = base + src;
        last_e = __shfl_sync(0xffffffffu, e, src);
      }
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (pick < 0) { pick = last_valid; pick_e = last_e; }   // rounding: u*total >= running sum
    if (lane == 0) {
      s_pick = pick;
      s_bcast[1] = pick_e;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const long long pick = s_pick;
    action[b] = pick;
    prob[b] = (pick >= 0) ? s_bcast[1] / total : 0.f;
    if (pick >= 0) {
      tk[pick >> 5] |= 1u << (pick & 31);
      if (pick == A - 1) done[b] = 1;
    }
  }
}

// K4b — whole-trajectory sampling support (Gumbel-top-k / Plackett-Luce): drawing
// ids one at a time from softmax(logits) without replacement until the terminal
// id appears is, in distribution, the same as sorting the ids by
// key_i = logit_i + Gumbel_i and cutting at the terminal's key. The set of
// removed edges of sample b is therefore {i : key[b,i] > key[b,A-1]}: this kernel
// packs that set (plus the terminal bit) into the sample's taken-bitmask with
// warp ballots and counts it. One block row per sample, coalesced key reads.
__global__ void __launch_bounds__(256)
k4_pack_taken_kernel(const float* __restrict__ keys, int64_t ld, int64_t A,
                     uint32_t* __restrict__ taken, int64_t words_ld, int32_t* __restrict__ length) {
  const int64_t b = blockIdx.x;
  const float* kb = keys + b * ld;
  const float thr = kb[A - 1];
  const int lane = threadIdx.x & 31;
  int cnt = 0;
  const int64_t words = (A + 31) >> 5;
  for (int64_t w = threadIdx.x >> 5; w < words; w += blockDim.x >> 5) {
    const int64_t i = w * 32 + lane;
    const bool hit = i < A && (kb[i] > thr || i == A - 1);
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) taken[b * words_ld + w] = bal;
    cnt += __popc(bal);
  }
  __shared__ int part[8];
  if (lane == 0) part[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    length[b] = t;                       // ids drawn, terminal included
  }
}

}  // namespace spai
