// K4p — many masked-categorical environment steps per launch, with running sums.
//
// Same step as K4 (policy.py:64-73 masked softmax, gflownet/gflownet.py:148 draw, :177-179 terminal
// test, log.py:67-87 bookkeeping), but one warp owns one sample for the whole launch and keeps, in
// shared memory, S[j] = sum of exp(l - max) over the untaken ids of block j (~sqrt(A) ids per block) and,
// in registers, one chunk sum per lane. A step is then: scan 32 chunk sums -> scan one chunk's block
// sums -> scan ONE block's logits (L2) -> draw -> refresh that block's and chunk's sums from scratch
// (no drift: a sum is never updated by subtraction across steps). O(sqrt(A)) per step instead of
// K4's three passes over all A logits. Uniforms are injected (tests, f32[nsteps, B]) or Philox4x32-10
// (key = seed ^ K4P tag, counter = (step / 4, sample)).
//
// The inverse-CDF order is id order at every level (chunks, blocks, lanes own contiguous ids), so a
// draw satisfies the same interval property as K4's: cdf[x-1] <= u * total <= cdf[x] (fp32 sums).
#pragma once

#include "k4g_gumbel.cuh"
#include "spai_internal.cuh"

namespace spai {

constexpr int K4P_WARPS = 4;
constexpr int K4P_MAX_NBLK = 2048;

__device__ __forceinline__ float k4p_wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float k4p_scan(float v, int lane) {      // inclusive
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

struct K4pState {
  const float* lg;
  uint32_t* tk;
  float* S;
  int64_t A, words;
  int blk, per, nblk, c, lane;       // per = ids per lane of a block (a multiple of 4)
  bool vec_ok;
  float mx, C;

  // bit j = id0 + j is untaken and < A (cnt <= 32 ids)
  __device__ __forceinline__ uint32_t untaken_bits(int64_t id0, int cnt) const {
    const int64_t w = id0 >> 5;
    const int sh = (int)(id0 & 31);
    const uint32_t w0 = (w < words) ? tk[w] : 0xffffffffu;
    const uint32_t w1 = (sh && w + 1 < words) ? tk[w + 1] : 0xffffffffu;
    uint32_t bits = ~__funnelshift_r(w0, w1, sh);
    if (cnt < 32) bits &= (1u << cnt) - 1u;
    const int64_t room = A - id0;
    if (room < 32) bits &= room <= 0 ? 0u : ((1u << (int)room) - 1u);
    return bits;
  }

  // sum of exp(l - mx) over this lane's contiguous ids of the block starting at `base`: the mask words
  // and the logits are loaded unconditionally (independent 16-byte loads in flight), then selected
  __device__ __forceinline__ float lane_sum(int64_t base) const {
    const int64_t lo = base + (int64_t)lane * per;
    float s = 0.f;
    for (int i0 = 0; i0 < per; i0 += 32) {
      const int cnt = per - i0 < 32 ? per - i0 : 32;
      const int64_t id0 = lo + i0;
      const uint32_t bits = untaken_bits(id0, cnt);
#pragma unroll 4
      for (int j = 0; j < cnt; j += 4) {
        float4 v;
        if (vec_ok && id0 + j + 3 < A) v = *reinterpret_cast<const float4*>(lg + id0 + j);
        else {
          v.x = id0 + j < A ? lg[id0 + j] : 0.f;
          v.y = id0 + j + 1 < A ? lg[id0 + j + 1] : 0.f;
          v.z = id0 + j + 2 < A ? lg[id0 + j + 2] : 0.f;
          v.w = id0 + j + 3 < A ? lg[id0 + j + 3] : 0.f;
        }
        const uint32_t nb = bits >> j;
        s += (nb & 1u) ? __expf(v.x - mx) : 0.f;
        s += (nb & 2u) ? __expf(v.y - mx) : 0.f;
        s += (nb & 4u) ? __expf(v.z - mx) : 0.f;
        s += (nb & 8u) ? __expf(v.w - mx) : 0.f;
      }
    }
    return s;
  }

  // max over the untaken logits, then every block sum and this lane's chunk sum
  __device__ void rebuild() {
    float m = -INFINITY;
    for (int64_t base = 0; base < A; base += 32) {
      const int64_t id = base + lane;
      const uint32_t word = tk[base >> 5];
      if (id < A && !((word >> lane) & 1u)) m = fmaxf(m, lg[id]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    mx = m;
    for (int j = 0; j < nblk; ++j) {
      const float s = k4p_wsum(lane_sum((int64_t)j * blk));
      if (lane == 0) S[j] = s;
    }
    __syncwarp();
    float cs = 0.f;
    for (int j = 0; j < c; ++j) cs += S[lane * c + j];
    C = cs;
  }
};

template <typename OutT>
__global__ void __launch_bounds__(K4P_WARPS * 32)
k4p_steps_kernel(const float* __restrict__ logits, int64_t A, uint32_t* __restrict__ taken, int64_t words_ld,
                 uint8_t* __restrict__ done, const float* __restrict__ uniforms, uint64_t seed, int64_t sample0,
                 int64_t step0, int64_t B, int64_t nsteps, int blk, int nblk,
                 OutT* __restrict__ actions, float* __restrict__ probs, int64_t ld, int32_t* __restrict__ steps_taken) {
  extern __shared__ float k4p_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * K4P_WARPS + warp;
  if (b >= B) return;                                   // warps are independent: no block-level barrier below
  OutT* act = actions + b * ld + step0;
  float* pr = probs ? probs + b * ld + step0 : nullptr;
  int64_t s = 0;
  if (!done[b]) {
    K4pState st;
    st.lg = logits; st.tk = taken + b * words_ld; st.S = k4p_smem + (size_t)warp * nblk; st.A = A;
    st.words = (A + 31) >> 5;
    st.blk = blk; st.per = blk / 32; st.nblk = nblk; st.c = nblk / 32; st.lane = lane;
    st.vec_ok = (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
    st.rebuild();
    uint4 rnd = make_uint4(0, 0, 0, 0);
    bool finished = false;
    for (; s < nsteps && !finished; ++s) {
      float total = k4p_wsum(st.C);
      if (!(total > 0.f)) {                             // the remaining ids underflow against the old maximum
        st.rebuild();
        total = k4p_wsum(st.C);
        if (!(total > 0.f)) break;                      // non-finite logits: leave the sample unfinished (caller checks)
      }
      float u;
      if (uniforms) u = uniforms[s * B + b];
      else {
        const int64_t gs = step0 + s;
        if ((gs & 3) == 0 || s == 0)
          rnd = philox4x32_10(make_uint4((uint32_t)(gs >> 2), (uint32_t)((uint64_t)(gs >> 2) >> 32), (uint32_t)(sample0 + b),
                                         (uint32_t)((uint64_t)(sample0 + b) >> 32)),
                              make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x4B345053u));
        const uint32_t x = (gs & 3) == 0 ? rnd.x : (gs & 3) == 1 ? rnd.y : (gs & 3) == 2 ? rnd.z : rnd.w;
        u = (float)(x >> 8) * 5.9604644775390625e-8f;   // [0, 1)
      }
      const float target = u * total;
      // ---- chunk
      const float incC = k4p_scan(st.C, lane);
      unsigned bal = __ballot_sync(0xffffffffu, incC > target && st.C > 0.f);
      const unsigned nzC = __ballot_sync(0xffffffffu, st.C > 0.f);
      const int k = bal ? __ffs(bal) - 1 : 31 - __clz(nzC);
      const float t1 = target - __shfl_sync(0xffffffffu, incC - st.C, k);
      // ---- block inside chunk k
      int bj = -1, last_bj = -1;
      float before2 = 0.f, last_before = 0.f, run = 0.f;
      for (int j0 = 0; j0 < st.c && bj < 0; j0 += 32) {
        const float v = (j0 + lane < st.c) ? st.S[k * st.c + j0 + lane] : 0.f;
        const float inc = k4p_scan(v, lane);
        bal = __ballot_sync(0xffffffffu, run + inc > t1 && v > 0.f);
        const unsigned nz = __ballot_sync(0xffffffffu, v > 0.f);
        if (bal) {
          const int src = __ffs(bal) - 1;
          bj = k * st.c + j0 + src;
          before2 = run + __shfl_sync(0xffffffffu, inc - v, src);
        } else if (nz) {
          const int src = 31 - __clz(nz);
          last_bj = k * st.c + j0 + src;
          last_before = run + __shfl_sync(0xffffffffu, inc - v, src);
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (bj < 0) { bj = last_bj; before2 = last_before; }
      const float t2 = t1 - before2;
      // ---- id inside block bj
      const int64_t base = (int64_t)bj * st.blk;
      const float ls = st.lane_sum(base);
      const float incL = k4p_scan(ls, lane);
      bal = __ballot_sync(0xffffffffu, incL > t2 && ls > 0.f);
      const unsigned nzL = __ballot_sync(0xffffffffu, ls > 0.f);
      if (!nzL) {                                       // stale sum (cannot happen: sums are exact zeros when empty)
        if (lane == 0) st.S[bj] = 0.f;
        __syncwarp();
        float cs = 0.f;
        for (int j = lane; j < st.c; j += 32) cs += st.S[k * st.c + j];
        cs = k4p_wsum(cs);
        if (lane == k) st.C = cs;
        --s;
        continue;
      }
      const int o = bal ? __ffs(bal) - 1 : 31 - __clz(nzL);
      const float t3 = t2 - __shfl_sync(0xffffffffu, incL - ls, o);
      // the whole warp walks the owner lane's ids, 32 per round
      long long pick = -1, lastv = -1;
      float pick_e = 0.f, last_e = 0.f, acc = 0.f;
      const int64_t lo_o = base + (int64_t)o * st.per;
      for (int r0 = 0; r0 < st.per && pick < 0; r0 += 32) {
        const int64_t id = lo_o + r0 + lane;
        const bool ok = r0 + lane < st.per && id < A && !((st.tk[id >> 5] >> (id & 31)) & 1u);
        const float e = ok ? __expf(st.lg[id] - st.mx) : 0.f;
        const float inc = k4p_scan(e, lane);
        const unsigned cb = __ballot_sync(0xffffffffu, acc + inc > t3 && e > 0.f);
        const unsigned vb = __ballot_sync(0xffffffffu, e > 0.f);
        if (cb) {
          const int src = __ffs(cb) - 1;
          pick = lo_o + r0 + src;
          pick_e = __shfl_sync(0xffffffffu, e, src);
        } else if (vb) {
          const int src = 31 - __clz(vb);
          lastv = lo_o + r0 + src;
          last_e = __shfl_sync(0xffffffffu, e, src);
        }
        acc += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (pick < 0) { pick = lastv; pick_e = last_e; }
      if (lane == 0) {
        act[s] = (OutT)pick;
        if (pr) pr[s] = pick_e / total;
        st.tk[pick >> 5] |= 1u << (pick & 31);
      }
      // ---- refresh the block's and the chunk's sums
      const float ns = k4p_wsum(lane == o ? fmaxf(ls - pick_e, 0.f) : ls);
      if (lane == 0) st.S[bj] = ns;
      __syncwarp();
      float cs = 0.f;
      for (int j = lane; j < st.c; j += 32) cs += st.S[k * st.c + j];
      cs = k4p_wsum(cs);
      if (lane == k) st.C = cs;
      if (pick == A - 1) finished = true;
    }
    if (finished && lane == 0) done[b] = 1;
  }
  if (steps_taken && lane == 0) steps_taken[b] = (int32_t)s;
  for (int64_t q = s + lane; q < nsteps; q += 32) {     // finished rows: -1 / 1.0 (log.py:67,78,84-86)
    act[q] = (OutT)-1;
    if (pr) pr[q] = 1.0f;
  }
}

}  // namespace spai
