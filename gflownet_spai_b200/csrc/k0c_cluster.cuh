// K0c — action lists -> kept-edge bitmasks in ONE pass for patterns whose bitmask does not fit
// one CTA's shared memory (E > 819 200 slots: BASELINE configs 3, 4, 5), by a thread-block
// cluster that holds the whole bitmask of a trajectory in DISTRIBUTED shared memory.
//
// Replaces gflownet/utils.py:315-323 like K0 (k0_masks.cuh) and K0b (k0b_bucket.cuh). K0b needs
// two kernels (sort by 64 K-slot segment, then build) and moves 12 bytes per 8-byte id through
// HBM with one shared-memory atomic per id in EACH pass; the round-1 cluster experiment cleared
// bits with remote shared-memory atomics, one 4-byte DSMEM transaction per id, and lost to L2
// atomics. Here the ids cross the cluster in bulk:
//
//   * cluster of cs <= 8 CTAs per trajectory; CTA r owns slots [r*S, (r+1)*S), S = spc * 65 536,
//     as a bitmask slice in its own shared memory (spc * 8 KB, <= 184 KB);
//   * every CTA streams its share of the trajectory (2048-id chunks, coalesced, read ONCE from
//     HBM, prefetched into L2 two rounds ahead), splits each chunk by OWNER (<= 8 buckets: packed
//     16-bit counters + one warp scan, no atomics), stages the runs in its shared memory and warp k
//     pushes run k with coalesced 16-byte st.shared::cluster into a fixed slot of owner k's inbox;
//     a run longer than the slot (capacity = mean + 4..6 sigma of a random trajectory) clears its
//     excess bits with remote red.and — exact for any input, slow only for adversarially sorted ones;
//   * the owner clears the bits of everything in its inbox with LOCAL shared-memory atomics;
//   * hand-shake per (group, round) through two mbarriers per CTA with remote arrivals
//     (full: every source has pushed — the one cluster-scope release, by one lane per run, at a
//     point where that thread has no id loads in flight; empty: every owner has consumed). The
//     round is software-pipelined: split(j) -> consume(j-1) -> push(j), so a push has a whole
//     split phase to land before its owner looks at it. A CTA runs NH independent groups of 256
//     threads on alternating rounds, so one group's split overlaps another's atomics;
//   * after the last round: cluster barrier, coalesced write of the slice, popcount fused.
//
// HBM traffic per id: the 8 (4) bytes of the id, once. Shared-memory atomics per id: one.
#pragma once

#include "k3_copy.cuh"
#include "spai_internal.cuh"

namespace spai {

constexpr int K0C_SEG_SHIFT = 16;                        // slice granularity: 65 536 slots = 8 KB of bitmask
constexpr int K0C_SEG_WORDS = 1 << (K0C_SEG_SHIFT - 5);
constexpr int K0C_HT = 256;                              // threads of one group
constexpr int K0C_HW = K0C_HT / 32;                      // 8 warps: warp k pushes the run of owner k
constexpr int K0C_IDS = 8;                               // ids per thread and round
constexpr int K0C_CHUNK = K0C_HT * K0C_IDS;              // 2048 ids per (CTA, group, round)
constexpr int K0C_MAX_CS = 8;                            // portable cluster size
constexpr int K0C_MAX_SMEM = 232448 - 1536;              // 227 KB minus the static part (barriers, popcount partials)

// per-group scratch behind the slice: inbox [cs][4 + cap] (slot = 16-byte header {count} + ids), staging buffer,
// warp totals [HW][4]
constexpr int K0C_STAGE_WORDS = K0C_CHUNK + 8 * K0C_MAX_CS;      // cs runs, each: 4-word header + ids padded to 4 words
__host__ __device__ inline int k0c_group_words(int cs, int cap) { return cs * (cap + 4) + K0C_STAGE_WORDS + K0C_HW * 4; }
inline size_t k0c_smem(int spc, int cs, int cap, int nh) {
  return (size_t)spc * K0C_SEG_WORDS * 4 + (size_t)nh * k0c_group_words(cs, cap) * 4;
}
// inbox slot of one (source, owner) pair: mean + sigmas * sigma of a binomial(CHUNK, 1/cs), a multiple of 4
inline int k0c_cap(int cs, double sigmas) {
  const double mean = (double)K0C_CHUNK / cs;
  int cap = (int)(mean + sigmas * sqrt(mean * (1.0 - 1.0 / cs)) + 4.0);
  cap = (cap + 3) & ~3;
  return cap > K0C_CHUNK ? K0C_CHUNK : cap;
}

__device__ __forceinline__ uint32_t k0c_mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void k0c_st_remote4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void k0c_red_and_remote(uint32_t addr, uint32_t v) {
  asm volatile("red.relaxed.cluster.shared::cluster.and.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// full signal: publishes the warp's remote stores (cluster-scope release = MEMBAR.ALL.GPU, which waits for everything
// the thread has in flight: it is issued where the next round's id loads have had a whole split phase to land)
__device__ __forceinline__ void k0c_arrive_remote_release(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
// empty signal: follows reads only, no cluster-scope release needed
__device__ __forceinline__ void k0c_arrive_remote(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void k0c_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void k0c_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void k0c_group_sync(int g) {
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(K0C_HT) : "memory");
}

// grid = B * cs CTAs, cluster = cs consecutive CTAs = one trajectory. inv = ceil(65536 / spc):
// owner of segment s is (s * inv) >> 16 (exact for s < 2048).
template <typename IdT, bool HAS_MAP, int NH>
__global__ void __launch_bounds__(NH * K0C_HT, 1)
k0c_cluster_kernel(const IdT* __restrict__ actions, int64_t T, int64_t ld, const int32_t* __restrict__ row_len,
                   const int32_t* __restrict__ edge_slot, int64_t E, int cs, int spc, int cap, uint32_t inv,
                   uint32_t* __restrict__ mask, int64_t W, unsigned long long* __restrict__ nnz) {
  extern __shared__ __align__(16) uint32_t k0c_sm[];
  __shared__ __align__(8) uint64_t bars[2 * NH];              // full[g], empty[g]
  __shared__ long long part[NH * K0C_HW];
  constexpr int THREADS = NH * K0C_HT;
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid / K0C_HT, ht = tid % K0C_HT, hw = ht >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int64_t b = blockIdx.x / cs;

  const int slice_words = spc * K0C_SEG_WORDS;
  const int slot = cap + 4;
  uint32_t* bits = k0c_sm;
  uint32_t* inbox = k0c_sm + slice_words + g * k0c_group_words(cs, cap);
  uint32_t* stage = inbox + cs * slot;
  uint32_t* wtot = stage + K0C_STAGE_WORDS;
  const int64_t w0 = (int64_t)rank * slice_words;
  const int nw = (int)max((int64_t)0, min((int64_t)slice_words, W - w0));
  const uint32_t tail = (E & 31) ? ((1u << (E & 31)) - 1u) : 0xffffffffu;
  for (int w = tid; w < nw; w += THREADS) bits[w] = (w0 + w == W - 1) ? tail : 0xffffffffu;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2 * NH; ++i) mbar_init(&bars[i], (uint32_t)cs);
    mbar_fence_init();
  }
  k0c_cluster_sync();                                         // slices and barriers of every CTA are ready

  int64_t len = T;
  if (row_len) len = min(T, (int64_t)row_len[b]);
  const int64_t nq = (len + K0C_CHUNK - 1) / K0C_CHUNK;       // chunks of the trajectory
  const int64_t G = (nq + cs - 1) / cs;                       // rounds: round r = chunks [r*cs, (r+1)*cs), one per CTA
  const IdT* row = actions + b * ld;
  const uint32_t Eu = (uint32_t)E;                            // E < 2^31 (checked at context creation)
  const uint32_t S = (uint32_t)spc << K0C_SEG_SHIFT;

  // warp hw of a group pushes the run of owner hw: my slot in its inbox, its full barrier, its slice
  uint32_t r_slot = 0, r_bits = 0, r_full = 0, r_empty = 0;
  if (hw < cs) {
    r_slot = k0c_mapa(smem_u32(inbox + rank * slot), (uint32_t)hw);
    r_bits = k0c_mapa(smem_u32(bits), (uint32_t)hw);
    r_full = k0c_mapa(smem_u32(&bars[2 * g]), (uint32_t)hw);
  }
  if (ht < cs) r_empty = k0c_mapa(smem_u32(&bars[2 * g + 1]), (uint32_t)ht);

  IdT v[K0C_IDS];
  auto load = [&](int64_t r) {
    const int64_t c0 = (r * cs + rank) * K0C_CHUNK + ht;
#pragma unroll
    for (int u = 0; u < K0C_IDS; ++u) {
      const int64_t t = c0 + u * K0C_HT;
      v[u] = (t < len) ? __ldcs(row + t) : (IdT)-1;
    }
  };
  // HBM -> L2 ahead of the register loads (which then hit L2): no registers, no fence ever waits for it
  auto prefetch = [&](int64_t r) {
    if (ht != 0 || r >= G) return;
    const int64_t c0 = (r * cs + rank) * K0C_CHUNK;
    const int64_t c1 = min(len, c0 + K0C_CHUNK);
    uintptr_t a0 = (reinterpret_cast<uintptr_t>(row + c0) + 15) & ~(uintptr_t)15;
    const uintptr_t a1 = reinterpret_cast<uintptr_t>(row + c1) & ~(uintptr_t)15;
    if (c1 > c0 && a1 > a0) k0c_prefetch_l2(reinterpret_cast<const void*>(a0), (uint32_t)(a1 - a0));
  };
  auto consume = [&](uint32_t jj) {                           // round jj of this group: every source's run has landed
    mbar_wait(&bars[2 * g], jj & 1u);
    for (int s = 0; s < cs; ++s) {
      const uint32_t* in = inbox + s * slot;
      const uint32_t n = in[0];
      for (uint32_t i = ht; i < n; i += K0C_HT) {
        const uint32_t l = in[4 + i];
        atomicAnd(bits + (l >> 5), ~(1u << (l & 31u)));
      }
    }
  };
  int64_t r = g;                                              // group g takes rounds g, g + NH, ...
  if (r < G) load(r);
  prefetch(r + NH);
  prefetch(r + 2 * NH);
  uint32_t j = 0;
  for (; r < G; r += NH, ++j) {
    // ---- split by owner: key = invalid << 31 | owner << 24 | local slot; per-thread 4-bit counters (<= 8 ids per owner)
    uint32_t key[K0C_IDS];
    uint32_t c = 0;
#pragma unroll
    for (int u = 0; u < K0C_IDS; ++u) {
      bool ok;
      uint32_t lo;
      if (sizeof(IdT) == 8) {
        const uint64_t a = (uint64_t)v[u];
        lo = (uint32_t)a;
        ok = ((uint32_t)(a >> 32) == 0u) & (lo < Eu);
      } else {
        lo = (uint32_t)v[u];
        ok = lo < Eu;
      }
      uint32_t s = lo;
      if (HAS_MAP) s = (uint32_t)__ldg(edge_slot + (ok ? lo : 0u));
      const uint32_t o = (((s >> K0C_SEG_SHIFT) * inv) >> 16) & 7u;
      const uint32_t l = s - o * S;
      key[u] = ok ? (l | (o << 24)) : 0x80000000u;
      c += ok ? (1u << (o << 2)) : 0u;
    }
    if (r + NH < G) load(r + NH);                             // next round's ids (L2 hits) fly during the rest of this one
    prefetch(r + 3 * NH);

    // ---- ranks: 4-bit fields -> 16-bit fields (two owners per register), inclusive warp scan (<= 256 per field)
    const uint32_t c0 = (c & 0xFu) | ((c & 0xF0u) << 12);
    const uint32_t c1 = ((c >> 8) & 0xFu) | ((c & 0xF000u) << 4);
    const uint32_t c2 = ((c >> 16) & 0xFu) | ((c >> 4) & 0xF0000u);
    const uint32_t c3 = ((c >> 24) & 0xFu) | ((c >> 12) & 0xF0000u);
    uint32_t i0 = c0, i1 = c1, i2 = c2, i3 = c3;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x0 = __shfl_up_sync(0xffffffffu, i0, o), x1 = __shfl_up_sync(0xffffffffu, i1, o);
      const uint32_t x2 = __shfl_up_sync(0xffffffffu, i2, o), x3 = __shfl_up_sync(0xffffffffu, i3, o);
      if (lane >= o) { i0 += x0; i1 += x1; i2 += x2; i3 += x3; }
    }
    if (lane == 31) *reinterpret_cast<uint4*>(wtot + hw * 4) = make_uint4(i0, i1, i2, i3);
    // ids of the lanes below me, per owner (<= 248): 8-bit fields, owners 0-3 / 4-7
    const uint32_t ex_lo = __byte_perm(i0 - c0, i1 - c1, 0x6420), ex_hi = __byte_perm(i2 - c2, i3 - c3, 0x6420);
    k0c_group_sync(g);                                        // also: the previous push has read `stage`
    // lane L < 8: run of owner L = 4-word header + ids (padded to 4 words) at `off`, this warp's ids from off + 4 + pre
    uint32_t pre = 0, tot = 0;
    if (lane < 8) {
#pragma unroll
      for (int w = 0; w < K0C_HW; ++w) {
        const uint32_t x = (wtot[w * 4 + (lane >> 1)] >> ((lane & 1) << 4)) & 0xffffu;
        pre += (w < hw) ? x : 0u;
        tot += x;
      }
    }
    const uint32_t padded = (lane < 8) ? 4u + ((tot + 3u) & ~3u) : 0u;
    uint32_t run = padded;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, run, o);
      if (lane >= o) run += x;
    }
    const uint32_t off = run - padded;
    const uint32_t wbase = off + 4u + pre;
    uint32_t rc = 0;                                          // ids of this thread already placed, per owner
#pragma unroll
    for (int u = 0; u < K0C_IDS; ++u) {
      const uint32_t k = key[u];
      const uint32_t o = (k >> 24) & 7u, sh4 = o << 2;
      const uint32_t exw = (o & 4u) ? ex_hi : ex_lo;
      const uint32_t excl = (exw >> ((o & 3u) << 3)) & 0xFFu;
      const uint32_t mine = (rc >> sh4) & 0xFu;
      const uint32_t base = __shfl_sync(0xffffffffu, wbase, (int)o);
      const bool ok = (int32_t)k >= 0;
      const uint32_t pos = ok ? base + excl + mine : (uint32_t)(K0C_STAGE_WORDS - 1);   // invalid ids: a dummy word
      stage[pos] = k & 0xffffffu;
      rc += ok ? (1u << sh4) : 0u;
    }
    const uint32_t n_run = __shfl_sync(0xffffffffu, tot, hw), o_run = __shfl_sync(0xffffffffu, off, hw);
    const uint32_t m_run = min(n_run, (uint32_t)cap);
    if (hw < cs && lane == 0) stage[o_run] = m_run;           // header of the run this warp will push
    k0c_group_sync(g);                                        // `stage` holds this round's runs

    // ---- the previous round's runs have had this whole split phase to land in my inbox
    if (j > 0) {
      consume(j - 1);
      k0c_group_sync(g);
      if (ht < cs) k0c_arrive_remote(r_empty);
    }

    // ---- push run hw to owner hw once every owner has consumed what this group sent in its previous round
    if (hw < cs) {
      if (j > 0) mbar_wait(&bars[2 * g + 1], (j - 1) & 1u);
      const uint32_t nv = 1u + ((m_run + 3u) >> 2);           // 16-byte vectors: header + ids
      const uint4* src = reinterpret_cast<const uint4*>(stage + o_run);
      for (uint32_t i = lane; i < nv; i += 32) k0c_st_remote4(r_slot + i * 16u, src[i]);
      for (uint32_t i = m_run + lane; i < n_run; i += 32) {    // over the slot's capacity: remote atomics
        const uint32_t l = stage[o_run + 4u + i];
        k0c_red_and_remote(r_bits + (l >> 5) * 4u, ~(1u << (l & 31u)));
      }
      __syncwarp();
      if (lane == 0) k0c_arrive_remote_release(r_full);
    }
  }
  if (j > 0) consume(j - 1);
  k0c_cluster_sync();                                         // every push, remote atomic and local atomic has landed

  long long cntv = 0;
  uint32_t* out = mask + b * W + w0;
  for (int w = tid; w < nw; w += THREADS) {
    const uint32_t x = bits[w];
    out[w] = x;
    cntv += __popc(x);
  }
  for (int o = 16; o; o >>= 1) cntv += __shfl_xor_sync(0xffffffffu, cntv, o);
  if (lane == 0) part[tid >> 5] = cntv;
  __syncthreads();
  if (tid == 0 && nnz) {
    long long t = 0;
    for (int i = 0; i < NH * K0C_HW; ++i) t += part[i];
    atomicAdd(nnz + b, (unsigned long long)t);
  }
}

}  // namespace spai
