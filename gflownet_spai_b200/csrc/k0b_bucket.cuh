// K0b — action lists -> kept-edge bitmasks for patterns whose bitmask does not fit one
// CTA's shared memory (E > 819 200 slots: BASELINE configs 3, 4, 5).
//
// Replaces gflownet/utils.py:315-323 like K0 (k0_masks.cuh). Round 1 cleared one bit per
// deletion with a global RED.AND: one 32-byte L2 transaction per id, 67 G ids/s = 8 % of
// the HBM rate of the action stream. Here no id ever touches L2 atomics:
//
//   pass 1  k0b_sort_kernel    one CTA per chunk of 8192 ids of one trajectory: ids are
//           mapped to slots, split into SEGMENTS of 65 536 slots (8 KB of bitmask), and the
//           chunk is counting-sorted by segment in shared memory (rank = returning
//           shared-memory atomic on a per-warp counter, so lanes of a warp rarely collide:
//           32 lanes over C >= 13 segments). The sorted chunk is written back as 16-BIT
//           local ids (2 bytes per id instead of the 8 read) plus its C+1 segment offsets.
//           Ids outside [0, E) (-1 padding, the terminal id) are dropped here.
//   pass 2  k0b_build_kernel   one CTA per (trajectory, R consecutive segments): the R*8 KB
//           slice of the bitmask lives in shared memory exactly as in
//           k0_mask_build_smem_kernel; each warp walks the chunks of the trajectory, reads the
//           contiguous run of the CTA's segments and clears bits with shared-memory atomics;
//           the finished slice is written with coalesced stores, popcount fused.
//
// DRAM traffic per id: 8 (read) + 2 (write) + 2 (read) + headers ~0.1 = ~12.1 bytes against
// 8 compulsory; with the batch processed in groups whose staged ids fit L2 the 2+2 bytes
// never reach DRAM (host loop in spai_b200.cu, SPAI_K0B_GROUP).
#pragma once

#include "spai_internal.cuh"

namespace spai {

constexpr int K0B_SEG_SHIFT = 16;                       // 65 536 slots per segment
constexpr int K0B_SEG_WORDS = 1 << (K0B_SEG_SHIFT - 5); // 2048 mask words = 8 KB
constexpr int K0B_IDS = 16;                             // ids per thread of a sort CTA
constexpr int K0B_THREADS = 512;                        // build CTA; default sort CTA (chunk = 16 * threads = 8192 ids)
constexpr int K0B_WARPS = K0B_THREADS / 32;
constexpr int K0B_CHUNK_MAX = K0B_THREADS * K0B_IDS;
constexpr int K0B_MAX_SEGS = 512;                       // E <= 33.5 M slots
constexpr int K0B_R = 8;                                // segments per build CTA (64 KB of shared memory)

inline size_t k0b_sort_smem(int C, int threads) {
  return (size_t)(threads / 32) * (C + 1) * 4 + (size_t)threads * K0B_IDS * 2 + 128;
}

// hdr u16[B][nchunks][C + 1]: hdr[..][s] = first position of segment s inside the sorted
// chunk, hdr[..][C] = number of valid ids of the chunk. Chunks at or beyond the row's length
// are not written (pass 2 derives the same chunk count from the same length).
// Straight-line code: every id takes the same path (ids that match no edge are counted in a
// trash bucket C of the warp and never stored), the edge -> slot map is a template switch.
template <typename IdT, bool HAS_MAP, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k0b_sort_kernel(const IdT* __restrict__ actions, int64_t T, int64_t ld,
                const int32_t* __restrict__ row_len, const int32_t* __restrict__ edge_slot,
                int64_t E, int C, uint16_t* __restrict__ stage, int64_t ld_stage,
                uint16_t* __restrict__ hdr, int64_t nchunks) {
  constexpr int NWARPS = THREADS / 32, CHUNK = THREADS * K0B_IDS;
  extern __shared__ __align__(16) uint32_t k0b_sm[];
  const int C1 = C + 1;                                           // + trash bucket
  uint32_t* cnt = k0b_sm;                                         // [WARPS][C1] counters, then bases
  uint16_t* stg = reinterpret_cast<uint16_t*>(k0b_sm + ((NWARPS * C1 + 3) & ~3));
  __shared__ uint32_t wtot[NWARPS];
  const int64_t b = blockIdx.x / nchunks;
  const int64_t chunk = blockIdx.x % nchunks;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t len = T;
  if (row_len) len = min(T, (int64_t)row_len[b]);
  const int64_t c0 = chunk * CHUNK;
  if (c0 >= len) return;
  const IdT* row = actions + b * ld + c0;
  const int left = (int)min(len - c0, (int64_t)CHUNK);        // ids of this chunk

  IdT v[K0B_IDS];
  if (left == CHUNK) {
#pragma unroll
    for (int u = 0; u < K0B_IDS; ++u) v[u] = __ldcs(row + u * THREADS + tid);
  } else {
#pragma unroll
    for (int u = 0; u < K0B_IDS; ++u) {
      const int t = u * THREADS + tid;
      v[u] = (t < left) ? __ldcs(row + t) : (IdT)-1;
    }
  }
  for (int i = tid; i < NWARPS * C1; i += THREADS) cnt[i] = 0;
  __syncthreads();

  uint32_t key[K0B_IDS];       // slot; 0xffffffff (segment = trash) for ids that match no edge
  uint32_t pos[K0B_IDS];       // rank inside the (warp, segment) sub-list
  uint32_t* wc = cnt + warp * C1;
  const uint32_t Eu = (uint32_t)E;                                // E < 2^31 (checked at context creation)
#pragma unroll
  for (int u = 0; u < K0B_IDS; ++u) {
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
    const uint32_t seg = ok ? (s >> K0B_SEG_SHIFT) : (uint32_t)C;
    key[u] = ok ? s : 0xffffffffu;
    pos[u] = atomicAdd(wc + seg, 1u);
  }
  __syncthreads();

  // exclusive scan in (segment, warp) order: thread s owns segment s (C <= THREADS); the trash
  // bucket is left out
  uint32_t tot = 0;
  if (tid < C) {
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) tot += cnt[w * C1 + tid];
  }
  uint32_t inc = tot;                                             // inclusive scan over segments
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += x;
  }
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < NWARPS; ++w) {
    const uint32_t x = wtot[w];
    if (w < warp) wbase += x;
    total += x;
  }
  const uint32_t segbase = wbase + inc - tot;
  uint16_t* h = hdr + (b * nchunks + chunk) * (int64_t)C1;
  if (tid < C) {
    uint32_t run = segbase;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) {
      const uint32_t x = cnt[w * C1 + tid];
      cnt[w * C1 + tid] = run;
      run += x;
    }
    h[tid] = (uint16_t)segbase;
  }
  if (tid == 0) h[C] = (uint16_t)total;
  if (tid < NWARPS) cnt[tid * C1 + C] = (uint32_t)CHUNK;    // trash ids land behind the chunk (never copied out)
  __syncthreads();

#pragma unroll
  for (int u = 0; u < K0B_IDS; ++u) {
    const uint32_t s = key[u];
    const uint32_t seg = min(s >> K0B_SEG_SHIFT, (uint32_t)C);
    const uint32_t at = wc[seg] + ((s == 0xffffffffu) ? 0u : pos[u]);
    stg[at] = (uint16_t)(s & 0xffffu);
  }
  __syncthreads();

  // sorted chunk -> stage[b][c0 ..): 16-byte stores (ld_stage is a multiple of 8, c0 of 8192)
  uint4* dst = reinterpret_cast<uint4*>(stage + b * ld_stage + c0);
  const uint4* src = reinterpret_cast<const uint4*>(stg);
  const int nvec = (int)((total + 7) >> 3);
  for (int i = tid; i < nvec; i += THREADS) dst[i] = src[i];
}

// grid = B * tasks_per_b, task = R consecutive segments of one trajectory.
__global__ void __launch_bounds__(K0B_THREADS)
k0b_build_kernel(const uint16_t* __restrict__ stage, int64_t ld_stage, const uint16_t* __restrict__ hdr,
                 int64_t nchunks, int C, const int32_t* __restrict__ row_len, int64_t T, int64_t E,
                 uint32_t* __restrict__ mask, int64_t W, unsigned long long* __restrict__ nnz,
                 int tasks_per_b, int chunk_ids) {
  extern __shared__ __align__(16) uint32_t k0b_sm[];
  __shared__ long long part[K0B_WARPS];
  const int64_t b = blockIdx.x / tasks_per_b;
  const int task = (int)(blockIdx.x % tasks_per_b);
  const int seg0 = task * K0B_R;
  const int nseg = min(K0B_R, C - seg0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t w0 = (int64_t)seg0 * K0B_SEG_WORDS;
  const int nw = (int)min((int64_t)nseg * K0B_SEG_WORDS, W - w0);   // words of this slice
  const uint32_t tail = (E & 31) ? ((1u << (E & 31)) - 1u) : 0xffffffffu;
  for (int w = tid; w < nw; w += K0B_THREADS) k0b_sm[w] = (w0 + w == W - 1) ? tail : 0xffffffffu;
  __syncthreads();
  int64_t len = T;
  if (row_len) len = min(T, (int64_t)row_len[b]);
  const int64_t nch = (len + chunk_ids - 1) / chunk_ids;
  const uint16_t* srow = stage + b * ld_stage;
  const uint16_t* hrow = hdr + b * nchunks * (int64_t)(C + 1) + seg0;
  // A warp takes a chunk; its 8 groups of 4 lanes walk the chunk's R = 8 segment runs side by side
  // (group g = segment seg0 + g: the segment's base word is loop-invariant, no search for the
  // segment of an id, short runs cost no extra loop overhead). The next chunk's offsets are
  // fetched while this chunk's runs are walked.
  static_assert(K0B_R == 8, "lane groups assume 8 segments per build CTA");
  const int grp = lane >> 2, sub = lane & 3;
  uint32_t* seg = k0b_sm + grp * K0B_SEG_WORDS;
  int o_next = 0;
  if (warp < nch) o_next = (lane <= nseg) ? (int)__ldg(hrow + (int64_t)warp * (C + 1) + lane) : 0;
  for (int64_t c = warp; c < nch; c += K0B_WARPS) {
    const int o = o_next;
    if (c + K0B_WARPS < nch) o_next = (lane <= nseg) ? (int)__ldg(hrow + (c + K0B_WARPS) * (C + 1) + lane) : 0;
    const uint16_t* src = srow + c * chunk_ids;
    const int start = __shfl_sync(0xffffffffu, o, grp);
    int end = __shfl_sync(0xffffffffu, o, grp + 1);
    if (grp >= nseg) end = start;
    for (int j0 = start + sub; j0 < end; j0 += 16) {                // 4 loads in flight per lane
      uint32_t l[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) l[u] = (j0 + 4 * u < end) ? (uint32_t)__ldg(src + j0 + 4 * u) : 0xffffffffu;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (l[u] != 0xffffffffu) atomicAnd(seg + (l[u] >> 5), ~(1u << (l[u] & 31)));
    }
  }
  __syncthreads();
  long long cntv = 0;
  uint32_t* out = mask + b * W + w0;
  for (int w = tid; w < nw; w += K0B_THREADS) {
    const uint32_t x = k0b_sm[w];
    out[w] = x;
    cntv += __popc(x);
  }
  for (int o = 16; o; o >>= 1) cntv += __shfl_xor_sync(0xffffffffu, cntv, o);
  if (lane == 0) part[warp] = cntv;
  __syncthreads();
  if (tid == 0 && nnz) {
    long long t = 0;
    for (int i = 0; i < K0B_WARPS; ++i) t += part[i];
    atomicAdd(nnz + b, (unsigned long long)t);
  }
}

}  // namespace spai
