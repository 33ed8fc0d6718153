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
constexpr int K0B_CHUNK = 8192;                         // ids per sort CTA
constexpr int K0B_THREADS = 512;
constexpr int K0B_WARPS = K0B_THREADS / 32;
constexpr int K0B_IDS = K0B_CHUNK / K0B_THREADS;        // 16 ids per thread
constexpr int K0B_MAX_SEGS = 512;                       // E <= 33.5 M slots
constexpr int K0B_R = 8;                                // segments per build CTA (64 KB of shared memory)

inline size_t k0b_sort_smem(int C) { return (size_t)K0B_WARPS * C * 4 + (size_t)K0B_CHUNK * 2 + 64; }

// hdr u16[B][nchunks][C + 1]: hdr[..][s] = first position of segment s inside the sorted
// chunk, hdr[..][C] = number of valid ids of the chunk. Chunks at or beyond the row's length
// are not written (pass 2 derives the same chunk count from the same length).
template <typename IdT>
__global__ void __launch_bounds__(K0B_THREADS, 2)
k0b_sort_kernel(const IdT* __restrict__ actions, int64_t T, int64_t ld,
                const int32_t* __restrict__ row_len, const int32_t* __restrict__ edge_slot,
                int64_t E, int C, uint16_t* __restrict__ stage, int64_t ld_stage,
                uint16_t* __restrict__ hdr, int64_t nchunks) {
  extern __shared__ __align__(16) uint32_t k0b_sm[];
  uint32_t* cnt = k0b_sm;                                         // [WARPS][C] counters, then bases
  uint16_t* stg = reinterpret_cast<uint16_t*>(k0b_sm + ((K0B_WARPS * C + 3) & ~3));
  __shared__ uint32_t wtot[K0B_WARPS];
  const int64_t b = blockIdx.x / nchunks;
  const int64_t chunk = blockIdx.x % nchunks;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t len = T;
  if (row_len) len = min(T, (int64_t)row_len[b]);
  const int64_t c0 = chunk * K0B_CHUNK;
  if (c0 >= len) return;
  const IdT* row = actions + b * ld + c0;
  const int64_t left = len - c0;                                  // ids of this chunk (<= CHUNK used)

  IdT v[K0B_IDS];
#pragma unroll
  for (int u = 0; u < K0B_IDS; ++u) {
    const int t = u * K0B_THREADS + tid;
    v[u] = (t < left) ? __ldcs(row + t) : (IdT)-1;
  }
  for (int i = tid; i < K0B_WARPS * C; i += K0B_THREADS) cnt[i] = 0;
  __syncthreads();

  uint32_t key[K0B_IDS];       // slot, or 0xffffffff for ids that match no edge
  uint32_t pos[K0B_IDS];       // rank inside the (warp, segment) sub-list
  uint32_t* wc = cnt + warp * C;
#pragma unroll
  for (int u = 0; u < K0B_IDS; ++u) {
    const int64_t a = (int64_t)v[u];
    uint32_t s = 0xffffffffu;
    if ((uint64_t)a < (uint64_t)E) s = edge_slot ? (uint32_t)__ldg(edge_slot + a) : (uint32_t)a;
    key[u] = s;
    pos[u] = 0;
    if (s != 0xffffffffu) pos[u] = atomicAdd(wc + (s >> K0B_SEG_SHIFT), 1u);
  }
  __syncthreads();

  // exclusive scan in (segment, warp) order: thread s owns segment s (C <= THREADS)
  uint32_t tot = 0;
  uint32_t pre[K0B_WARPS];
  if (tid < C) {
#pragma unroll
    for (int w = 0; w < K0B_WARPS; ++w) { pre[w] = tot; tot += cnt[w * C + tid]; }
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
  for (int w = 0; w < K0B_WARPS; ++w) {
    const uint32_t x = wtot[w];
    if (w < warp) wbase += x;
    total += x;
  }
  const uint32_t segbase = wbase + inc - tot;
  uint16_t* h = hdr + (b * nchunks + chunk) * (int64_t)(C + 1);
  if (tid < C) {
#pragma unroll
    for (int w = 0; w < K0B_WARPS; ++w) cnt[w * C + tid] = segbase + pre[w];
    h[tid] = (uint16_t)segbase;
  }
  if (tid == 0) h[C] = (uint16_t)total;
  __syncthreads();

#pragma unroll
  for (int u = 0; u < K0B_IDS; ++u) {
    const uint32_t s = key[u];
    if (s != 0xffffffffu) stg[wc[s >> K0B_SEG_SHIFT] + pos[u]] = (uint16_t)(s & 0xffffu);
  }
  __syncthreads();

  // sorted chunk -> stage[b][c0 ..): 16-byte stores (ld_stage is a multiple of 8, c0 of 8192)
  uint4* dst = reinterpret_cast<uint4*>(stage + b * ld_stage + c0);
  const uint4* src = reinterpret_cast<const uint4*>(stg);
  const int nvec = (int)((total + 7) >> 3);
  for (int i = tid; i < nvec; i += K0B_THREADS) dst[i] = src[i];
}

// grid = B * tasks_per_b, task = R consecutive segments of one trajectory.
__global__ void __launch_bounds__(K0B_THREADS)
k0b_build_kernel(const uint16_t* __restrict__ stage, int64_t ld_stage, const uint16_t* __restrict__ hdr,
                 int64_t nchunks, int C, const int32_t* __restrict__ row_len, int64_t T, int64_t E,
                 uint32_t* __restrict__ mask, int64_t W, unsigned long long* __restrict__ nnz,
                 int tasks_per_b) {
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
  const int64_t nch = (len + K0B_CHUNK - 1) / K0B_CHUNK;
  const uint16_t* srow = stage + b * ld_stage;
  for (int64_t c = warp; c < nch; c += K0B_WARPS) {
    const uint16_t* h = hdr + (b * nchunks + c) * (int64_t)(C + 1) + seg0;
    const int o = (lane <= nseg) ? (int)__ldg(h + lane) : 0x7fffffff;
    int off[K0B_R + 1];
#pragma unroll
    for (int q = 0; q <= K0B_R; ++q) off[q] = __shfl_sync(0xffffffffu, o, q);   // off[q > nseg] = INT_MAX
    const int start = off[0];
    const int end = __shfl_sync(0xffffffffu, o, nseg);
    const uint16_t* src = srow + c * K0B_CHUNK;
    for (int j0 = start + lane; j0 < end; j0 += 128) {              // 4 loads in flight per lane
      uint32_t l[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) l[u] = (j0 + 32 * u < end) ? (uint32_t)__ldcs(src + j0 + 32 * u) : 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 32 * u;
        if (j < end) {
          int r = 0;
#pragma unroll
          for (int q = 1; q < K0B_R; ++q) r += (j >= off[q]) ? 1 : 0;
          atomicAnd(&k0b_sm[r * K0B_SEG_WORDS + (l[u] >> 5)], ~(1u << (l[u] & 31)));
        }
      }
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
