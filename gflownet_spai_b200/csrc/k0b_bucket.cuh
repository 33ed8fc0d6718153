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

inline size_t k0b_sort_smem(int C, int threads, int ids = K0B_IDS) {
  return (size_t)(threads / 32) * (C + 1) * 4 + (size_t)threads * ids * 2 + 128;
}

// hdr u16[B][nchunks][C + 1]: hdr[..][s] = first position of segment s inside the sorted
// chunk, hdr[..][C] = number of valid ids of the chunk. Chunks at or beyond the row's length
// are not written (pass 2 derives the same chunk count from the same length).
// Straight-line code: every id takes the same path (ids that match no edge are counted in a
// trash bucket C of the warp and never stored), the edge -> slot map is a template switch.
template <typename IdT, bool HAS_MAP, int THREADS, int IDS = K0B_IDS>
__global__ void __launch_bounds__(THREADS, (IDS == K0B_IDS ? 1024 : 1536) / THREADS)
k0b_sort_kernel(const IdT* __restrict__ actions, int64_t T, int64_t ld,
                const int32_t* __restrict__ row_len, const int32_t* __restrict__ edge_slot,
                int64_t E, int C, uint16_t* __restrict__ stage, int64_t ld_stage,
                uint16_t* __restrict__ hdr, int64_t nchunks) {
  constexpr int NWARPS = THREADS / 32, CHUNK = THREADS * IDS;
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

  IdT v[IDS];
  if (left == CHUNK) {
#pragma unroll
    for (int u = 0; u < IDS; ++u) v[u] = __ldcs(row + u * THREADS + tid);
  } else {
#pragma unroll
    for (int u = 0; u < IDS; ++u) {
      const int t = u * THREADS + tid;
      v[u] = (t < left) ? __ldcs(row + t) : (IdT)-1;
    }
  }
  for (int i = tid; i < NWARPS * C1; i += THREADS) cnt[i] = 0;
  __syncthreads();

  uint32_t key[IDS];       // slot; 0xffffffff (segment = trash) for ids that match no edge
  uint32_t pos[IDS];       // rank inside the (warp, segment) sub-list
  uint32_t* wc = cnt + warp * C1;
  const uint32_t Eu = (uint32_t)E;                                // E < 2^31 (checked at context creation)
#pragma unroll
  for (int u = 0; u < IDS; ++u) {
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
    key[u] = ok ? s : 0xffffffffu;                                  // `ok` dies here: 16 live predicates cost a packed register
    const uint32_t seg = min(key[u] >> K0B_SEG_SHIFT, (uint32_t)C);
    pos[u] = atomicAdd(wc + seg, 1u);
  }
  __syncthreads();

  // exclusive scan in (segment, warp) order: thread s owns segment s (C <= THREADS); the trash
  // bucket is left out. Only the warps that own segments take part (C = 64: two of eight).
  const int nwc = (C + 31) >> 5;
  __shared__ uint32_t s_total;
  uint32_t tot = 0, inc = 0;
  if (warp < nwc) {
    if (tid < C) {
#pragma unroll
      for (int w = 0; w < NWARPS; ++w) tot += cnt[w * C1 + tid];
    }
    inc = tot;                                                    // inclusive scan over segments
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += x;
    }
    if (lane == 31) wtot[warp] = inc;
  }
  __syncthreads();
  if (warp < nwc) {
    uint32_t wbase = 0, tsum = 0;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) {
      const uint32_t x = (w < nwc) ? wtot[w] : 0u;
      if (w < warp) wbase += x;
      tsum += x;
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
    if (tid == 0) { h[C] = (uint16_t)tsum; s_total = tsum; }
  }
  if (tid < NWARPS) cnt[tid * C1 + C] = (uint32_t)CHUNK;    // trash ids land behind the chunk (never copied out)
  __syncthreads();

#pragma unroll
  for (int u = 0; u < IDS; ++u) {
    const uint32_t s = key[u];
    const uint32_t seg = min(s >> K0B_SEG_SHIFT, (uint32_t)C);
    const uint32_t at = wc[seg] + ((s == 0xffffffffu) ? 0u : pos[u]);
    stg[at] = (uint16_t)(s & 0xffffu);
  }
  __syncthreads();

  // sorted chunk -> stage[b][c0 ..): 16-byte stores (ld_stage is a multiple of 8, c0 of 8192)
  uint4* dst = reinterpret_cast<uint4*>(stage + b * ld_stage + c0);
  const uint4* src = reinterpret_cast<const uint4*>(stg);
  const int nvec = (int)((s_total + 7) >> 3);
  for (int i = tid; i < nvec; i += THREADS) dst[i] = src[i];
}

// Sort pass, second version (opt-in, SPAI_K0B_SORT=2; measured slower than version 1): PERSISTENT CTAs (one wave) walk the (trajectory, chunk) items grid-stride; the ids of the
// CTA's NEXT item are already on their way into shared memory (cp.async, one element per copy: rows are element-aligned
// only) while the current chunk is ranked, scattered and written out, so the HBM stream never pauses for a CTA's compute
// phases and no CTA starts with an empty pipeline. The chunk body is the one of k0b_sort_kernel (same staged layout).
inline size_t k0b_sort2_smem(int C, int threads, int elem) {
  return k0b_sort_smem(C, threads) + (size_t)threads * K0B_IDS * elem;
}
template <int BYTES>
__device__ __forceinline__ void k0b_cp_async(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}
template <typename IdT, bool HAS_MAP, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k0b_sort2_kernel(const IdT* __restrict__ actions, int64_t T, int64_t ld,
                 const int32_t* __restrict__ row_len, const int32_t* __restrict__ edge_slot,
                 int64_t E, int C, uint16_t* __restrict__ stage, int64_t ld_stage,
                 uint16_t* __restrict__ hdr, int64_t nchunks, int64_t total) {
  constexpr int NWARPS = THREADS / 32, CHUNK = THREADS * K0B_IDS;
  extern __shared__ __align__(16) uint32_t k0b_sm[];
  const int C1 = C + 1;                                           // + trash bucket
  uint32_t* cnt = k0b_sm;                                         // [WARPS][C1] counters, then bases
  uint16_t* stg = reinterpret_cast<uint16_t*>(k0b_sm + ((NWARPS * C1 + 3) & ~3));
  IdT* buf = reinterpret_cast<IdT*>(stg + CHUNK);                 // the next item's ids (CHUNK * 2 bytes of stg keep it 16-byte aligned)
  __shared__ uint32_t wtot[NWARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t Eu = (uint32_t)E;                                // E < 2^31 (checked at context creation)
  const uint32_t buf_addr = (uint32_t)__cvta_generic_to_shared(buf + tid);

  // item w = (trajectory b, chunk): valid when the chunk starts inside the row
  int64_t b = 0, chunk = 0;
  int left = 0;
  auto locate = [&](int64_t w) {                                  // first valid item at or after w (grid-stride), or total
    while (w < total) {
      b = w / nchunks;
      chunk = w - b * nchunks;
      int64_t len = T;
      if (row_len) len = min(T, (int64_t)row_len[b]);
      const int64_t c0 = chunk * CHUNK;
      if (c0 < len) { left = (int)min(len - c0, (int64_t)CHUNK); return w; }
      w += gridDim.x;
    }
    return w;
  };
  auto prefetch = [&]() {                                         // ids of item (b, chunk) -> buf
    const IdT* row = actions + b * ld + chunk * CHUNK + tid;
#pragma unroll
    for (int u = 0; u < K0B_IDS; ++u)
      if (u * THREADS + tid < left) k0b_cp_async<(int)sizeof(IdT)>(buf_addr + u * THREADS * (uint32_t)sizeof(IdT), row + u * THREADS);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int64_t w = locate(blockIdx.x);
  if (w < total) prefetch();
  while (w < total) {
    const int64_t cb = b, cchunk = chunk;
    const int cleft = left;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    IdT v[K0B_IDS];
#pragma unroll
    for (int u = 0; u < K0B_IDS; ++u) v[u] = (u * THREADS + tid < cleft) ? buf[u * THREADS + tid] : (IdT)-1;
    for (int i = tid; i < NWARPS * C1; i += THREADS) cnt[i] = 0;
    __syncthreads();                                              // buf is in registers, counters are zero
    const int64_t wn = locate(w + gridDim.x);
    if (wn < total) prefetch();                                   // flies during everything below

    uint32_t key[K0B_IDS];       // slot; 0xffffffff (segment = trash) for ids that match no edge
    uint32_t pos[K0B_IDS];       // rank inside the (warp, segment) sub-list
    uint32_t* wc = cnt + warp * C1;
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
      key[u] = ok ? s : 0xffffffffu;                                // `ok` dies here: 16 live predicates cost a packed register
      const uint32_t seg = min(key[u] >> K0B_SEG_SHIFT, (uint32_t)C);
      pos[u] = atomicAdd(wc + seg, 1u);
    }
    __syncthreads();

    uint32_t tot = 0;
    if (tid < C) {
#pragma unroll
      for (int ww = 0; ww < NWARPS; ++ww) tot += cnt[ww * C1 + tid];
    }
    uint32_t inc = tot;                                             // inclusive scan over segments
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += x;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, total_ids = 0;
#pragma unroll
    for (int ww = 0; ww < NWARPS; ++ww) {
      const uint32_t x = wtot[ww];
      if (ww < warp) wbase += x;
      total_ids += x;
    }
    const uint32_t segbase = wbase + inc - tot;
    uint16_t* h = hdr + (cb * nchunks + cchunk) * (int64_t)C1;
    if (tid < C) {
      uint32_t run = segbase;
#pragma unroll
      for (int ww = 0; ww < NWARPS; ++ww) {
        const uint32_t x = cnt[ww * C1 + tid];
        cnt[ww * C1 + tid] = run;
        run += x;
      }
      h[tid] = (uint16_t)segbase;
    }
    if (tid == 0) h[C] = (uint16_t)total_ids;
    if (tid < NWARPS) cnt[tid * C1 + C] = (uint32_t)CHUNK;    // trash ids land behind the chunk (never copied out)
    __syncthreads();

#pragma unroll
    for (int u = 0; u < K0B_IDS; ++u) {
      const uint32_t s = key[u];
      const uint32_t seg = min(s >> K0B_SEG_SHIFT, (uint32_t)C);
      const uint32_t at = wc[seg] + ((s == 0xffffffffu) ? 0u : pos[u]);
      if (at < (uint32_t)CHUNK) stg[at] = (uint16_t)(s & 0xffffu);          // trash ids: nothing behind stg but the next item's ids
    }
    __syncthreads();

    uint4* dst = reinterpret_cast<uint4*>(stage + cb * ld_stage + cchunk * CHUNK);
    const uint4* src = reinterpret_cast<const uint4*>(stg);
    const int nvec = (int)((total_ids + 7) >> 3);
    for (int i = tid; i < nvec; i += THREADS) dst[i] = src[i];
    w = wn;
  }
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

// Build pass, second version (ncu source view of the first, cfg3 B = 256: half of the samples wait on the four 2-byte
// loads a lane has in flight, and every shared-memory atomic drags five address instructions along). Same task
// decomposition; a lane group now reads its run as aligned 8-byte words (four ids per load, up to four loads = the whole
// run of a cfg3 chunk in flight), the loads of the warp's NEXT chunk are issued before the bits of the current one are
// cleared, and bits are cleared by red.shared on a 32-bit shared address.
__device__ __forceinline__ void k0b_red_and(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 k0b_ldg8(const uint2* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
// L = lanes per (chunk, segment) run: a warp-iteration covers 4 / L chunks (32 lanes = 8 segments x L lanes x 4 / L
// chunks), so the per-iteration bookkeeping is spread over 4 / L times the ids when runs are short (L = 4: runs of
// >= 48 ids, cfg3; L = 2: cfg4, cfg5).
template <int L>
__global__ void __launch_bounds__(K0B_THREADS)
k0b_build2_kernel(const uint16_t* __restrict__ stage, int64_t ld_stage, const uint16_t* __restrict__ hdr,
                  int64_t nchunks, int C, const int32_t* __restrict__ row_len, int64_t T, int64_t E,
                  uint32_t* __restrict__ mask, int64_t W, unsigned long long* __restrict__ nnz,
                  int tasks_per_b, int chunk_ids) {
  extern __shared__ __align__(16) uint32_t k0b_sm[];
  __shared__ long long part[K0B_WARPS];
  constexpr int CPI = 4 / L;                                         // chunks per warp-iteration
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
  const int nch = (int)((len + chunk_ids - 1) / chunk_ids);
  static_assert(K0B_R == 8, "lane groups assume 8 segments per build CTA");
  const int sub = lane % L, grp = (lane / L) & 7, cj = lane / (8 * L);
  const bool live = grp < nseg;                                      // segments beyond the pattern have no shared memory
  const uint32_t seg_addr = (uint32_t)__cvta_generic_to_shared(k0b_sm + grp * K0B_SEG_WORDS);
  const int C1 = C + 1, cw = chunk_ids >> 2;
  const uint16_t* hb = hdr + b * nchunks * (int64_t)C1 + seg0 + grp;                      // + chunk * C1
  const uint2* sb = reinterpret_cast<const uint2*>(stage + b * ld_stage);                  // + chunk * cw
  constexpr int CSTEP = K0B_WARPS * CPI;

  // clear the bits of one 8-byte word (ids at chunk positions 4q .. 4q+3; valid inside [start, end)): straight-line
  // code, the atomic of an id outside the run is predicated off
  auto clear4 = [&](const uint2& v, int q, int start, int end) {
    const uint32_t span = (uint32_t)(end - start);
    const uint32_t t = (uint32_t)(4 * q - start);                   // sub-position i is valid iff t + i < span (unsigned)
    const uint32_t id[4] = {v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %0, %1;\n\t@p red.shared.and.b32 [%2], %3;\n\t}" ::"r"(t + (uint32_t)i),
                   "r"(span), "r"(seg_addr + ((id[i] >> 3) & 0x1ffcu)), "r"(~(1u << (id[i] & 31u)))
                   : "memory");
  };
  auto header = [&](int c, int& start, int& end) {                  // run bounds of (chunk c, segment seg0 + grp)
    start = 0; end = 0;
    if (L == 4) {                                                   // one chunk per warp: lanes 0..8 fetch its nine entries once
      int o = 0;
      if (c < nch && lane <= nseg) o = (int)__ldg(hb - grp + (int64_t)c * C1 + lane);
      start = __shfl_sync(0xffffffffu, o, grp);
      end = __shfl_sync(0xffffffffu, o, grp + 1);
      if (!live) end = start;
    } else if (live && c < nch) {                                   // the lane's own two adjacent 16-bit entries
      start = (int)__ldg(hb + (int64_t)c * C1);
      end = (int)__ldg(hb + (int64_t)c * C1 + 1);
    }
  };
  auto fetch = [&](int c, int start, int end, uint2 (&v)[4]) {
    const uint2* src = sb + (int64_t)c * cw;
    const int q = (start >> 2) + sub, q1 = (end + 3) >> 2;
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (q + L * u < q1) ? k0b_ldg8(src + q + L * u) : make_uint2(0u, 0u);
  };

  int c = warp * CPI + cj;
  int start, end, nstart, nend;
  uint2 cur[4];
  header(c, start, end);
  header(c + CSTEP, nstart, nend);
  fetch(c, start, end, cur);
  for (int cb = warp * CPI; cb < nch; cb += CSTEP) {                 // warp-uniform
    // the next iteration's words fly while this one's bits are cleared; the header after that one is fetched too
    uint2 nxt[4];
    fetch(c + CSTEP, nstart, nend, nxt);
    int n2s, n2e;
    header(c + 2 * CSTEP, n2s, n2e);

    const uint2* src = sb + (int64_t)c * cw;
    const int q = (start >> 2) + sub, q1 = (end + 3) >> 2;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (q + L * u < q1) clear4(cur[u], q + L * u, start, end);
    for (int qq = q + 4 * L; qq < q1; qq += L) clear4(k0b_ldg8(src + qq), qq, start, end);   // longer runs: not pipelined

#pragma unroll
    for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
    start = nstart; end = nend;
    nstart = n2s; nend = n2e;
    c += CSTEP;
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

// Build pass, third version: a LANE owns the run of one (chunk, segment) pair. Warp w of the CTA takes segment w & 7 of
// the task and 32 chunks at a time (warps 0-7 the even blocks of 32 chunks, warps 8-15 the odd ones); every lane reads
// its run — contiguous in its chunk — as aligned 16-byte words (8 ids), up to four in flight, while the next block's two
// header entries are already being fetched. No lane waits on a short run of its neighbours' chunk and the per-chunk work
// of versions 1 and 2 (nine header entries, shuffles and bound checks per 4096 / C * 8 ids) is gone: what remains per run
// is two header loads and the word loop. Short runs (cfg4: 32 ids, cfg5: 24 ids per chunk and segment) gain most.
__device__ __forceinline__ uint4 k0b_ldg16(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__global__ void __launch_bounds__(K0B_THREADS)
k0b_build3_kernel(const uint16_t* __restrict__ stage, int64_t ld_stage, const uint16_t* __restrict__ hdr,
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
  const int nch = (int)((len + chunk_ids - 1) / chunk_ids);
  static_assert(K0B_R == 8 && K0B_WARPS == 16, "warp -> (segment, chunk block parity) map");
  const int g = warp & 7, half = warp >> 3;
  if (g < nseg) {
    const uint32_t seg_addr = (uint32_t)__cvta_generic_to_shared(k0b_sm + g * K0B_SEG_WORDS);
    const int C1 = C + 1, cw = chunk_ids >> 3;
    const uint16_t* hb = hdr + b * nchunks * (int64_t)C1 + seg0 + g;
    const uint4* sb = reinterpret_cast<const uint4*>(stage + b * ld_stage);
    auto clear8 = [&](const uint4& v, int q, int start, int end) {
      const int lo = start - 8 * q, hi = end - 8 * q;                // valid sub-positions: lo <= i < hi
      const uint32_t id[8] = {v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16, v.z & 0xffffu, v.z >> 16, v.w & 0xffffu, v.w >> 16};
      if (lo <= 0 && hi >= 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) k0b_red_and(seg_addr + ((id[i] >> 3) & 0x1ffcu), ~(1u << (id[i] & 31u)));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i >= lo && i < hi) k0b_red_and(seg_addr + ((id[i] >> 3) & 0x1ffcu), ~(1u << (id[i] & 31u)));
      }
    };
    int c = half * 32 + lane;
    int o0 = 0, o1 = 0;
    if (c < nch) { o0 = (int)__ldg(hb + (int64_t)c * C1); o1 = (int)__ldg(hb + (int64_t)c * C1 + 1); }
    for (int cb = half * 32; cb < nch; cb += 64) {                   // warp-uniform block loop
      const int cn = c + 64;
      int n0 = 0, n1 = 0;
      if (cn < nch) { n0 = (int)__ldg(hb + (int64_t)cn * C1); n1 = (int)__ldg(hb + (int64_t)cn * C1 + 1); }
      if (c < nch) {
        const uint4* src = sb + (int64_t)c * cw;
        const int q1 = (o1 + 7) >> 3;
        for (int q = o0 >> 3; q < q1; q += 4) {
          uint4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = (q + u < q1) ? k0b_ldg16(src + q + u) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (q + u < q1) clear8(v[u], q + u, o0, o1);
        }
      }
      c = cn; o0 = n0; o1 = n1;
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
