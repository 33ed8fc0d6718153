// K0 — action lists -> kept-edge bitmasks.
//
// Replaces gflownet/utils.py:315-323 (`set(actions)` + O(E) list comprehension
// per trajectory) and the two coalesce() sorts of utils.py:348-353, :124: the
// pattern's slots are sorted once at context creation, so "building M" for a
// trajectory is clearing one bit per action.
//
// Layouts
//   mask  u32[B][W]   slot-order bit p of trajectory b (row-major per trajectory:
//                     all atomics of a trajectory land in one W*4-byte region)
//   maskT u32[W][Bp]  the same bits transposed: the reward kernels put 32
//                     trajectories on the 32 lanes of a warp, so word w of 32
//                     consecutive trajectories is one coalesced 128-byte load.
#pragma once

#include "spai_internal.cuh"

namespace spai {

__global__ void k0_mask_init_kernel(uint32_t* __restrict__ mask, int64_t W, int64_t E, int64_t B) {
  const int64_t total = W * B;
  const uint32_t tail = (E & 31) ? ((1u << (E & 31)) - 1u) : 0xffffffffu;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t w = idx % W;
    mask[idx] = (w == W - 1) ? tail : 0xffffffffu;
  }
}

// grid: x = action chunks of one trajectory, y-major flattened into x as
// blockIdx.x = b * chunks + chunk so concurrently resident blocks work on
// neighbouring trajectories. Ids outside [0, E) (the -1 padding, the terminal
// id, anything else) match no edge, exactly like `i not in actions_set`.
template <int UNROLL, typename IdT = int64_t>
__global__ void __launch_bounds__(256)
k0_mask_clear_kernel(const IdT* __restrict__ actions, int64_t B, int64_t T, int64_t ld,
                     const int32_t* __restrict__ edge_slot, int64_t E,
                     uint32_t* __restrict__ mask, int64_t W, int64_t chunks,
                     const int32_t* __restrict__ row_len) {
  const int64_t b = blockIdx.x / chunks;
  const int64_t chunk = blockIdx.x % chunks;
  if (b >= B) return;
  if (row_len) T = min(T, (int64_t)row_len[b]);
  const IdT* row = actions + b * ld;
  uint32_t* mrow = mask + b * W;
  const int64_t t0 = chunk * (256 * UNROLL) + threadIdx.x;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const int64_t t = t0 + (int64_t)u * 256;
    if (t < T) {
      const int64_t a = (int64_t)__ldcs(row + t);          // streamed once
      if ((uint64_t)a < (uint64_t)E) {
        const int s = edge_slot ? __ldg(edge_slot + a) : (int)a;
        atomicAnd(mrow + (s >> 5), ~(1u << (s & 31)));
      }
    }
  }
}

// Shared-memory variant of init + clear + popcount for patterns whose bitmask fits
// in one CTA's shared memory (W*4 bytes <= K0S_MAX_SMEM): one block per trajectory
// streams the trajectory's action row once (16-byte loads), clears bits with
// shared-memory atomics (no L2 atomic traffic: ncu showed the global-RED variant
// L2-bound at 74 % lts throughput, one 32-byte L2 transaction per deletion),
// then writes the finished mask row with coalesced stores and its popcount.
constexpr int K0S_THREADS = 512;
constexpr int K0S_MAX_SMEM = 100 * 1024;

template <typename IdT> struct IdVec;
template <> struct IdVec<int64_t> {
  using type = longlong2;
  static constexpr int N = 2;
  template <typename F> static __device__ __forceinline__ void each(const longlong2& v, F&& f) { f(v.x); f(v.y); }
};
template <> struct IdVec<int32_t> {
  using type = int4;
  static constexpr int N = 4;
  template <typename F> static __device__ __forceinline__ void each(const int4& v, F&& f) {
    f((int64_t)v.x); f((int64_t)v.y); f((int64_t)v.z); f((int64_t)v.w);
  }
};

template <typename IdT>
__global__ void __launch_bounds__(K0S_THREADS)
k0_mask_build_smem_kernel(const IdT* __restrict__ actions, int64_t B, int64_t T, int64_t ld,
                          const int32_t* __restrict__ edge_slot, int64_t E,
                          uint32_t* __restrict__ mask, int64_t W, long long* __restrict__ nnz,
                          const int32_t* __restrict__ row_len) {
  extern __shared__ uint32_t k0_sm[];
  __shared__ long long part[K0S_THREADS / 32];
  using V = IdVec<IdT>;
  const int64_t b = blockIdx.x;
  if (b >= B) return;
  const int tid = threadIdx.x;
  const uint32_t tail = (E & 31) ? ((1u << (E & 31)) - 1u) : 0xffffffffu;
  for (int64_t w = tid; w < W; w += K0S_THREADS) k0_sm[w] = (w == W - 1) ? tail : 0xffffffffu;
  __syncthreads();
  const IdT* row = actions + b * ld;
  // caller-supplied / host-trimmed lengths: entries at positions >= row_len[b] are padding
  // the caller vouches for and are never read (4.35 of 8.59 GB on the cfg2 batch)
  if (row_len) T = min(T, (int64_t)row_len[b]);
  auto clear = [&](int64_t a) {
    if ((uint64_t)a < (uint64_t)E) {
      const int s = edge_slot ? __ldg(edge_slot + a) : (int)a;
      atomicAnd(&k0_sm[s >> 5], ~(1u << (s & 31)));
    }
  };
  {
    // rows are element-aligned only; peel until the next 16-byte granule
    int64_t head = (int64_t)((16 - (reinterpret_cast<uintptr_t>(row) & 15)) & 15) / (int64_t)sizeof(IdT);
    if (head > T) head = T;
    if (tid < head) clear((int64_t)row[tid]);
    const int64_t Tv = T - head;
    const typename V::type* rowv = reinterpret_cast<const typename V::type*>(row + head);
    const int64_t Tn = Tv / V::N;
    constexpr int LD = 8;                 // independent 16-byte loads in flight per thread
    int64_t t = tid;
    for (; t + (int64_t)(LD - 1) * K0S_THREADS < Tn; t += (int64_t)LD * K0S_THREADS) {
      typename V::type v[LD];
#pragma unroll
      for (int u = 0; u < LD; ++u) v[u] = __ldcs(rowv + t + (int64_t)u * K0S_THREADS);
#pragma unroll
      for (int u = 0; u < LD; ++u) V::each(v[u], clear);
    }
    for (; t < Tn; t += K0S_THREADS) {
      const typename V::type v = __ldcs(rowv + t);
      V::each(v, clear);
    }
    const int64_t done = head + Tn * V::N;
    if (tid < T - done) clear((int64_t)row[done + tid]);
  }
  __syncthreads();
  long long cnt = 0;
  uint32_t* out = mask + b * W;
  for (int64_t w = tid; w < W; w += K0S_THREADS) {
    const uint32_t v = k0_sm[w];
    out[w] = v;
    cnt += __popc(v);
  }
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((tid & 31) == 0) part[tid >> 5] = cnt;
  __syncthreads();
  if (tid == 0 && nnz) {
    long long t = 0;
    for (int i = 0; i < K0S_THREADS / 32; ++i) t += part[i];
    nnz[b] = t;
  }
}

// taken u32[B][ld] in EDGE order (bit set = edge removed) -> mask u32[B][W] in
// slot order (bit set = kept).
__global__ void k0_mask_from_taken_kernel(const uint32_t* __restrict__ taken, int64_t ld,
                                          const int32_t* __restrict__ slot_edge, int64_t E,
                                          uint32_t* __restrict__ mask, int64_t W, int64_t B) {
  const int64_t total = W * B;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / W, w = idx % W;
    const uint32_t tail = (w == W - 1 && (E & 31)) ? ((1u << (E & 31)) - 1u) : 0xffffffffu;
    uint32_t out;
    if (!slot_edge) {
      out = ~taken[b * ld + w] & tail;
    } else {
      out = 0;
      const int64_t p0 = w * 32;
      for (int j = 0; j < 32 && p0 + j < E; ++j) {
        const int e = __ldg(slot_edge + p0 + j);
        const uint32_t t = taken[b * ld + (e >> 5)];
        out |= (((t >> (e & 31)) & 1u) ^ 1u) << j;
      }
    }
    mask[idx] = out;
  }
}

// mask u32[B][W] -> maskT u32[W][Bp]; the padding columns b in [B, Bp) are written as
// all-ones ("nothing removed") so they never defeat the untouched-row vote of K3;
// their results are discarded.
__global__ void __launch_bounds__(256)
k0_transpose_kernel(const uint32_t* __restrict__ mask, int64_t B, int64_t W,
                    uint32_t* __restrict__ maskT, int64_t Bp) {
  __shared__ uint32_t tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int64_t w0 = (int64_t)blockIdx.x * 32, b0 = (int64_t)blockIdx.y * 32;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t b = b0 + r, w = w0 + tx;
    tile[r][tx] = (w < W) ? ((b < B) ? mask[b * W + w] : 0xffffffffu) : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t w = w0 + r, b = b0 + tx;
    if (w < W && b < Bp) maskT[w * Bp + b] = tile[tx][r];
  }
}

// The same transpose on 64 x 64 tiles with 16-byte accesses on both sides (W % 4 == 0): four independent 16-byte loads in
// flight per thread instead of four 4-byte ones, 256-byte segments per row on both sides.
__global__ void __launch_bounds__(256)
k0_transpose64_kernel(const uint32_t* __restrict__ mask, int64_t B, int64_t W,
                      uint32_t* __restrict__ maskT, int64_t Bp) {
  __shared__ uint32_t tile[64][65];
  const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;        // 16 x 16 threads, 4 columns each
  const int64_t w0 = (int64_t)blockIdx.x * 64, b0 = (int64_t)blockIdx.y * 64;
  uint4 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t b = b0 + r0 + 16 * i, w = w0 + c4;
    if (w + 3 < W && b < B) v[i] = *reinterpret_cast<const uint4*>(mask + b * W + w);
    else {
      uint32_t t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = (w + j < W) ? ((b < B) ? mask[b * W + w + j] : 0xffffffffu) : 0u;
      v[i] = make_uint4(t[0], t[1], t[2], t[3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t* dst = &tile[r0 + 16 * i][c4];
    dst[0] = v[i].x; dst[1] = v[i].y; dst[2] = v[i].z; dst[3] = v[i].w;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int wr = r0 + 16 * i;
    const int64_t w = w0 + wr, b = b0 + c4;
    if (w < W && b + 3 < Bp)
      *reinterpret_cast<uint4*>(maskT + w * Bp + b) = make_uint4(tile[c4][wr], tile[c4 + 1][wr], tile[c4 + 2][wr], tile[c4 + 3][wr]);
  }
}

// nnz(M) per trajectory = kept slots, minus the surplus of repeated coordinates
// (coalesce() merges them, preconditioner.py:71 counts stored entries).
__global__ void __launch_bounds__(256)
k0_popcount_kernel(const uint32_t* __restrict__ mask, int64_t W, int64_t B,
                   long long* __restrict__ nnz) {
  const int64_t b = blockIdx.x;
  if (b >= B) return;
  long long s = 0;
  for (int64_t w = threadIdx.x; w < W; w += blockDim.x) s += __popc(mask[b * W + w]);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ long long part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    nnz[b] = t;
  }
}

__global__ void k0_dup_correction_kernel(const uint32_t* __restrict__ mask, int64_t W, int64_t B,
                                         const int32_t* __restrict__ dup_start,
                                         const int32_t* __restrict__ dup_len, int64_t ndup,
                                         long long* __restrict__ nnz) {
  const int64_t total = ndup * B;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / ndup, g = idx % ndup;
    int kept = 0;
    const int s0 = dup_start[g];
    for (int j = 0; j < dup_len[g]; ++j) {
      const int s = s0 + j;
      kept += (mask[b * W + (s >> 5)] >> (s & 31)) & 1u;
    }
    if (kept > 1) atomicAdd((unsigned long long*)(nnz + b), (unsigned long long)(-(long long)(kept - 1)));
  }
}

// slot-order mask -> EDGE-order bytes (the reference's `remaining_edges_mask`
// as a 0/1 vector, gflownet/utils.py:323).
__global__ void k0_kept_bytes_kernel(const uint32_t* __restrict__ mask, int64_t W, int64_t B,
                                     const int32_t* __restrict__ edge_slot, int64_t E,
                                     uint8_t* __restrict__ out) {
  const int64_t total = E * B;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / E, e = idx % E;
    const int s = edge_slot ? edge_slot[e] : (int)e;
    out[idx] = (uint8_t)((mask[b * W + (s >> 5)] >> (s & 31)) & 1u);
  }
}

}  // namespace spai
