// K3s — copy-mode reward driven by the DELETED slots instead of by rows x trajectories
// (SURVEY §8f-2 "incremental reward, touched rows only"), for trajectories that delete a small
// fraction of the candidates (early training, huge n).
//
// Row i of M*A - I has entries  f_x - sum_{deleted e} w_e(x)  where f_x is the entry with every
// candidate kept and w_e(x) = m_e * A[c_e, x] the contribution of slot e. Expanding the square,
//
//   r_i(kept) = base_i + sum_{deleted e} sum_x w_e(x) (w_e(x) - 2 f_x)
//                      + 2 sum_{deleted e' < e} sum_x w_e(x) w_e'(x)
//
// base_i = the row's residual^2 with every candidate kept (cached per context). So a deleted slot
// costs one pass over ITS OWN r_c (f, w) pairs (a slot-major copy of the plan, k3s_build_kernel),
// plus one sorted merge of two rows of A per pair of deleted slots in the same row (rare when
// deletions are sparse) — never the row's k * r_c records. One warp scans 32 words of a
// trajectory's kept-mask (trajectory-major, coalesced), compacts the zero bits (= deleted slots;
// duplicates and invalid action ids never show up here) into a shared-memory list with
// ballot-free prefix sums and handles 32 list entries per pass, one per lane. Work is
// proportional to the number of deletions, not to n * B. Blocks that handle the same word range
// of different trajectories are adjacent in the grid, so the pairs of the range stay in L2 while
// the batch sweeps it. Deltas are formed in the data's precision, the per-trajectory sum in fp64.
//
// Replaces the same reference lines as K3 (preconditioner.py:79-93 on the pattern of
// gflownet/utils.py:315-353); fp32 results differ from K3's by the rounding of the expanded form
// (~1e-7 relative), inside the 1e-4 bar.
#pragma once

#include "k3_copy.cuh"
#include "spai_internal.cuh"

namespace spai {

constexpr int K3S_THREADS = 128;
constexpr int K3S_LIST = 1024 + 32;      // zero bits of one 32-word batch + carry-over

// Everything a deleted slot needs, in one 16-byte load: its row's slot range (the pair check)
// and its own (f, w) list.
struct alignas(16) SlotMeta {
  int32_t sp;        // first slot of the slot's row
  uint16_t k;        // candidate slots of the row
  uint16_t rc;       // pairs of this slot = nnz(A[c_e, :])
  int64_t off;       // first pair in sl_rec
};

template <typename T> struct Pair;
template <> struct alignas(8) Pair<float> { float f, w; };
template <> struct alignas(16) Pair<double> { double f, w; };

// ---------------------------------------------------------------- build (once per context)
// One thread per row: walks the row's records segment by segment and stores (segment sum,
// contribution) at the record's place in its slot's list = the rank of the output column in
// A[c_e, :] (binary search), so list j of a slot lines up with entry j of that row of A.
template <typename T>
__global__ void k3s_build_kernel(const typename RecOf<T>::type* __restrict__ recs,
                                 const int64_t* __restrict__ cptr, const int32_t* __restrict__ c_col,
                                 const RowHdr* __restrict__ rhdr, const int32_t* __restrict__ slot_col,
                                 const int32_t* __restrict__ a_ptr, const int32_t* __restrict__ a_col,
                                 const SlotMeta* __restrict__ meta, int64_t n, Pair<T>* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RowHdr h = rhdr[i];
  const int64_t cb = cptr[i];
  const auto* rp = recs + cb;
  T acc = (h.flags & 1) ? T(-1) : T(0);
  int c0 = 0;
  for (int c = 0; c < h.cnt; ++c) {
    const auto r = rp[c];
    acc += rec_w(r);
    if (r.flags & F_END) {
      const int x = c_col[cb + c];                   // output column of the segment
      for (int q = c0; q <= c; ++q) {
        const auto rq = rp[q];
        const int slot = h.sp + (int)rec_e(rq.flags);
        const int cc = slot_col[slot];
        int lo = a_ptr[cc], hi = a_ptr[cc + 1];
        const int first = lo;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (a_col[mid] < x) lo = mid + 1; else hi = mid; }
        Pair<T> p;
        p.f = acc; p.w = rec_w(rq);
        out[meta[slot].off + (lo - first)] = p;
      }
      acc = (r.flags & F_NEXT_DIAG) ? T(-1) : T(0);
      c0 = c + 1;
    }
  }
}

// accumulation type of the deltas: fp32 data is combined in fp32 (2 FMAs per pair), the running
// sum of a trajectory is fp64 either way
template <typename T> __device__ __forceinline__ T k3s_single(const Pair<T>* __restrict__ pp, int rc) {
  T d = T(0);
  for (int q = 0; q < rc; ++q) {
    const Pair<T> p = pp[q];
    d = fma(p.w, fma(T(-2), p.f, p.w), d);
  }
  return d;
}

// sum over the common output columns of two deleted slots of w_1(x) * w_2(x): sorted merge of
// two rows of A
template <typename T>
__device__ __forceinline__ T k3s_cross(int s1, int s2, const int32_t* __restrict__ slot_col,
                                       const int32_t* __restrict__ a_ptr, const int32_t* __restrict__ a_col,
                                       const SlotMeta* __restrict__ meta, const Pair<T>* __restrict__ sl_rec) {
  const SlotMeta m1 = meta[s1], m2 = meta[s2];
  const int32_t* x1 = a_col + a_ptr[slot_col[s1]];
  const int32_t* x2 = a_col + a_ptr[slot_col[s2]];
  const int n1 = m1.rc, n2 = m2.rc;
  const Pair<T>* p1 = sl_rec + m1.off;
  const Pair<T>* p2 = sl_rec + m2.off;
  T acc = T(0);
  int i = 0, j = 0;
  while (i < n1 && j < n2) {
    const int a = x1[i], b = x2[j];
    if (a == b) acc = fma(p1[i].w, p2[j].w, acc);
    i += (a <= b);
    j += (b <= a);
  }
  return acc;
}

constexpr int K3S_XCAP = 256;            // deferred (slot, earlier deleted slot) pairs per warp

template <typename T>
__global__ void __launch_bounds__(K3S_THREADS)
k3s_sparse_kernel(const SlotMeta* __restrict__ meta, const int32_t* __restrict__ slot_col,
                  const int32_t* __restrict__ a_ptr, const int32_t* __restrict__ a_col,
                  const Pair<T>* __restrict__ sl_rec, const uint32_t* __restrict__ mask, int64_t W,
                  int64_t B, int64_t Bp, int64_t w_lo, int64_t w_hi, int64_t chunk_words,
                  int slot_lo, int slot_hi, double base_sum, double* __restrict__ partial,
                  long long* __restrict__ nnz) {
  __shared__ int32_t light[K3S_THREADS / 32][K3S_LIST];
  __shared__ int32_t xs1[K3S_THREADS / 32][K3S_XCAP], xs2[K3S_THREADS / 32][K3S_XCAP];
  __shared__ double wsum[K3S_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;                       // trajectory (x: neighbours share the word range)
  const int64_t c_lo = w_lo + (int64_t)blockIdx.y * chunk_words;
  const int64_t c_hi = (c_lo + chunk_words < w_hi) ? c_lo + chunk_words : w_hi;
  const uint32_t* mrow = mask + b * W;
  int32_t* L = light[warp];
  int32_t* X1 = xs1[warp];
  int32_t* X2 = xs2[warp];
  int nl = 0, nx = 0;                                 // warp-uniform list fills
  double tot = 0.0;
  int ones = 0;                                       // kept slots seen by this lane (nnz(M), when asked)

  // the pairs of deleted slots that share a row are rare and long (a merge of two rows of A):
  // they are queued and run 32 at a time so they never stall a pass of short single deltas
  auto cross_pass = [&](int first, int count) {
    if (lane < count)
      tot += 2.0 * (double)k3s_cross<T>(X1[first + lane], X2[first + lane], slot_col, a_ptr, a_col, meta, sl_rec);
  };
  auto cross_flush = [&](bool all) {
    int done = 0;
    while (nx - done >= 32) { cross_pass(done, 32); done += 32; }
    if (all && nx > done) { cross_pass(done, nx - done); done = nx; }
    const int rest = nx - done;
    const int m1 = (lane < rest) ? X1[done + lane] : 0;
    const int m2 = (lane < rest) ? X2[done + lane] : 0;
    __syncwarp();
    if (lane < rest) { X1[lane] = m1; X2[lane] = m2; }
    nx = rest;
    __syncwarp();
  };

  auto pass = [&](int first, int count) {             // slots L[first..first+count), one per lane
    uint32_t before = 0u;
    int s = 0, sp = 0;
    if (lane < count) {
      s = L[first + lane];
      if (s >= slot_lo && s < slot_hi) {              // row range; also drops the zero bits past E in the last word
        const SlotMeta h = meta[s];
        sp = h.sp;
        tot += (double)k3s_single<T>(sl_rec + h.off, (int)h.rc);
        if (h.k <= 32) {
          const int64_t w0 = h.sp >> 5;
          const uint32_t lo = mrow[w0];
          const uint32_t hi = (w0 + 1 < W) ? mrow[w0 + 1] : 0u;
          const uint32_t kmask = (h.k >= 32) ? 0xffffffffu : ((1u << h.k) - 1u);
          before = ~__funnelshift_r(lo, hi, h.sp & 31) & kmask & ((1u << (s - h.sp)) - 1u);
        } else {                                      // wide row: earlier deleted slots found bit by bit, done in place
          T x = T(0);
          for (int o = h.sp; o < s; ++o)
            if (!((mrow[o >> 5] >> (o & 31)) & 1u)) x += k3s_cross<T>(s, o, slot_col, a_ptr, a_col, meta, sl_rec);
          tot += 2.0 * (double)x;
        }
      }
    }
    const int cnt = __popc(before);
    if (!__any_sync(0xffffffffu, cnt != 0)) return;
    int pre = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    if (nx + total > K3S_XCAP) cross_flush(true);
    if (total > K3S_XCAP) {                           // (dense deletions, forced): in place
      T x = T(0);
      while (before) {
        const int e = __ffs(before) - 1;
        before &= before - 1u;
        x += k3s_cross<T>(s, sp + e, slot_col, a_ptr, a_col, meta, sl_rec);
      }
      tot += 2.0 * (double)x;
      return;
    }
    int at = nx + pre - cnt;
    while (before) {
      const int e = __ffs(before) - 1;
      before &= before - 1u;
      X1[at] = s;
      X2[at] = sp + e;
      ++at;
    }
    nx += total;
    __syncwarp();
    if (nx >= 32) cross_flush(false);
  };

  for (int64_t w0 = c_lo + (int64_t)warp * 32; w0 < c_hi; w0 += (K3S_THREADS / 32) * 32) {
    const int64_t w = w0 + lane;
    uint32_t z = (w < c_hi) ? ~mrow[w] : 0u;
    // exclusive prefix of the per-lane zero-bit counts
    const int cnt = __popc(z);
    if (w < c_hi) ones += 32 - cnt;                   // the bits past E in the last word are zero
    int pre = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    if (total == 0) continue;
    int at = nl + pre - cnt;
    while (z) {
      const int bit = __ffs(z) - 1;
      z &= z - 1u;
      L[at++] = (int32_t)(w * 32 + bit);
    }
    nl += total;
    __syncwarp();
    int done = 0;
    while (nl - done >= 32) { pass(done, 32); done += 32; }
    const int rest = nl - done;                       // < 32: move to the front
    const int mv = (lane < rest) ? L[done + lane] : 0;
    __syncwarp();
    if (lane < rest) L[lane] = mv;
    nl = rest;
    __syncwarp();
  }
  if (nl) pass(0, nl);
  __syncwarp();
  cross_flush(true);

#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xffffffffu, tot, o);
  if (lane == 0) wsum[warp] = tot;
  if (nnz) {                                          // integer atomics: order-independent
    ones = __reduce_add_sync(0xffffffffu, ones);
    if (lane == 0 && ones) atomicAdd(reinterpret_cast<unsigned long long*>(nnz + b), (unsigned long long)ones);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = (blockIdx.y == 0) ? base_sum : 0.0;
    for (int x = 0; x < K3S_THREADS / 32; ++x) s += wsum[x];
    partial[(int64_t)blockIdx.y * Bp + b] = s;
  }
}

}  // namespace spai
