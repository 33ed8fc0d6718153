// K3 — fused row-of-(M*A) residual, per-row norm and Frobenius reduction
// ("copy" mode: surviving entries keep their values, the reference's semantics).
//
// Replaces preconditioner.py:82-93 (sparse identity, torch.mm SpGEMM, sparse
// subtraction, torch.norm) for a whole batch of trajectories.
//
// Mapping: the 32 lanes of a warp are 32 TRAJECTORIES working on the same row
// of M, so every read of the row's gathered tile is a warp-uniform broadcast
// from shared memory and control flow is uniform; the only per-lane state is
// the row's kept-mask word and the running sums. A block owns a contiguous
// range of row tiles and THREADS*NT trajectories. Tiles (<= K3_TILE_C plan
// records + their row headers) are brought into shared memory by
// cp.async.bulk (TMA 1-D bulk copy) into a two-stage ring signalled through
// mbarriers, so the copy of tile t+1 overlaps the arithmetic of tile t. Row
// sums are reduced in registers over the block's rows, written once as
// partial[row_block][b], and combined in a fixed order by the finalize kernel
// (deterministic, no atomics).
//
// Kept-mask words are read from maskT[w][b] (one coalesced 128-byte load per
// warp and word) and carried in registers across the rows that share a word.
//
// Per record: 1 broadcast LDS.128 + NT x (bit test -> predicate, predicated
// FADD); per output segment: NT x FFMA into the row sum (the accumulator starts
// at -delta_ij, so no separate subtraction).
#pragma once

#include "spai_internal.cuh"

namespace spai {

constexpr int K3_TILE_C = 1024;      // records per staged tile (16 KB)
constexpr int K3_TILE_R = 128;       // rows per staged tile (2 KB of row headers)
constexpr int K3_LIST = 128;         // per-warp compaction list (touching trajectories of a row)
constexpr int K3_THREADS = 128;
constexpr int K3_STAGE_BYTES = (K3_TILE_C + 1) * 16 + K3_TILE_R * 16;   // +1: prefetch slack
constexpr int K3_WARPS = 4;
// the compaction lists exist only in the COMPACT variant (6 vs 5 CTAs/SM for fp32)
template <typename T> constexpr int k3_smem_bytes(bool compact) {
  return 2 * K3_STAGE_BYTES + 64 + (compact ? K3_WARPS * K3_LIST * (4 + (int)sizeof(T)) : 0);
}

// ---- mbarrier / bulk-copy primitives (sm_90+ PTX; SASS: SYNCS / UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void k3_cond_add(float& acc, uint32_t m, uint32_t ebit, float w) {
  asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
      "and.b32 t, %1, %2;\n\t"
      "setp.ne.u32 p, t, 0;\n\t"
      "@p add.f32 %0, %0, %3;\n\t}"
      : "+f"(acc)
      : "r"(m), "r"(ebit), "f"(w));
}
__device__ __forceinline__ void k3_cond_add(double& acc, uint32_t m, uint32_t ebit, double w) {
  asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
      "and.b32 t, %1, %2;\n\t"
      "setp.ne.u32 p, t, 0;\n\t"
      "@p add.f64 %0, %0, %3;\n\t}"
      : "+d"(acc)
      : "r"(m), "r"(ebit), "d"(w));
}

// Row residual^2 with EVERY candidate kept, one thread per row, same operation
// order as the main kernel. Used for the incremental fast path: a row none of
// whose slots was removed by any trajectory of the warp is not re-evaluated.
template <typename T>
__global__ void k3_row_base_kernel(const typename RecOf<T>::type* __restrict__ recs,
                                   const int64_t* __restrict__ cptr, const RowHdr* __restrict__ rhdr,
                                   int64_t n, T* __restrict__ base) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T acc = (rhdr[i].flags & 1) ? T(-1) : T(0);
  T rs = T(0);
  for (int64_t c = cptr[i]; c < cptr[i + 1]; ++c) {
    const auto r = recs[c];
    acc += rec_w(r);
    if (r.flags & F_END) {
      rs = fma(acc, acc, rs);
      acc = (r.flags & F_NEXT_DIAG) ? T(-1) : T(0);
    }
  }
  base[i] = rs;
}

// The per-lane mask window: words wcur and wcur+1 of every trajectory handled by
// the lane (trajectory j of the lane is column `j * K3_THREADS` after `mp`);
// advanced monotonically as the rows walk through the slot range.
template <int NT>
struct MaskWindow {
  int64_t wcur;
  uint32_t lo[NT], hi[NT];
  __device__ __forceinline__ void load(const uint32_t* __restrict__ mp, int64_t Bp, int64_t W, int64_t w) {
    wcur = w;
    const uint32_t* p0 = mp + w * Bp;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      lo[j] = (w < W) ? p0[j * K3_THREADS] : 0u;
      hi[j] = (w + 1 < W) ? p0[Bp + j * K3_THREADS] : 0u;
    }
  }
  __device__ __forceinline__ void seek(const uint32_t* __restrict__ mp, int64_t Bp, int64_t W, int64_t w) {
    if (w == wcur) return;
    if (w == wcur + 1) {
      wcur = w;
      const uint32_t* p1 = mp + (w + 1) * Bp;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        lo[j] = hi[j];
        hi[j] = (w + 1 < W) ? p1[j * K3_THREADS] : 0u;
      }
      return;
    }
    load(mp, Bp, W, w);
  }
};

// One trajectory per lane: the compacted path below.
template <typename T, typename Rec>
__device__ __forceinline__ T k3_row_single(const Rec* __restrict__ rp, int cnt, bool first_diag, uint32_t m) {
  T acc = first_diag ? T(-1) : T(0);
  T rs = T(0);
  int c = 0;
  while (c < cnt) {
    int fl;
    do {
      const Rec r = rp[c++];
      fl = (int)r.flags;
      k3_cond_add(acc, m, r.ebit, rec_w(r));
    } while (fl >= 0);
    rs = fma(acc, acc, rs);
    acc = ((uint32_t)fl & F_NEXT_DIAG) ? T(-1) : T(0);
  }
  return rs;
}

// One row for NT trajectories per lane. The record stream is walked segment by
// segment (do-while up to the END record: a real branch, so the per-segment
// epilogue is not issued for every record).
// Incremental evaluation: a trajectory that removed no candidate of
// this row contributes the cached all-kept row residual `base`. If no trajectory
// of the warp touched the row it is skipped; if only a few did (<= 12 per NT
// slot, at most K3_LIST * 3/4), their masks are compacted into a per-warp
// shared-memory list (ballot + popc) and evaluated 32 at a time, one trajectory
// per lane; otherwise all 32*NT run (cfg5 with 1 % deletions: 66.7 -> 36.9 ms;
// the dense headline workload pays ~3 % for the votes).
template <typename T, int NT, bool COMPACT, typename Rec>
__device__ __forceinline__ void k3_row_fast(const Rec* __restrict__ rp, int cnt, int sp, int k,
                                            bool first_diag, T base, MaskWindow<NT>& mw,
                                            const uint32_t* __restrict__ mp, int64_t Bp, int64_t W,
                                            uint32_t* __restrict__ lm, T* __restrict__ lv, T (&rs)[NT]) {
  mw.seek(mp, Bp, W, sp >> 5);
  const int sh = sp & 31;
  const uint32_t kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  uint32_t m[NT];
  uint32_t all = kmask;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    m[j] = __funnelshift_r(mw.lo[j], mw.hi[j], sh) & kmask;
    all &= m[j];
  }
  // one vote first: while every lane holds at least one touching trajectory (more than
  // ~1/3 of them touch) the per-slot ballots below are never issued
  const unsigned lanes_touch = __ballot_sync(0xffffffffu, all != kmask);
  if (lanes_touch == 0) {
#pragma unroll
    for (int j = 0; j < NT; ++j) rs[j] += base;
    return;
  }
  if (COMPACT && lanes_touch != 0xffffffffu) {
    unsigned bal[NT];
    int touching = 0;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      bal[j] = __ballot_sync(0xffffffffu, m[j] != kmask);
      touching += __popc(bal[j]);
    }
    if (touching <= (K3_LIST * 3) / 4 && touching <= 12 * NT) {
      int at = 0;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if (m[j] != kmask) lm[at + __popc(bal[j] & lt)] = m[j]; else rs[j] += base;
        at += __popc(bal[j]);
      }
      __syncwarp();
      for (int p0 = 0; p0 < touching; p0 += 32) {
        const int idx = p0 + (threadIdx.x & 31);
        const uint32_t mm = (idx < touching) ? lm[idx] : kmask;
        const T v = k3_row_single<T, Rec>(rp, cnt, first_diag, mm);
        if (idx < touching) lv[idx] = v;
      }
      __syncwarp();
      at = 0;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if (m[j] != kmask) rs[j] += lv[at + __popc(bal[j] & lt)];
        at += __popc(bal[j]);
      }
      __syncwarp();
      return;
    }
  }
  T acc[NT];
  const T a0 = first_diag ? T(-1) : T(0);
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j] = a0;
  int c = 0;
  while (c < cnt) {
    int fl;
    do {
      const Rec r = rp[c++];
      fl = (int)r.flags;
      const T w = rec_w(r);
#pragma unroll
      for (int j = 0; j < NT; ++j) k3_cond_add(acc[j], m[j], r.ebit, w);
    } while (fl >= 0);                                   // END is the sign bit
    const T nxt = ((uint32_t)fl & F_NEXT_DIAG) ? T(-1) : T(0);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      rs[j] = fma(acc[j], acc[j], rs[j]);
      acc[j] = nxt;
    }
  }
}

// rows with more than 32 candidate slots: the kept bit of every record is read
// from the mask on demand (coalesced across lanes; L1-resident within a row).
template <typename T, int NT, typename Rec>
__device__ __forceinline__ void k3_row_wide(const Rec* __restrict__ rp, int cnt, int sp,
                                            bool first_diag, const uint32_t* __restrict__ mp,
                                            int64_t Bp, T (&rs)[NT]) {
  T acc[NT];
  const T a0 = first_diag ? T(-1) : T(0);
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j] = a0;
  for (int c = 0; c < cnt; ++c) {
    const Rec r = rp[c];
    const T w = rec_w(r);
    const int64_t bit = (int64_t)sp + rec_e(r.flags);
    const uint32_t* wp = mp + (bit >> 5) * Bp;
    const int sh = bit & 31;
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j] += ((wp[j * K3_THREADS] >> sh) & 1u) ? w : T(0);
    if (r.flags & F_END) {
      const T nxt = (r.flags & F_NEXT_DIAG) ? T(-1) : T(0);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        rs[j] = fma(acc[j], acc[j], rs[j]);
        acc[j] = nxt;
      }
    }
  }
}

// COMPACT = false drops the compaction code (a few % faster when most trajectories touch
// most rows: the host picks it from the trajectory length, see eval_masks).
template <typename T, int NT, bool COMPACT>
__global__ void __launch_bounds__(K3_THREADS)
k3_copy_kernel(const typename RecOf<T>::type* __restrict__ recs, const int64_t* __restrict__ cptr,
               const RowHdr* __restrict__ rhdr, const T* __restrict__ row_base,
               const int32_t* __restrict__ tile_row, int ntiles,
               const uint32_t* __restrict__ maskT, int64_t Bp, int64_t W,
               double* __restrict__ partial, int row_lo, int row_hi) {
  using Rec = typename RecOf<T>::type;
  extern __shared__ __align__(128) unsigned char k3_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(k3_smem);                    // 2 mbarriers
  unsigned char* stage0 = k3_smem + 64;
  unsigned char* lists = stage0 + 2 * K3_STAGE_BYTES;                       // per-warp compaction lists
  T* lv = reinterpret_cast<T*>(lists) + (threadIdx.x >> 5) * K3_LIST;
  uint32_t* lm = reinterpret_cast<uint32_t*>(lists + K3_WARPS * K3_LIST * sizeof(T)) + (threadIdx.x >> 5) * K3_LIST;

  const int t0 = (int)((int64_t)ntiles * blockIdx.x / gridDim.x);
  const int t1 = (int)((int64_t)ntiles * (blockIdx.x + 1) / gridDim.x);
  // trajectory j of this lane: b = bbase + j*K3_THREADS + tid. maskT is padded to
  // a multiple of K3_THREADS*NT columns by the host, so every column is readable.
  const int64_t bbase = (int64_t)blockIdx.y * (K3_THREADS * NT);
  const uint32_t* mp = maskT + bbase + threadIdx.x;
  double tot[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) tot[j] = 0.0;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // producer: one thread issues the bulk copies of a tile into stage s
  auto issue = [&](int t, int s) {
    const int r0 = tile_row[t], r1 = tile_row[t + 1];
    const int64_t c0 = cptr[r0], c1 = cptr[r1];
    unsigned char* st = stage0 + (size_t)s * K3_STAGE_BYTES;
    const uint32_t hbytes = (uint32_t)(r1 - r0) * 16u;
    const bool staged = (c1 - c0) <= K3_TILE_C;
    const uint32_t rbytes = staged ? (uint32_t)(c1 - c0) * 16u : 0u;
    mbar_expect_tx(&bars[s], hbytes + rbytes);
    bulk_g2s(st + (K3_TILE_C + 1) * 16, rhdr + r0, hbytes, &bars[s]);
    if (rbytes) bulk_g2s(st, recs + c0, rbytes, &bars[s]);
  };

  uint32_t phase[2] = {0u, 0u};
  if (t0 < t1 && threadIdx.x == 0) issue(t0, 0);
  MaskWindow<NT> mw;
  mw.wcur = -2;
  for (int t = t0; t < t1; ++t) {
    const int s = (t - t0) & 1;
    if (t + 1 < t1 && threadIdx.x == 0) issue(t + 1, s ^ 1);   // stage s^1 was released by the last sync
    mbar_wait(&bars[s], phase[s]);
    phase[s] ^= 1u;
    const unsigned char* st = stage0 + (size_t)s * K3_STAGE_BYTES;
    const Rec* tile = reinterpret_cast<const Rec*>(st);
    const RowHdr* hdr = reinterpret_cast<const RowHdr*>(st + (K3_TILE_C + 1) * 16);
    const int r0 = tile_row[t], r1 = tile_row[t + 1];
    const int64_t c0 = cptr[r0];
    const bool staged = (cptr[r1] - c0) <= K3_TILE_C;
    T rs[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) rs[j] = T(0);
    int off = 0;
    for (int i = 0; i < r1 - r0; ++i) {
      const RowHdr h = hdr[i];
      if (h.cnt == 0) continue;
      if (r0 + i < row_lo || r0 + i >= row_hi) { off += h.cnt; continue; }     // row-range evaluation
      const bool fd = h.flags & 1;
      if (h.k <= 32) {
        const T base = row_base[r0 + i];
        if (staged) k3_row_fast<T, NT, COMPACT, Rec>(tile + off, h.cnt, h.sp, h.k, fd, base, mw, mp, Bp, W, lm, lv, rs);
        else        k3_row_fast<T, NT, COMPACT, Rec>(recs + c0 + off, h.cnt, h.sp, h.k, fd, base, mw, mp, Bp, W, lm, lv, rs);
      } else {
        if (staged) k3_row_wide<T, NT, Rec>(tile + off, h.cnt, h.sp, fd, mp, Bp, rs);
        else        k3_row_wide<T, NT, Rec>(recs + c0 + off, h.cnt, h.sp, fd, mp, Bp, rs);
      }
      off += h.cnt;
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) tot[j] += (double)rs[j];   // fp32 sums live for one tile only
    __syncthreads();                       // stage s fully consumed -> may be refilled
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int64_t bj = bbase + (int64_t)j * K3_THREADS + threadIdx.x;
    if (bj < Bp) partial[(int64_t)blockIdx.x * Bp + bj] = tot[j];
  }
}

// residual^2 = sum_g partial[g][b] + rows_missing_diag (each such row adds the
// uncovered -1 of -I); residual = sqrt; reward mix of preconditioner.py:154-163
// and :64, all in fp64. res2_extra (optional) is added as well (generic-path sums).
__global__ void k3_finalize_kernel(const double* __restrict__ partial, int nparts, int64_t Bp,
                                   int64_t B, const double* __restrict__ res2_extra,
                                   double rows_missing_diag, const long long* __restrict__ nnz,
                                   double n, double res0, double flops0, double alpha,
                                   double* __restrict__ reward, double* __restrict__ residual,
                                   long long* __restrict__ nnz_out,
                                   const unsigned int* __restrict__ fail_count, unsigned int fail_cap,
                                   int partial_only) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = rows_missing_diag;
  // more rank-deficient tiles than the hand-over list can hold: fail loudly (NaN)
  if (fail_count && *fail_count > fail_cap) s = __longlong_as_double(0x7ff8000000000000LL);
  for (int g = 0; g < nparts; ++g) s += partial[(int64_t)g * Bp + b];
  if (res2_extra) s += res2_extra[b];
  const long long z = nnz[b];
  if (partial_only) {                 // row-range evaluation: the caller sums over ranks first
    if (residual) residual[b] = s;
    if (nnz_out) nnz_out[b] = z;
    return;
  }
  const double res = sqrt(s < 0.0 ? 0.0 : s);      // (the delta form of K3s can land a hair below zero; NaN stays NaN)
  const double flops = 2.0 * (double)z * n;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const double rr = (res0 != 0.0) ? res / res0 : inf;
  const double cr = (flops0 != 0.0) ? flops / flops0 : inf;
  const double metric = alpha * (1.0 - rr) + (1.0 - alpha) * (1.0 - cr);
  if (reward) reward[b] = metric * 1000.0;
  if (residual) residual[b] = res;
  if (nnz_out) nnz_out[b] = z;
}

}  // namespace spai
