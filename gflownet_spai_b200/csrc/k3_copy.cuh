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
// range of rows (tiles of <= TILE_C plan records staged in shared memory) and
// THREADS*NT trajectories; row sums are reduced in registers over the block's
// rows, written once as partial[row_block][b], and combined in a fixed order
// by the finalize kernel (deterministic, no atomics).
//
// Per record: 1 broadcast LDS.128 + NT x (mask test, predicated add); per output
// segment: NT x (subtract delta, FMA into the row sum).
#pragma once

#include "spai_internal.cuh"

namespace spai {

constexpr int K3_TILE_C = 2048;      // records per staged tile (32 KB)
constexpr int K3_THREADS = 256;

template <typename T, int NT, typename Rec>
__device__ __forceinline__ void k3_row_fast(const Rec* __restrict__ rp, int cnt, int sp, int k,
                                            const uint32_t* __restrict__ maskT, int64_t Bp,
                                            const int64_t (&b)[NT], double (&tot)[NT]) {
  const int64_t w0 = sp >> 5;
  const int sh = sp & 31;
  const bool two = sh + k > 32;
  const uint32_t kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
  uint32_t m[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const uint32_t lo = maskT[w0 * Bp + b[j]];
    const uint32_t hi = two ? maskT[(w0 + 1) * Bp + b[j]] : 0u;
    m[j] = __funnelshift_r(lo, hi, sh) & kmask;
  }
  T acc[NT], rs[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) { acc[j] = T(0); rs[j] = T(0); }
#pragma unroll 4
  for (int c = 0; c < cnt; ++c) {
    const Rec r = rp[c];
    const T w = rec_w(r);
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j] += (m[j] & r.ebit) ? w : T(0);
    if (r.flags & F_END) {
      const T d = (r.flags & F_DIAG) ? T(1) : T(0);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const T t = acc[j] - d;
        rs[j] = fma(t, t, rs[j]);
        acc[j] = T(0);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) tot[j] += (double)rs[j];
}

// rows with more than 32 candidate slots: the kept bit of every record is read
// from the mask on demand (coalesced across lanes; L1-resident within a row).
template <typename T, int NT, typename Rec>
__device__ __forceinline__ void k3_row_wide(const Rec* __restrict__ rp, int cnt, int sp,
                                            const uint32_t* __restrict__ maskT, int64_t Bp,
                                            const int64_t (&b)[NT], double (&tot)[NT]) {
  T acc[NT], rs[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) { acc[j] = T(0); rs[j] = T(0); }
  for (int c = 0; c < cnt; ++c) {
    const Rec r = rp[c];
    const T w = rec_w(r);
    const int64_t bit = (int64_t)sp + rec_e(r.flags);
    const uint32_t* wp = maskT + (bit >> 5) * Bp;
    const int sh = bit & 31;
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j] += ((wp[b[j]] >> sh) & 1u) ? w : T(0);
    if (r.flags & F_END) {
      const T d = (r.flags & F_DIAG) ? T(1) : T(0);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const T t = acc[j] - d;
        rs[j] = fma(t, t, rs[j]);
        acc[j] = T(0);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) tot[j] += (double)rs[j];
}

template <typename T, int NT>
__global__ void __launch_bounds__(K3_THREADS)
k3_copy_kernel(const typename RecOf<T>::type* __restrict__ recs, const int64_t* __restrict__ cptr,
               const int32_t* __restrict__ sptr, const int32_t* __restrict__ tile_row, int ntiles,
               const uint32_t* __restrict__ maskT, int64_t Bp, double* __restrict__ partial) {
  using Rec = typename RecOf<T>::type;
  extern __shared__ __align__(16) unsigned char k3_smem[];
  Rec* tile = reinterpret_cast<Rec*>(k3_smem);

  const int t0 = (int)((int64_t)ntiles * blockIdx.x / gridDim.x);
  const int t1 = (int)((int64_t)ntiles * (blockIdx.x + 1) / gridDim.x);
  const int64_t bbase = (int64_t)blockIdx.y * (K3_THREADS * NT);
  int64_t b[NT];
  bool live[NT];
  double tot[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int64_t bj = bbase + (int64_t)j * K3_THREADS + threadIdx.x;
    live[j] = bj < Bp;
    b[j] = live[j] ? bj : 0;
    tot[j] = 0.0;
  }

  for (int t = t0; t < t1; ++t) {
    const int r0 = tile_row[t], r1 = tile_row[t + 1];
    const int64_t c0 = cptr[r0], c1 = cptr[r1];
    const bool staged = (c1 - c0) <= K3_TILE_C;
    __syncthreads();                       // previous tile fully consumed
    if (staged) {
      const int cnt = (int)(c1 - c0);
      const uint4* src = reinterpret_cast<const uint4*>(recs + c0);
      uint4* dst = reinterpret_cast<uint4*>(tile);
      for (int x = threadIdx.x; x < cnt; x += K3_THREADS) dst[x] = __ldcs(src + x);
    }
    __syncthreads();
    for (int i = r0; i < r1; ++i) {
      const int64_t cb = cptr[i], ce = cptr[i + 1];
      if (cb == ce) continue;
      const int sp = sptr[i];
      const int k = sptr[i + 1] - sp;
      const int cnt = (int)(ce - cb);
      if (k <= 32) {
        if (staged) k3_row_fast<T, NT, Rec>(tile + (cb - c0), cnt, sp, k, maskT, Bp, b, tot);
        else        k3_row_fast<T, NT, Rec>(recs + cb, cnt, sp, k, maskT, Bp, b, tot);
      } else {
        if (staged) k3_row_wide<T, NT, Rec>(tile + (cb - c0), cnt, sp, maskT, Bp, b, tot);
        else        k3_row_wide<T, NT, Rec>(recs + cb, cnt, sp, maskT, Bp, b, tot);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NT; ++j)
    if (live[j]) partial[(int64_t)blockIdx.x * Bp + b[j]] = tot[j];
}

// residual^2 = sum_g partial[g][b] + rows_missing_diag (each such row adds the
// uncovered -1 of -I); residual = sqrt; reward mix of preconditioner.py:154-163
// and :64, all in fp64. res2_in (optional) is added as well (generic-path sums).
__global__ void k3_finalize_kernel(const double* __restrict__ partial, int nparts, int64_t Bp,
                                   int64_t B, const double* __restrict__ res2_extra,
                                   double rows_missing_diag, const long long* __restrict__ nnz,
                                   double n, double res0, double flops0, double alpha,
                                   double* __restrict__ reward, double* __restrict__ residual,
                                   long long* __restrict__ nnz_out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = rows_missing_diag;
  for (int g = 0; g < nparts; ++g) s += partial[(int64_t)g * Bp + b];
  if (res2_extra) s += res2_extra[b];
  const double res = sqrt(s);
  const long long z = nnz[b];
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
