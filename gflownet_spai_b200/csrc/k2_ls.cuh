// K2 — batched small dense Householder-QR least squares ("ls" mode, the SPAI
// re-solve named by BASELINE.json's north star; the reference has no such code).
//
// For row i with kept candidate columns J (a subset of the row's candidate
// slots) and union index set I_i the problem is
//     min_m || A(J, I_i)^T m - e_i(I_i) ||_2 ,      row residual^2 = ||.||^2
// (+1 if i is not in I_i: the uncovered diagonal of -I; added by the finalize
// kernel through rows_missing_diag).
//
// Register kernel (k2_ls_kernel): the trajectory-independent dense tile
// D = A(S_i, I_i)^T (|I_i| x k, gathered by K1) is staged once per (block, row)
// in shared memory; a group of G lanes owns one (row, trajectory) problem with
// the tile's rows distributed cyclically over the G lanes and held in registers
// (QL rows x KMAX columns per lane, statically indexed). Columns of removed
// candidates are zeroed by the kept-mask, which leaves the least-squares
// residual unchanged; reflectors of zero / dependent columns are skipped and the
// pivot row p advances only on active columns. Column norms and v^T a_c dot
// products are reduced over the G lanes with xor-shuffles. All FP64/FP32 CUDA
// cores; the tiles (<= 64 x 32) are far too small for tensor-core MMA shapes.
//
// Generic kernel (k2_ls_generic_kernel): one warp per problem with the compacted
// tile in a global scratch slab — any |I_i| and k, used for the rows outside the
// register kernels' classes.
#pragma once

#include <utility>

#include "spai_internal.cuh"

namespace spai {

template <int G> __device__ __forceinline__ float k2_gsum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int G> __device__ __forceinline__ double k2_gsum(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T> struct K2Tol;
template <> struct K2Tol<float>  { static constexpr float  v = 1e-10f; };   // (~85 eps)^2
template <> struct K2Tol<double> { static constexpr double v = 1e-26;  };   // (~450 eps)^2

__device__ __forceinline__ float k2_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double k2_rsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ float k2_rcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double k2_rcp(double x) { return __drcp_rn(x); }

// One Householder step on column E of the lane-distributed tile. E is a
// template parameter so that every register-array index is a compile-time
// constant (a runtime column loop would demote `a` to local memory). `kkw` is
// the largest number of kept columns among the problems of this warp: the kept
// columns are compacted to the front when the tile is loaded, so steps and
// trailing columns >= kkw are all-zero for every lane and are skipped with
// warp-uniform branches. Full column rank is assumed, i.e. the pivot row of step
// E is row E: "row r >= E" is then a compile-time fact for all but the one
// register row that straddles E (rows are dealt cyclically, r = t*G + lg), which
// removed ~half of the issued instructions (64-bit selects on a runtime pivot).
// A tile that turns out rank-deficient sets `bad` and is redone by the generic
// kernel (same hand-over list as the column kernel).
template <typename T, int KMAX, int G, int QL, int E>
__device__ __forceinline__ void k2_step(T (&a)[QL][KMAX], T (&y)[QL], T (&cn2)[KMAX], int lg, int kk,
                                        int kkw, bool& bad) {
  if (E >= kkw) return;
  constexpr int TB = E / G;                 // register row that holds matrix row E (on lane E % G)
  constexpr int LB = E % G;
  if constexpr (TB < QL) {
    // ||a[E:, E]||^2 : rows t > TB are below the pivot on every lane, row TB only on lanes >= LB
    T sig = (lg >= LB) ? a[TB][E] * a[TB][E] : T(0);
    T sig1 = T(0);                            // two chains: FMA latency binds at this occupancy
#pragma unroll
    for (int t = TB + 1; t < QL; ++t) {
      if ((t - TB) & 1) sig1 = fma(a[t][E], a[t][E], sig1); else sig = fma(a[t][E], a[t][E], sig);
    }
    sig = k2_gsum<G>(sig + sig1);
    const T alp = __shfl_sync(0xffffffffu, a[TB][E], LB, G);          // pivot element a[E][E]
    const bool act = sig > cn2[E] * K2Tol<T>::v;
    const T sg = act ? sig : T(1);
    const T nrm = sg * k2_rsqrt(sg);
    const T beta = (alp >= T(0)) ? -nrm : nrm;
    const T inv = act ? k2_rcp(fma(-alp, beta, sg)) : T(0);
    bad |= (E < kk) && !act;
    // v: row E -> alp - beta, rows > E -> a[.,E], rows < E -> 0 (only register row TB is lane-dependent)
    const T vb = (lg > LB) ? a[TB][E] : ((lg == LB) ? alp - beta : T(0));
#pragma unroll
    for (int c = E + 1; c < KMAX; ++c) {
      if (c < kkw) {
        T dot = vb * a[TB][c];
        T dot1 = T(0);
#pragma unroll
        for (int t = TB + 1; t < QL; ++t) {
          if ((t - TB) & 1) dot1 = fma(a[t][E], a[t][c], dot1); else dot = fma(a[t][E], a[t][c], dot);
        }
        dot = k2_gsum<G>(dot + dot1);
        const T f = dot * inv;
        a[TB][c] = fma(-f, vb, a[TB][c]);
#pragma unroll
        for (int t = TB + 1; t < QL; ++t) a[t][c] = fma(-f, a[t][E], a[t][c]);
      }
    }
    {
      T dot = vb * y[TB];
      T dot1 = T(0);
#pragma unroll
      for (int t = TB + 1; t < QL; ++t) {
        if ((t - TB) & 1) dot1 = fma(a[t][E], y[t], dot1); else dot = fma(a[t][E], y[t], dot);
      }
      dot = k2_gsum<G>(dot + dot1);
      const T f = dot * inv;
      y[TB] = fma(-f, vb, y[TB]);
#pragma unroll
      for (int t = TB + 1; t < QL; ++t) y[t] = fma(-f, a[t][E], y[t]);
    }
  }
}

template <typename T, int KMAX, int G, int QL, int... Es>
__device__ __forceinline__ void k2_all_steps(T (&a)[QL][KMAX], T (&y)[QL], T (&cn2)[KMAX], int lg, int kk,
                                             int kkw, bool& bad, std::integer_sequence<int, Es...>) {
  (k2_step<T, KMAX, G, QL, Es>(a, y, cn2, lg, kk, kkw, bad), ...);
}

constexpr int K2_NW = 4;          // warps per block
constexpr int K2_MAX_NTG = 16;    // trajectories per lane group

template <typename T, int KMAX, int G, int QL>
// fp64, G = 2: 204 registers allow 2 CTAs/SM and the kernel is latency-bound (ncu: 12 % warps
// active, stall = wait); capping at 168 registers (3 CTAs/SM, ~200 B of spills) is 1.25x faster,
// 128 registers (4 CTAs/SM) spills ~1 KB and is slower again.
__global__ void __launch_bounds__(K2_NW * 32, (sizeof(T) == 8 && G == 2) ? 3 : 1)
k2_ls_kernel(const typename RecOf<T>::type* __restrict__ recs, const int64_t* __restrict__ cptr,
             const int32_t* __restrict__ sptr, const int32_t* __restrict__ r_diag,
             const int32_t* __restrict__ rows, int64_t nrows, const uint32_t* __restrict__ maskT,
             int64_t Bp, int ntg, double* __restrict__ partial, const double* __restrict__ row_base,
             int64_t B, int2* __restrict__ fail_pairs, unsigned int* __restrict__ fail_count,
             unsigned int fail_cap) {
  using Rec = typename RecOf<T>::type;
  constexpr int QP = QL * G;
  constexpr int NG = 32 / G;
  constexpr int NTHREADS = K2_NW * 32;
  constexpr int GROUPS = K2_NW * NG;
  __shared__ __align__(16) T tile[QP * KMAX];
  __shared__ T coln2[KMAX];
  __shared__ double totsm[K2_MAX_NTG * GROUPS];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int grp = lane / G, lg = lane % G;
  const int gid = warp * NG + grp;
  const int64_t bbase = (int64_t)blockIdx.y * ((int64_t)GROUPS * ntg);

  for (int x = tid; x < K2_MAX_NTG * GROUPS; x += NTHREADS) totsm[x] = 0.0;

  const int64_t ri0 = nrows * blockIdx.x / gridDim.x;
  const int64_t ri1 = nrows * (blockIdx.x + 1) / gridDim.x;
  for (int64_t ri = ri0; ri < ri1; ++ri) {
    const int i = rows[ri];
    const int64_t cb = cptr[i], ce = cptr[i + 1];
    const int sp = sptr[i];
    const int k = sptr[i + 1] - sp;
    const int diag = r_diag[i];
    __syncthreads();
    for (int x = tid; x < QP * KMAX; x += NTHREADS) tile[x] = T(0);
    __syncthreads();
    for (int64_t c = cb + tid; c < ce; c += NTHREADS) {
      const Rec r = recs[c];
      tile[rec_s(r.flags) * KMAX + rec_e(r.flags)] = rec_a(r);
    }
    __syncthreads();
    if (tid < KMAX) {
      T s = T(0);
      for (int q = 0; q < QP; ++q) { const T v = tile[q * KMAX + tid]; s = fma(v, v, s); }
      coln2[tid] = s;
    }
    __syncthreads();

    const int64_t w0 = sp >> 5;
    const int sh = sp & 31;
    const bool two = sh + k > 32;
    const uint32_t kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);

#pragma unroll 1
    for (int j = 0; j < ntg; ++j) {
      const int64_t breal = bbase + (int64_t)j * GROUPS + gid;
      const int64_t b = (breal < Bp) ? breal : 0;           // result discarded below
      const uint32_t lo = maskT[w0 * Bp + b];
      const uint32_t hi = two ? maskT[(w0 + 1) * Bp + b] : 0u;
      const uint32_t m = __funnelshift_r(lo, hi, sh) & kmask;
      // incremental path: no problem of this warp lost a candidate of this row
      if (__all_sync(0xffffffffu, m == kmask)) {
        if (lg == 0) totsm[j * GROUPS + gid] += row_base[i];
        continue;
      }

      // compact the kept columns to the front: column c' <- c'-th kept slot
      T a[QL][KMAX];
      T y[QL];
      T cn2[KMAX];
      uint32_t mm = m;
#pragma unroll
      for (int cc = 0; cc < KMAX; ++cc) {
        const int e = __ffs(mm) - 1;           // -1 once the kept slots are exhausted
        mm &= mm - 1u;
        cn2[cc] = (e >= 0) ? coln2[e] : T(0);
#pragma unroll
        for (int t = 0; t < QL; ++t) a[t][cc] = (e >= 0) ? tile[(t * G + lg) * KMAX + e] : T(0);
      }
#pragma unroll
      for (int t = 0; t < QL; ++t) y[t] = ((t * G + lg) == diag) ? T(1) : T(0);
      const int kk = __popc(m);
      const int kkw = __reduce_max_sync(0xffffffffu, kk);
      bool bad = false;
      k2_all_steps<T, KMAX, G, QL>(a, y, cn2, lg, kk, kkw, bad, std::make_integer_sequence<int, KMAX>{});
      T r2 = T(0);
#pragma unroll
      for (int t = 0; t < QL; ++t) {
        const int r = t * G + lg;
        r2 += (r >= kk) ? y[t] * y[t] : T(0);               // full rank: the first kk rows are eliminated
      }
      r2 = k2_gsum<G>(r2);
      if (lg == 0) {
        if (bad && breal < B) {                              // rank-deficient tile: redone by the generic kernel
          const unsigned int slot = atomicAdd(fail_count, 1u);
          if (slot < fail_cap) fail_pairs[slot] = make_int2(i, (int)breal);
          r2 = T(0);
        }
        totsm[j * GROUPS + gid] += (double)r2;
      }
    }
  }
  __syncthreads();
  for (int x = tid; x < ntg * GROUPS; x += NTHREADS) {
    const int j = x / GROUPS, g = x % GROUPS;
    const int64_t b = bbase + (int64_t)j * GROUPS + g;
    if (b < Bp) partial[(int64_t)blockIdx.x * Bp + b] = totsm[x];
  }
}

// ---------------------------------------------------------------- column-per-lane kernel
// For k > 8 the row-distributed layout above is bound by warp shuffles (one
// xor-butterfly per trailing column and step: the shuffle unit retires one warp
// instruction per SM and clock). Here a problem owns W lanes (W = 16 or 32) and
// lane c holds COLUMN c of the compacted tile in registers (QMAX rows, static
// indices), so v^T a_c and the update are lane-local; the reflector v of step E
// is published by its owner lane through a per-problem shared-memory buffer and
// read back as warp-broadcast LDS.128. The right-hand side e_i lives in shared
// memory with its rows dealt over the W lanes (one butterfly per step). Full
// column rank is assumed (pivot row == step index: every index is static);
// a tile that turns out rank-deficient is pushed to `fail_pairs` and redone by
// the generic kernel.
// One Householder step; E is a RUNTIME value (the step loop is not unrolled: a
// fully unrolled 32-step body is ~180 KB of straight-line SASS that every warp
// streams once per problem — instruction fetch became the bottleneck). Register
// indices stay static because the column is SHIFTED UP by one row per step (the
// update writes row r into register r-1, the finished pivot row drops out and a
// zero enters at the bottom): the pivot is always a[0], the live rows are always
// a[0 .. QMAX-1-E], and no per-row predicates or selects are needed (a version
// with `r >= E` selects issued 2.2x more instructions, ALU-pipe bound per ncu).
// The right-hand side lives in shared memory and is indexed with the offset E.
template <typename T, int W, int QMAX>
__device__ __forceinline__ void k2c_step(T (&a)[QMAX], T cn, T* __restrict__ vb, T* __restrict__ yb,
                                         int lc, int kk, int E, bool& bad) {
  T sg4[4] = {T(0), T(0), T(0), T(0)};                    // 4 chains: FMA latency binds at low occupancy
#pragma unroll
  for (int r = 0; r < QMAX; ++r) sg4[r & 3] = fma(a[r], a[r], sg4[r & 3]);
  const T sig = (sg4[0] + sg4[1]) + (sg4[2] + sg4[3]);
  const T alp = a[0];
  const bool act = sig > cn * K2Tol<T>::v;
  const T sg = act ? sig : T(1);
  const T nrm = sg * k2_rsqrt(sg);
  const T beta = (alp >= T(0)) ? -nrm : nrm;
  const T inv_o = act ? k2_rcp(fma(-alp, beta, sg)) : T(0);
  if (lc == E) {                                          // owner publishes v
    vb[0] = alp - beta;
#pragma unroll
    for (int r = 1; r < QMAX; ++r) vb[r] = a[r];
  }
  const T inv = __shfl_sync(0xffffffffu, inv_o, E, W);
  bad |= (E < kk) && (inv == T(0));
  __syncwarp();
  // trailing columns (lanes > E): dot, update + shift
  T dt4[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
  for (int r = 0; r < QMAX; ++r) dt4[r & 3] = fma(vb[r], a[r], dt4[r & 3]);
  const T dot = (dt4[0] + dt4[1]) + (dt4[2] + dt4[3]);
  const T f = (lc > E) ? dot * inv : T(0);
  if constexpr (sizeof(T) == 8) {
    // fp64: re-read v from shared memory (volatile) instead of keeping the QMAX values of the dot
    // pass live: 244 -> 164 registers, 3 CTAs/SM (cfg4 +8 %; for fp32 the hoisted copy is faster)
    const volatile T* vbv = vb;
#pragma unroll
    for (int r = 1; r < QMAX; ++r) a[r - 1] = fma(-f, vbv[r], a[r]);
  } else {
#pragma unroll
    for (int r = 1; r < QMAX; ++r) a[r - 1] = fma(-f, vb[r], a[r]);
  }
  a[QMAX - 1] = T(0);
  // right-hand side: live rows E .. QMAX-1 dealt over the W lanes
  T py = T(0);
#pragma unroll
  for (int r = lc; r < QMAX; r += W) py = fma(vb[r], yb[E + r], py);
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) py += __shfl_xor_sync(0xffffffffu, py, o);
  const T fy = py * inv;
#pragma unroll
  for (int r = lc; r < QMAX; r += W) yb[E + r] = fma(-fy, vb[r], yb[E + r]);
  __syncwarp();
}

template <typename T, int W, int QMAX>
__global__ void __launch_bounds__(K2_NW * 32)          // (forcing 3 CTAs/SM for fp64 spilled the column: 1.9x slower)
k2c_ls_kernel(const typename RecOf<T>::type* __restrict__ recs, const int64_t* __restrict__ cptr,
              const int32_t* __restrict__ sptr, const int32_t* __restrict__ r_diag,
              const int32_t* __restrict__ rows, int64_t nrows, const uint32_t* __restrict__ maskT,
              int64_t Bp, int64_t B, int ntg, double* __restrict__ partial, int2* __restrict__ fail_pairs,
              unsigned int* __restrict__ fail_count, unsigned int fail_cap,
              const double* __restrict__ row_base) {
  using Rec = typename RecOf<T>::type;
  constexpr int NP = 32 / W;                       // problems per warp
  constexpr int NTHREADS = K2_NW * 32;
  constexpr int GROUPS = K2_NW * NP;
  __shared__ __align__(16) T tile[QMAX * W];       // [row][slot]
  __shared__ T coln2[W];
  __shared__ __align__(16) T vbuf[GROUPS][QMAX];
  __shared__ __align__(16) T ybuf[GROUPS][QMAX + W];      // + W: step E reads rows E .. E+QMAX-1
  __shared__ double totsm[K2_MAX_NTG * GROUPS];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int pr = lane / W, lc = lane % W;
  const int gid = warp * NP + pr;
  const int64_t bbase = (int64_t)blockIdx.y * ((int64_t)GROUPS * ntg);
  T* vb = vbuf[gid];
  T* yb = ybuf[gid];

  for (int x = tid; x < K2_MAX_NTG * GROUPS; x += NTHREADS) totsm[x] = 0.0;

  const int64_t ri0 = nrows * blockIdx.x / gridDim.x;
  const int64_t ri1 = nrows * (blockIdx.x + 1) / gridDim.x;
  for (int64_t ri = ri0; ri < ri1; ++ri) {
    const int i = rows[ri];
    const int64_t cb = cptr[i], ce = cptr[i + 1];
    const int sp = sptr[i];
    const int k = sptr[i + 1] - sp;
    const int diag = r_diag[i];
    __syncthreads();
    for (int x = tid; x < QMAX * W; x += NTHREADS) tile[x] = T(0);
    __syncthreads();
    for (int64_t c = cb + tid; c < ce; c += NTHREADS) {
      const Rec r = recs[c];
      tile[rec_s(r.flags) * W + rec_e(r.flags)] = rec_a(r);
    }
    __syncthreads();
    if (tid < W) {
      T s = T(0);
      for (int q = 0; q < QMAX; ++q) { const T v = tile[q * W + tid]; s = fma(v, v, s); }
      coln2[tid] = s;
    }
    __syncthreads();

    const int64_t w0 = sp >> 5;
    const int sh = sp & 31;
    const bool two = sh + k > 32;
    const uint32_t kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);

#pragma unroll 1
    for (int j = 0; j < ntg; ++j) {
      const int64_t breal = bbase + (int64_t)j * GROUPS + gid;
      const int64_t b = (breal < Bp) ? breal : 0;
      const uint32_t lo = maskT[w0 * Bp + b];
      const uint32_t hi = two ? maskT[(w0 + 1) * Bp + b] : 0u;
      const uint32_t m = __funnelshift_r(lo, hi, sh) & kmask;
      if (__all_sync(0xffffffffu, m == kmask)) {            // incremental path (see k2_ls_kernel)
        if (lc == 0) totsm[j * GROUPS + gid] += row_base[i];
        continue;
      }
      const int kk = __popc(m);
      const int kkw = __reduce_max_sync(0xffffffffu, kk);
      const int e = (lc < kk) ? (int)__fns(m, 0, lc + 1) : -1;       // my column = lc-th kept slot
      T a[QMAX];
#pragma unroll
      for (int r = 0; r < QMAX; ++r) a[r] = (e >= 0) ? tile[r * W + e] : T(0);
      const T cn = (e >= 0) ? coln2[e] : T(0);
#pragma unroll
      for (int r = lc; r < QMAX + W; r += W) yb[r] = (r == diag) ? T(1) : T(0);
      __syncwarp();
      bool bad = false;
#pragma unroll 1
      for (int E = 0; E < kkw; ++E) k2c_step<T, W, QMAX>(a, cn, vb, yb, lc, kk, E, bad);
      T r2 = T(0);
#pragma unroll
      for (int r = lc; r < QMAX; r += W) r2 += (r >= kk) ? yb[r] * yb[r] : T(0);
#pragma unroll
      for (int o = W / 2; o > 0; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      __syncwarp();
      if (lc == 0) {
        if (bad && breal < B) {
          const unsigned int slot = atomicAdd(fail_count, 1u);
          if (slot < fail_cap) fail_pairs[slot] = make_int2(i, (int)breal);
          r2 = T(0);                                       // redone by the generic kernel
        }
        totsm[j * GROUPS + gid] += (double)r2;
      }
    }
  }
  __syncthreads();
  for (int x = tid; x < ntg * GROUPS; x += NTHREADS) {
    const int j = x / GROUPS, g = x % GROUPS;
    const int64_t b = bbase + (int64_t)j * GROUPS + g;
    if (b < Bp) partial[(int64_t)blockIdx.x * Bp + b] = totsm[x];
  }
}

// ---------------------------------------------------------------- generic path
// One warp per (row, trajectory). Scratch per warp: kk columns of q entries
// (column-major) + y[q]. Any k (kept bits fetched word by word) and any q.
template <typename T>
__device__ __forceinline__ T k2_wsum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(128)
k2_ls_generic_kernel(const typename RecOf<T>::type* __restrict__ recs,
                     const int64_t* __restrict__ cptr, const int32_t* __restrict__ sptr,
                     const int32_t* __restrict__ r_q, const int32_t* __restrict__ r_diag,
                     const int32_t* __restrict__ rows, int64_t nrows,
                     const uint32_t* __restrict__ maskT, int64_t Bp, int64_t B, T* work,
                     int64_t work_stride, int32_t* colmap, int64_t colmap_stride,
                     double* __restrict__ res2, const int2* __restrict__ pairs,
                     const unsigned int* __restrict__ npairs, unsigned int pair_cap,
                     double* __restrict__ row_out, int64_t row_out_ld = 1) {
  using Rec = typename RecOf<T>::type;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  T* A = work + warp * work_stride;
  int32_t* cmap = colmap + warp * colmap_stride;
  // two modes: every (row of `rows`) x (trajectory), or the (row, trajectory)
  // pairs a register kernel could not finish (rank-deficient tiles)
  const int64_t items = pairs ? (int64_t)min(*npairs, pair_cap) : nrows * B;
  for (int64_t it = warp; it < items; it += nwarps) {
    int64_t b;
    int i;
    if (pairs) { i = pairs[it].x; b = pairs[it].y; }
    else { i = rows[it / B]; b = it % B; }
    const int q = r_q[i];
    const int sp = sptr[i];
    const int k = sptr[i + 1] - sp;
    const int diag = r_diag[i];
    const int64_t cb = cptr[i], ce = cptr[i + 1];
    if (q == 0) continue;                        // nothing gathered: row adds only the constant
    // compact the kept columns: cmap[e] = column index or -1
    int kk = 0;
    for (int e0 = 0; e0 < k; e0 += 32) {
      const int e = e0 + lane;
      bool kept = false;
      if (e < k) {
        const int64_t bit = (int64_t)sp + e;
        kept = (maskT[(bit >> 5) * Bp + b] >> (bit & 31)) & 1u;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, kept);
      if (e < k) cmap[e] = kept ? kk + __popc(bal & ((1u << lane) - 1u)) : -1;
      kk += __popc(bal);
    }
    T* y = A + (int64_t)kk * q;
    for (int64_t x = lane; x < (int64_t)kk * q + q; x += 32) A[x] = T(0);
    __syncwarp();
    for (int64_t c = cb + lane; c < ce; c += 32) {
      const Rec r = recs[c];
      const int col = cmap[rec_e(r.flags)];
      if (col >= 0) A[(int64_t)col * q + rec_s(r.flags)] = rec_a(r);
    }
    if (lane == 0 && diag >= 0) y[diag] = T(1);
    __syncwarp();
    int p = 0;
    for (int j = 0; j < kk && p < q; ++j) {
      T* aj = A + (int64_t)j * q;
      T full = T(0), sig = T(0);
      for (int r = lane; r < q; r += 32) {
        const T v = aj[r];
        full = fma(v, v, full);
        if (r >= p) sig = fma(v, v, sig);
      }
      full = k2_wsum(full);
      sig = k2_wsum(sig);
      if (!(sig > full * K2Tol<T>::v)) continue;          // dependent / zero column
      const T alp = aj[p];
      const T nrm = sqrt(sig);
      const T beta = (alp >= T(0)) ? -nrm : nrm;
      const T inv = T(1) / (sig - alp * beta);
      __syncwarp();
      if (lane == 0) aj[p] = alp - beta;
      __syncwarp();
      for (int c = j + 1; c <= kk; ++c) {                  // c == kk is the right-hand side
        T* ac = A + (int64_t)c * q;
        T dot = T(0);
        for (int r = p + lane; r < q; r += 32) dot = fma(aj[r], ac[r], dot);
        dot = k2_wsum(dot);
        const T f = dot * inv;
        for (int r = p + lane; r < q; r += 32) ac[r] = fma(-f, aj[r], ac[r]);
      }
      __syncwarp();
      ++p;
    }
    T r2 = T(0);
    for (int r = p + lane; r < q; r += 32) r2 = fma(y[r], y[r], r2);
    r2 = k2_wsum(r2);
    if (lane == 0) {
      if (row_out) row_out[(int64_t)i * row_out_ld + b] = (double)r2;   // per-row mode: row_out[row][trajectory]
      else atomicAdd(res2 + b, (double)r2);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- solve values
// ls-mode values of M for ONE trajectory (column 0 of maskT): one warp per row,
// Householder QR as above but R is kept (diagonal in `beta`, pivot rows in
// `piv`) and the triangular system is solved by back substitution. Output in
// EDGE order (slot_edge), 0 for removed / dependent candidates.
template <typename T>
__global__ void __launch_bounds__(128)
k2_ls_solve_kernel(const typename RecOf<T>::type* __restrict__ recs, const int64_t* __restrict__ cptr,
                   const int32_t* __restrict__ sptr, const int32_t* __restrict__ slot_edge,
                   const int32_t* __restrict__ r_q, const int32_t* __restrict__ r_diag, int64_t n,
                   const uint32_t* __restrict__ maskT, int64_t Bp, T* work, int64_t work_stride,
                   int32_t* imap, int64_t imap_stride, double* __restrict__ m_val) {
  using Rec = typename RecOf<T>::type;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const int q = r_q[i];
    const int sp = sptr[i];
    const int k = sptr[i + 1] - sp;
    if (q == 0 || k == 0) continue;
    const int diag = r_diag[i];
    const int64_t cb = cptr[i], ce = cptr[i + 1];
    int32_t* cmap = imap + warp * imap_stride;          // [k] slot -> column
    int32_t* piv = cmap + k;                            // [k] column -> pivot row or -1
    int kk = 0;
    for (int e0 = 0; e0 < k; e0 += 32) {
      const int e = e0 + lane;
      bool kept = false;
      if (e < k) {
        const int64_t bit = (int64_t)sp + e;
        kept = (maskT[(bit >> 5) * Bp] >> (bit & 31)) & 1u;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, kept);
      if (e < k) cmap[e] = kept ? kk + __popc(bal & ((1u << lane) - 1u)) : -1;
      kk += __popc(bal);
    }
    T* A = work + warp * work_stride;                   // kk columns of q, then y[q], beta[kk], x[kk]
    T* y = A + (int64_t)kk * q;
    T* beta_v = y + q;
    T* x = beta_v + kk;
    for (int64_t t = lane; t < (int64_t)kk * q + q + 2 * kk; t += 32) A[t] = T(0);
    __syncwarp();
    for (int64_t c = cb + lane; c < ce; c += 32) {
      const Rec r = recs[c];
      const int col = cmap[rec_e(r.flags)];
      if (col >= 0) A[(int64_t)col * q + rec_s(r.flags)] = rec_a(r);
    }
    if (lane == 0 && diag >= 0) y[diag] = T(1);
    __syncwarp();
    int p = 0;
    for (int j = 0; j < kk; ++j) {
      T* aj = A + (int64_t)j * q;
      T full = T(0), sig = T(0);
      for (int r = lane; r < q; r += 32) {
        const T v = aj[r];
        full = fma(v, v, full);
        if (r >= p) sig = fma(v, v, sig);
      }
      full = k2_wsum(full);
      sig = k2_wsum(sig);
      if (p >= q || !(sig > full * K2Tol<T>::v)) {
        if (lane == 0) piv[j] = -1;
        continue;
      }
      const T alp = aj[p];
      const T nrm = sqrt(sig);
      const T beta = (alp >= T(0)) ? -nrm : nrm;
      const T inv = T(1) / (sig - alp * beta);
      __syncwarp();
      if (lane == 0) { aj[p] = alp - beta; beta_v[j] = beta; piv[j] = p; }
      __syncwarp();
      for (int c = j + 1; c <= kk; ++c) {
        T* ac = A + (int64_t)c * q;
        T dot = T(0);
        for (int r = p + lane; r < q; r += 32) dot = fma(aj[r], ac[r], dot);
        dot = k2_wsum(dot);
        const T f = dot * inv;
        for (int r = p + lane; r < q; r += 32) ac[r] = fma(-f, aj[r], ac[r]);
      }
      __syncwarp();
      ++p;
    }
    __syncwarp();
    // back substitution over the active columns (R[piv[j]][c] lives in A[c][piv[j]])
    for (int j = kk - 1; j >= 0; --j) {
      const int pj = piv[j];
      if (pj < 0) continue;                    // uniform: piv is read by all lanes
      T acc = T(0);
      for (int c = j + 1 + lane; c < kk; c += 32)
        if (piv[c] >= 0) acc = fma(A[(int64_t)c * q + pj], x[c], acc);
      acc = k2_wsum(acc);
      if (lane == 0) x[j] = (y[pj] - acc) / beta_v[j];
      __syncwarp();
    }
    for (int e = lane; e < k; e += 32) {
      const int col = cmap[e];
      if (col >= 0) m_val[slot_edge[sp + e]] = (double)x[col];
    }
    __syncwarp();
  }
}

}  // namespace spai
