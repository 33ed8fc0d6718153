// K2g — least-squares row residuals through the semi-normal equations ("ls_gram" mode).
//
// For row i the dense tile D = A(S_i, I_i)^T (|I_i| x k, candidates S_i of the row) does
// not depend on the trajectory; a trajectory only selects the kept columns J of D. With
//   G = D^T D  (k x k, symmetric),   b = D^T e_i = row `diag` of D,
// the least-squares residual of  min || D(:,J) m - e_i ||  is
//   res^2 = ||e_i||^2 - b_J^T (G_JJ)^-1 b_J,
// and G_JJ is a principal sub-matrix of G because a kept column is zero outside the union of
// the kept columns' supports. So the gather AND the Gram products are done once per context
// (k2g_build_kernel, fp64 accumulation), and a (row, trajectory) solve is one masked
// right-looking LDL^T elimination of a k x k matrix held in registers: ~k^3/6 + k^2 FMAs instead
// of the 2 |I| k^2 of Householder QR on the tile. The pivot d_j is the squared norm of column j
// orthogonal to the kept columns before it — the same quantity the QR kernels test — so a small
// relative pivot (ill-conditioned or rank-deficient tile, where squaring the condition number
// would cost accuracy) hands the (row, trajectory) pair to the Householder kernel through the
// same fail list the register QR kernels use.
//
// Layout of the solve kernel = K3's: the 32 lanes of a warp are 32 trajectories on the SAME row,
// so every read of G/b is a warp-uniform shared-memory broadcast; rows arrive by cp.async.bulk
// into a 2-stage mbarrier ring; untouched rows (no lane of the warp lost a candidate) add the
// cached all-kept residual.
//
// No reference counterpart (the reference copies values, SURVEY §8a "ls-mode additions");
// checked against oracle/spai_oracle.py:reward_batch_ls (Householder / lstsq).
#pragma once

#include <utility>

#include "k3_copy.cuh"
#include "spai_internal.cuh"

namespace spai {

struct alignas(16) GramHdr {     // 32 bytes in front of every row's values
  int32_t sp;                    // bit offset of the row's slots in the kept-mask
  int32_t k;                     // candidate slots of the row (<= K)
  int32_t row;                   // row index (fail list)
  int32_t ee;                    // 1 when the row index is in the union (||e_i||^2)
  double base;                   // all-kept residual^2 (incremental path)
  double pad;
};

// Values of a row: the Gram matrix, then b[K], then thr[K] = the smallest pivot accepted for
// column j (GramTol * G_jj, at least a tiny positive number).
// FULL = false: packed lower triangle (row-major) — the one-lane kernel.
// FULL = true : the whole symmetric matrix, row stride K + 1 values (bank spread), then b —
//               the lane-distributed kernel (a lane reads its own rows with fixed offsets).
template <typename T, int K, bool FULL>
struct GramGeom {
  static constexpr int KS = K + 1;
  static constexpr int NG = FULL ? K * KS : K * (K + 1) / 2;
  static constexpr int NV = NG + 2 * K;                        // + b + pivot thresholds
  static constexpr int VB = (NV * (int)sizeof(T) + 15) / 16 * 16;
  static constexpr int RB = (int)sizeof(GramHdr) + VB;         // bytes per row
  static constexpr int R = (8192 / RB) < 1 ? 1 : (8192 / RB > 16 ? 16 : 8192 / RB);   // rows per stage
  static constexpr int STAGES = 64 + 2 * R * RB;
};

__host__ __device__ constexpr int gram_idx(int i, int l) { return i * (i + 1) / 2 + l; }   // l <= i

// relative pivot below which a pair goes to the Householder kernel
template <typename T> struct GramTol;
template <> struct GramTol<float>  { static constexpr float  v = 1e-3f; static constexpr float  tiny = 1e-30f; };
template <> struct GramTol<double> { static constexpr double v = 1e-6;  static constexpr double tiny = 1e-280; };

// 1/d for a positive normal d: hardware approximation + Newton steps (no IEEE slow path; the
// pivot test has already rejected zeros, denormals and NaNs when the result is used)
__device__ __forceinline__ float gram_rcp(float d) {
  float x;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(d));
  return fmaf(x, fmaf(-d, x, 1.0f), x);
}
__device__ __forceinline__ double gram_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  x = fma(x, fma(-d, x, 1.0), x);
  return fma(x, fma(-d, x, 1.0), x);
}

// ---------------------------------------------------------------- build (once per context)
// One warp per row of the class list. Walks the row's records segment by segment (records are
// sorted by (output column, slot)); a segment is one row of D.
template <typename T, int K, bool FULL>
__global__ void __launch_bounds__(128)
k2g_build_kernel(const typename RecOf<T>::type* __restrict__ recs, const int64_t* __restrict__ cptr,
                 const int32_t* __restrict__ sptr, const int32_t* __restrict__ r_diag,
                 const double* __restrict__ row_base, const int32_t* __restrict__ rows, int64_t nrows,
                 unsigned char* __restrict__ gram) {
  using Geo = GramGeom<T, K, FULL>;
  constexpr int NTRI = K * (K + 1) / 2;
  constexpr int PAIRS = (NTRI + 31) / 32;
  __shared__ double vec[4][K];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ri = (int64_t)blockIdx.x * 4 + warp;
  if (ri >= nrows) return;
  const int i = rows[ri];
  const int64_t cb = cptr[i], ce = cptr[i + 1];
  const int sp = sptr[i], k = sptr[i + 1] - sp, diag = r_diag[i];
  int pi[PAIRS], pl[PAIRS];
  double acc[PAIRS];
#pragma unroll
  for (int p = 0; p < PAIRS; ++p) {
    const int x = p * 32 + lane;
    int a = 0;
    while (gram_idx(a + 1, 0) <= x) ++a;       // row of the packed index
    pi[p] = a; pl[p] = x - gram_idx(a, 0);
    acc[p] = 0.0;
  }
  double bval = 0.0, dval = 0.0;                // lane e (< K) keeps b[e] and G[e][e]
  if (lane < K) vec[warp][lane] = 0.0;
  __syncwarp();
  int seg = 0;
  int64_t c = cb;
  while (c < ce) {
    const bool in = c + lane < ce;
    uint32_t fl = 0u;
    double a = 0.0;
    if (in) { const auto r = recs[c + lane]; fl = r.flags; a = (double)rec_a(r); }
    const unsigned ends = __ballot_sync(0xffffffffu, in && (fl & F_END));
    const int last = ends ? __ffs(ends) - 1 : 31;               // a segment has <= k <= 32 records
    if (lane <= last && in) vec[warp][rec_e(fl)] = a;
    __syncwarp();
#pragma unroll
    for (int p = 0; p < PAIRS; ++p)
      if (p * 32 + lane < NTRI) acc[p] = fma(vec[warp][pi[p]], vec[warp][pl[p]], acc[p]);
    if (seg == diag && lane < K) bval = vec[warp][lane];
    if (lane < K) dval = fma(vec[warp][lane], vec[warp][lane], dval);
    __syncwarp();
    if (lane < K) vec[warp][lane] = 0.0;
    __syncwarp();
    c += last + 1;
    ++seg;
  }
  unsigned char* out = gram + ri * Geo::RB;
  if (lane == 0) {
    GramHdr h;
    h.sp = sp; h.k = k; h.row = i; h.ee = diag >= 0 ? 1 : 0; h.base = row_base[i]; h.pad = 0.0;
    *reinterpret_cast<GramHdr*>(out) = h;
  }
  T* v = reinterpret_cast<T*>(out + sizeof(GramHdr));
#pragma unroll
  for (int p = 0; p < PAIRS; ++p) {
    const int x = p * 32 + lane;
    if (x < NTRI) {
      double g = acc[p];
      if (pi[p] >= k) g = (pi[p] == pl[p]) ? 1.0 : 0.0;          // padding slots: identity
      if (FULL) { v[pi[p] * Geo::KS + pl[p]] = (T)g; v[pl[p] * Geo::KS + pi[p]] = (T)g; }
      else v[x] = (T)g;
    }
  }
  if (FULL) for (int x = lane; x < K; x += 32) v[x * Geo::KS + K] = T(0);       // row padding
  if (lane < K) {
    v[Geo::NG + lane] = (lane < k) ? (T)bval : T(0);
    const double gd = (lane < k) ? dval : 1.0;
    const double thr = fmax((double)GramTol<T>::v * gd, (double)GramTol<T>::tiny);
    v[Geo::NG + K + lane] = (T)thr;
  }
}

// ---------------------------------------------------------------- solve
// acc = b_J^T (G_JJ)^-1 b_J by a masked right-looking LDL^T: a removed column gets inv = 0, so it
// neither eliminates nor contributes; everything is unrolled onto registers.
template <typename T, int K>
__device__ __forceinline__ T k2g_solve(const T* __restrict__ g, uint32_t m, bool& bad) {
  constexpr int NG = K * (K + 1) / 2;
  T S[NG], b[K];
#pragma unroll
  for (int x = 0; x < NG; ++x) S[x] = g[x];
#pragma unroll
  for (int x = 0; x < K; ++x) b[x] = g[NG + x];
  T acc = T(0);
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const bool keep = (m >> j) & 1u;
    const T d = S[gram_idx(j, j)];
    bad |= keep && !(d > g[NG + K + j]);
    const T inv = keep ? gram_rcp(d) : T(0);
    const T t = b[j] * inv;
    acc = fma(b[j], t, acc);
#pragma unroll
    for (int i = j + 1; i < K; ++i) {
      const T sij = S[gram_idx(i, j)];
      const T f = sij * inv;
      b[i] = fma(-sij, t, b[i]);
#pragma unroll
      for (int l = j + 1; l <= i; ++l) S[gram_idx(i, l)] = fma(-f, S[gram_idx(l, j)], S[gram_idx(i, l)]);
    }
  }
  return acc;
}

template <typename T, int K>
__global__ void __launch_bounds__(K3_THREADS)
k2g_solve_kernel(const unsigned char* __restrict__ gram, int64_t nrows,
                 const uint32_t* __restrict__ maskT, int64_t Bp, int64_t W, int64_t B,
                 double* __restrict__ partial, int2* __restrict__ fail_pairs,
                 unsigned int* __restrict__ fail_count, unsigned int fail_cap) {
  using Geo = GramGeom<T, K, false>;
  extern __shared__ __align__(128) unsigned char k2g_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(k2g_smem);
  unsigned char* stage0 = k2g_smem + 64;

  const int64_t ri0 = nrows * blockIdx.x / gridDim.x;
  const int64_t ri1 = nrows * (blockIdx.x + 1) / gridDim.x;
  const int64_t breal = (int64_t)blockIdx.y * K3_THREADS + threadIdx.x;
  const bool live = breal < Bp;                      // Bp is a multiple of 32: warp-uniform
  const uint32_t* mp = maskT + (live ? breal : 0);
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int64_t r, int s) {
    const int64_t cnt = (ri1 - r < Geo::R) ? ri1 - r : Geo::R;
    const uint32_t bytes = (uint32_t)cnt * Geo::RB;
    mbar_expect_tx(&bars[s], bytes);
    bulk_g2s(stage0 + (size_t)s * Geo::R * Geo::RB, gram + r * Geo::RB, bytes, &bars[s]);
  };
  uint32_t phase[2] = {0u, 0u};
  if (ri0 < ri1 && threadIdx.x == 0) issue(ri0, 0);
  MaskWindow<1> mw;
  mw.wcur = -2;
  double tot = 0.0;
  int s = 0;
  for (int64_t r = ri0; r < ri1; r += Geo::R, s ^= 1) {
    if (r + Geo::R < ri1 && threadIdx.x == 0) issue(r + Geo::R, s ^ 1);
    mbar_wait(&bars[s], phase[s]);
    phase[s] ^= 1u;
    const unsigned char* st = stage0 + (size_t)s * Geo::R * Geo::RB;
    const int cnt = (int)((ri1 - r < Geo::R) ? ri1 - r : Geo::R);
    for (int x = 0; x < cnt; ++x) {
      const GramHdr h = *reinterpret_cast<const GramHdr*>(st + x * Geo::RB);
      mw.seek(mp, Bp, W, h.sp >> 5);
      const uint32_t kmask = (h.k >= 32) ? 0xffffffffu : ((1u << h.k) - 1u);
      const uint32_t m = __funnelshift_r(mw.lo[0], mw.hi[0], h.sp & 31) & kmask;
      if (__all_sync(0xffffffffu, m == kmask)) { tot += h.base; continue; }
      bool bad = false;
      const T acc = k2g_solve<T, K>(reinterpret_cast<const T*>(st + x * Geo::RB + sizeof(GramHdr)), m, bad);
      T r2 = (T)h.ee - acc;
      r2 = r2 > T(0) ? r2 : T(0);
      if (bad) {
        r2 = T(0);
        if (live && breal < B) {                     // redone by the Householder kernel
          const unsigned int slot = atomicAdd(fail_count, 1u);
          if (slot < fail_cap) fail_pairs[slot] = make_int2(h.row, (int)breal);
        }
      }
      tot += (double)r2;
    }
    __syncthreads();                                 // stage s consumed
  }
  if (live) partial[(int64_t)blockIdx.x * Bp + breal] = tot;
}

// ---------------------------------------------------------------- lane-distributed solve
// For k = 16 in fp64 and k = 32 the k x k matrix does not fit one lane's registers: G lanes share
// a problem, lane h owns rows i = r*G + h (block-cyclic), each stored up to column (r+1)*G - 1 (the
// few entries right of the diagonal are the symmetric copies and are updated alike, which keeps
// every register index a compile-time constant). Step j: every lane publishes its entries of
// column j into the problem's shared-memory line, the owner of row j also publishes 1/d_j and
// t_j = b_j/d_j; one __syncwarp; then each lane updates its own rows. A warp = 32/G trajectories
// on the same row, so the Gram reads stay warp-uniform per sub-lane.
template <typename T, int K, int G>
struct GramDist {
  static constexpr int NR = K / G;                       // rows per lane
  static constexpr int NS = G * NR * (NR + 1) / 2;       // stored entries per lane
  static constexpr int PPB = K3_THREADS / G;             // problems per block
  static constexpr int LINE = 2 * (K + 2);               // per problem: 2 buffers of (column, inv, t)
  static constexpr int SMEM = GramGeom<T, K, true>::STAGES + PPB * LINE * (int)sizeof(T);
  __host__ __device__ static constexpr int off(int r) { return G * r * (r + 1) / 2; }
};

template <typename T, int K, int G, int J>
__device__ __forceinline__ void k2gd_step(T (&S)[GramDist<T, K, G>::NS], T (&b)[K / G], T& acc, bool& bad,
                                          const T* __restrict__ g, uint32_t m, int h, T* __restrict__ line) {
  using D = GramDist<T, K, G>;
  constexpr int KS = K + 1, NR = K / G, RJ = J / G, HJ = J % G;
  T* cb = line + (J & 1) * (K + 2);
#pragma unroll
  for (int r = RJ; r < NR; ++r) cb[r * G + h] = S[D::off(r) + J];
  if (h == HJ) {
    const bool keep = (m >> J) & 1u;
    const T d = S[D::off(RJ) + J];
    bad |= keep && !(d > g[K * KS + K + J]);
    const T inv = keep ? gram_rcp(d) : T(0);
    const T t = b[RJ] * inv;
    acc = fma(b[RJ], t, acc);
    cb[K] = inv;
    cb[K + 1] = t;
  }
  __syncwarp();
  const T inv = cb[K], t = cb[K + 1];
  T c[K];
#pragma unroll
  for (int l = J + 1; l < K; ++l) c[l] = cb[l];
#pragma unroll
  for (int r = RJ; r < NR; ++r) {
    const T sij = S[D::off(r) + J];
    const T f = sij * inv;
    b[r] = fma(-sij, t, b[r]);
#pragma unroll
    for (int l = J + 1; l < (r + 1) * G; ++l) S[D::off(r) + l] = fma(-f, c[l], S[D::off(r) + l]);
  }
}

template <typename T, int K, int G, int... Js>
__device__ __forceinline__ void k2gd_steps(T (&S)[GramDist<T, K, G>::NS], T (&b)[K / G], T& acc, bool& bad,
                                           const T* __restrict__ g, uint32_t m, int h, T* __restrict__ line,
                                           std::integer_sequence<int, Js...>) {
  (k2gd_step<T, K, G, Js>(S, b, acc, bad, g, m, h, line), ...);
}

template <typename T, int K, int G>
__global__ void __launch_bounds__(K3_THREADS)
k2gd_solve_kernel(const unsigned char* __restrict__ gram, int64_t nrows,
                  const uint32_t* __restrict__ maskT, int64_t Bp, int64_t W, int64_t B,
                  double* __restrict__ partial, int2* __restrict__ fail_pairs,
                  unsigned int* __restrict__ fail_count, unsigned int fail_cap) {
  using Geo = GramGeom<T, K, true>;
  using D = GramDist<T, K, G>;
  constexpr int KS = K + 1, NR = K / G;
  extern __shared__ __align__(128) unsigned char k2g_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(k2g_smem);
  unsigned char* stage0 = k2g_smem + 64;
  const int h = threadIdx.x % G, pb = threadIdx.x / G;
  T* line = reinterpret_cast<T*>(k2g_smem + Geo::STAGES) + pb * D::LINE;

  const int64_t ri0 = nrows * blockIdx.x / gridDim.x;
  const int64_t ri1 = nrows * (blockIdx.x + 1) / gridDim.x;
  const int64_t breal = (int64_t)blockIdx.y * D::PPB + pb;
  const bool live = breal < Bp;                      // Bp % 32 == 0 and 32/G problems per warp: warp-uniform
  const uint32_t* mp = maskT + (live ? breal : 0);
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int64_t r, int s) {
    const int64_t cnt = (ri1 - r < Geo::R) ? ri1 - r : Geo::R;
    const uint32_t bytes = (uint32_t)cnt * Geo::RB;
    mbar_expect_tx(&bars[s], bytes);
    bulk_g2s(stage0 + (size_t)s * Geo::R * Geo::RB, gram + r * Geo::RB, bytes, &bars[s]);
  };
  uint32_t phase[2] = {0u, 0u};
  if (ri0 < ri1 && threadIdx.x == 0) issue(ri0, 0);
  MaskWindow<1> mw;
  mw.wcur = -2;
  double tot = 0.0;
  int s = 0;
  for (int64_t r = ri0; r < ri1; r += Geo::R, s ^= 1) {
    if (r + Geo::R < ri1 && threadIdx.x == 0) issue(r + Geo::R, s ^ 1);
    mbar_wait(&bars[s], phase[s]);
    phase[s] ^= 1u;
    const unsigned char* st = stage0 + (size_t)s * Geo::R * Geo::RB;
    const int cnt = (int)((ri1 - r < Geo::R) ? ri1 - r : Geo::R);
    if (live) {
      for (int x = 0; x < cnt; ++x) {
        const GramHdr hd = *reinterpret_cast<const GramHdr*>(st + x * Geo::RB);
        mw.seek(mp, Bp, W, hd.sp >> 5);
        const uint32_t kmask = (hd.k >= 32) ? 0xffffffffu : ((1u << hd.k) - 1u);
        const uint32_t m = __funnelshift_r(mw.lo[0], mw.hi[0], hd.sp & 31) & kmask;
        if (__all_sync(0xffffffffu, m == kmask)) { tot += hd.base; continue; }
        const T* g = reinterpret_cast<const T*>(st + x * Geo::RB + sizeof(GramHdr));
        T S[D::NS], b[NR];
        const T* gh = g + h * KS;
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
#pragma unroll
          for (int l = 0; l < (rr + 1) * G; ++l) S[D::off(rr) + l] = gh[rr * G * KS + l];
          b[rr] = g[Geo::NG + rr * G + h];
        }
        T acc = T(0);
        bool bad = false;
        k2gd_steps<T, K, G>(S, b, acc, bad, g, m, h, line, std::make_integer_sequence<int, K>{});
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
          acc += __shfl_xor_sync(0xffffffffu, acc, o);
          bad |= (bool)__shfl_xor_sync(0xffffffffu, (int)bad, o);
        }
        T r2 = (T)hd.ee - acc;
        r2 = r2 > T(0) ? r2 : T(0);
        if (bad) {
          r2 = T(0);
          if (h == 0 && breal < B) {
            const unsigned int slot = atomicAdd(fail_count, 1u);
            if (slot < fail_cap) fail_pairs[slot] = make_int2(hd.row, (int)breal);
          }
        }
        tot += (double)r2;
        __syncwarp();                                // the line buffers are reused by the next row
      }
    }
    __syncthreads();                                 // stage s consumed
  }
  if (live && h == 0) partial[(int64_t)blockIdx.x * Bp + breal] = tot;
}

}  // namespace spai
