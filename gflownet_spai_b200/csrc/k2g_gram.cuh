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

template <typename T, int K>
struct GramGeom {
  static constexpr int NG = K * (K + 1) / 2;                   // packed lower triangle, row-major
  static constexpr int NV = NG + K;                            // + b
  static constexpr int VB = (NV * (int)sizeof(T) + 15) / 16 * 16;
  static constexpr int RB = (int)sizeof(GramHdr) + VB;         // bytes per row
  static constexpr int R = (K <= 8) ? 16 : 8;                  // rows per stage
  static constexpr int SMEM = 64 + 2 * R * RB;
};

__host__ __device__ constexpr int gram_idx(int i, int l) { return i * (i + 1) / 2 + l; }   // l <= i

// relative pivot below which a pair goes to the Householder kernel
template <typename T> struct GramTol;
template <> struct GramTol<float>  { static constexpr float  v = 1e-3f; };
template <> struct GramTol<double> { static constexpr double v = 1e-6;  };

// ---------------------------------------------------------------- build (once per context)
// One warp per row of the class list. Walks the row's records segment by segment (records are
// sorted by (output column, slot)); a segment is one row of D.
template <typename T, int K>
__global__ void __launch_bounds__(128)
k2g_build_kernel(const typename RecOf<T>::type* __restrict__ recs, const int64_t* __restrict__ cptr,
                 const int32_t* __restrict__ sptr, const int32_t* __restrict__ r_diag,
                 const double* __restrict__ row_base, const int32_t* __restrict__ rows, int64_t nrows,
                 unsigned char* __restrict__ gram) {
  using Geo = GramGeom<T, K>;
  constexpr int PAIRS = (Geo::NG + 31) / 32;
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
  double bval = 0.0;                            // lane e (< K) keeps b[e]
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
    const int last = ends ? __ffs(ends) - 1 : 31;               // a segment has <= k <= 16 records
    if (lane <= last && in) vec[warp][rec_e(fl)] = a;
    __syncwarp();
#pragma unroll
    for (int p = 0; p < PAIRS; ++p)
      if (p * 32 + lane < Geo::NG) acc[p] = fma(vec[warp][pi[p]], vec[warp][pl[p]], acc[p]);
    if (seg == diag && lane < K) bval = vec[warp][lane];
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
    if (x < Geo::NG) {
      double g = acc[p];
      if (pi[p] >= k) g = (pi[p] == pl[p]) ? 1.0 : 0.0;          // padding slots: identity
      v[x] = (T)g;
    }
  }
  if (lane < K) v[Geo::NG + lane] = (lane < k) ? (T)bval : T(0);
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
    bad |= keep && !(d > GramTol<T>::v * g[gram_idx(j, j)]);
    const T inv = keep ? T(1) / d : T(0);
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
  using Geo = GramGeom<T, K>;
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

}  // namespace spai
