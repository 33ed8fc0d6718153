// K3t — copy-mode row residuals by table lookup, for patterns whose rows have at most 8
// candidates (BASELINE.json configs[0] and [1], the headline).
//
// A row with k <= 8 candidates has at most 256 kept-masks, and a batch holds thousands of
// trajectories, so the same (row, mask) pair is evaluated over and over by the row sweep (K3).
// k3t_build_kernel evaluates every (row, mask) ONCE per context — with K3's own row routine, so
// the table holds exactly the numbers K3 would produce — and a (row, trajectory) evaluation
// becomes: extract the row's 8 mask bits, one shared-memory load, one add. The layout is K3's:
// the 32 lanes of a warp are 32 trajectories on the same row, a block owns a range of rows x
// 128*NT trajectories, the table rows of the range arrive by cp.async.bulk into a 2-stage
// mbarrier ring (1 KB per row in fp32), fixed-order partial sums.
//
// Replaces the same reference lines as K3 (preconditioner.py:79-93 on the pattern of
// gflownet/utils.py:315-353).
#pragma once

#include "k2g_gram.cuh"
#include "k3_copy.cuh"
#include "spai_internal.cuh"

namespace spai {

constexpr int K3T_K = 8;                 // candidates per row the table covers
constexpr int K3T_ENTRIES = 1 << K3T_K;
template <typename T> __host__ __device__ constexpr int k3t_rows() { return 64 / (int)sizeof(T); }   // table rows per stage (16 KB)

template <typename T> __host__ __device__ constexpr int k3t_stage_bytes() { return k3t_rows<T>() * (K3T_ENTRIES * (int)sizeof(T) + 16); }
template <typename T> __host__ __device__ constexpr int k3t_smem_bytes() { return 64 + 2 * k3t_stage_bytes<T>(); }

// one block per row, one thread per mask
template <typename T>
__global__ void __launch_bounds__(K3T_ENTRIES)
k3t_build_kernel(const typename RecOf<T>::type* __restrict__ recs, const int64_t* __restrict__ cptr,
                 const RowHdr* __restrict__ rhdr, T* __restrict__ lut) {
  using Rec = typename RecOf<T>::type;
  const int64_t row = blockIdx.x;
  const RowHdr h = rhdr[row];
  const uint32_t m = threadIdx.x;
  T v = T(0);
  if (h.cnt > 0 && h.k <= K3T_K && m < (1u << h.k))
    v = k3_row_single<T, Rec>(recs + cptr[row], h.cnt, h.flags & 1, m);
  lut[row * K3T_ENTRIES + m] = v;
}

// ls_gram mode: the same table holds the least-squares residual of every (row, kept-mask),
// solved once per context from the row's Gram record (k2g_solve). A mask whose elimination meets a
// small pivot is stored as NaN; the lookup kernel hands such (row, trajectory) pairs to the
// Householder kernel through the usual fail list.
template <typename T>
__global__ void __launch_bounds__(K3T_ENTRIES)
k3t_build_ls_kernel(const unsigned char* __restrict__ gram, T* __restrict__ lut) {
  using Geo = GramGeom<T, 8, false>;
  const unsigned char* rec = gram + (int64_t)blockIdx.x * Geo::RB;
  const GramHdr h = *reinterpret_cast<const GramHdr*>(rec);
  const uint32_t m = threadIdx.x;
  if (m >= (1u << h.k)) return;                       // the table was zero-filled
  bool bad = false;
  const T acc = k2g_solve<T, 8>(reinterpret_cast<const T*>(rec + sizeof(GramHdr)), m, bad);
  T r2 = (T)h.ee - acc;
  r2 = r2 > T(0) ? r2 : T(0);
  if (bad) r2 = (T)__longlong_as_double(0x7ff8000000000000LL);
  lut[(int64_t)h.row * K3T_ENTRIES + m] = r2;
}

// Householder (`ls`) mode: the table is filled by the generic QR kernel itself, run on 256
// pseudo-trajectories — pseudo-trajectory m keeps, in EVERY row, exactly the candidates whose bit is
// set in m. This kernel writes that mask (transposed layout, Bp = 256 columns, zero-filled before).
__global__ void k3t_uniform_masks_kernel(const int32_t* __restrict__ sptr, int64_t n, uint32_t* __restrict__ maskT) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = idx >> K3T_K;
  const uint32_t m = (uint32_t)(idx & (K3T_ENTRIES - 1));
  if (row >= n) return;
  const int sp = sptr[row], k = sptr[row + 1] - sp;
  for (int j = 0; j < k && j < K3T_K; ++j)
    if ((m >> j) & 1u) atomicOr(&maskT[(int64_t)((sp + j) >> 5) * K3T_ENTRIES + m], 1u << ((sp + j) & 31));
}

template <typename T, int NT, bool FAILS = false>
__global__ void __launch_bounds__(K3_THREADS)
k3t_lookup_kernel(const T* __restrict__ lut, const RowHdr* __restrict__ rhdr,
                  const uint32_t* __restrict__ maskT, int64_t Bp, int64_t W,
                  double* __restrict__ partial, int row_lo, int row_hi, int64_t B = 0,
                  int2* __restrict__ fail_pairs = nullptr, unsigned int* __restrict__ fail_count = nullptr,
                  unsigned int fail_cap = 0) {
  extern __shared__ __align__(128) unsigned char k3t_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(k3t_smem);
  unsigned char* stage0 = k3t_smem + 64;
  constexpr int K3T_ROWS = k3t_rows<T>();
  constexpr int TB = K3T_ROWS * K3T_ENTRIES * (int)sizeof(T);      // table bytes of a stage

  const int64_t nrows = row_hi - row_lo;
  const int r0 = row_lo + (int)(nrows * blockIdx.x / gridDim.x);
  const int r1 = row_lo + (int)(nrows * (blockIdx.x + 1) / gridDim.x);
  const int64_t bbase = (int64_t)blockIdx.y * (K3_THREADS * NT);
  const uint32_t* mp = maskT + bbase + threadIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int r, int s) {
    const int cnt = (r1 - r < K3T_ROWS) ? r1 - r : K3T_ROWS;
    unsigned char* st = stage0 + (size_t)s * k3t_stage_bytes<T>();
    mbar_expect_tx(&bars[s], (uint32_t)cnt * (K3T_ENTRIES * (uint32_t)sizeof(T) + 16u));
    bulk_g2s(st, lut + (int64_t)r * K3T_ENTRIES, (uint32_t)cnt * K3T_ENTRIES * (uint32_t)sizeof(T), &bars[s]);
    bulk_g2s(st + TB, rhdr + r, (uint32_t)cnt * 16u, &bars[s]);
  };
  uint32_t phase[2] = {0u, 0u};
  if (r0 < r1 && threadIdx.x == 0) issue(r0, 0);
  MaskWindow<NT> mw;
  mw.wcur = -2;
  double tot[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) tot[j] = 0.0;
  int s = 0;
  for (int r = r0; r < r1; r += K3T_ROWS, s ^= 1) {
    if (r + K3T_ROWS < r1 && threadIdx.x == 0) issue(r + K3T_ROWS, s ^ 1);
    mbar_wait(&bars[s], phase[s]);
    phase[s] ^= 1u;
    const unsigned char* st = stage0 + (size_t)s * k3t_stage_bytes<T>();
    const T* tab = reinterpret_cast<const T*>(st);
    const RowHdr* hdr = reinterpret_cast<const RowHdr*>(st + TB);
    const int cnt = (r1 - r < K3T_ROWS) ? r1 - r : K3T_ROWS;
    T rs[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) rs[j] = T(0);
    for (int i = 0; i < cnt; ++i) {
      const RowHdr h = hdr[i];
      if (h.cnt == 0) continue;
      mw.seek(mp, Bp, W, h.sp >> 5);
      const int sh = h.sp & 31;
      const uint32_t kmask = (1u << h.k) - 1u;
      const T* row = tab + i * K3T_ENTRIES;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        T v = row[__funnelshift_r(mw.lo[j], mw.hi[j], sh) & kmask];
        if (FAILS && v != v) {                                // ill-conditioned (row, mask): Householder redoes it
          const int64_t bj = bbase + (int64_t)j * K3_THREADS + threadIdx.x;
          if (bj < B) {
            const unsigned int slot = atomicAdd(fail_count, 1u);
            if (slot < fail_cap) fail_pairs[slot] = make_int2(r + i, (int)bj);
          }
          v = T(0);
        }
        rs[j] += v;
      }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) tot[j] += (double)rs[j];     // fp32 sums live for one stage only
    __syncthreads();                                          // stage s consumed
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int64_t bj = bbase + (int64_t)j * K3_THREADS + threadIdx.x;
    if (bj < Bp) partial[(int64_t)blockIdx.x * Bp + bj] = tot[j];
  }
}

}  // namespace spai
