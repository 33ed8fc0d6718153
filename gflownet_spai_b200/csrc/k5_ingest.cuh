// K5 — device-side ingest: the step immediately before the reward path.
//
// Replaces, on the GPU,
//   gflownet/utils.py:54-63      market_matrix_to_sparse_tensor: (MatrixMarket text ->) COO -> coalesced CSR
//   GFlowNet100.py:126-153       the initial matrix from a product of sparse factors (L @ U): CSR SpGEMM
//   this repo's synth.superset_pattern / synth.neumann_values (the SURVEY §8d candidate supersets
//   pattern(A) U ... U pattern(A^p) truncated to k per row by (distance, column id), and initial values
//   omega * sum_j (I - omega A)^j restricted to the superset)
//
// All kernels put one WARP on one matrix row: rows are short (<= a few hundred entries), so the row's
// working set (a small open-addressing hash of column -> value plus an insertion-ordered list) lives in
// shared memory; lanes walk the entries of one source row in parallel (distinct columns: no atomics needed
// for the values, only for claiming hash slots), source rows are taken in list order, so every sum has a
// FIXED order (deterministic results). Int/byte work: HBM/L2-bound gathers, no tensor cores.
#pragma once

#include "spai_internal.cuh"

namespace spai {

constexpr int K5_WARPS = 4;                 // warps (rows) per CTA
constexpr int K5_CAP = 2048;                // hash slots per row: a working set stops growing at K5_LIST + one source row (<= 512) entries
constexpr int K5_LIST = 512;                // list entries per row
constexpr int K5_EMPTY = -1;

struct K5Row {                              // per-warp shared-memory working set
  int key[K5_CAP];                          // column id or K5_EMPTY
  int slot_of[K5_LIST];                     // list position -> hash slot (insertion order)
  double val[K5_CAP];
  int aux[K5_LIST];                         // per-entry tag (BFS distance)
};

__device__ __forceinline__ unsigned k5_hash(int c) { return ((unsigned)c * 2654435761u) >> 21; }   // 11 bits

// Claim (or find) the slot of column c. Returns the slot; *fresh = this call inserted it.
__device__ __forceinline__ int k5_claim(K5Row& r, int c, bool* fresh) {
  unsigned h = k5_hash(c) & (K5_CAP - 1);
  for (;;) {
    const int old = atomicCAS(&r.key[h], K5_EMPTY, c);
    if (old == K5_EMPTY) { *fresh = true; return (int)h; }
    if (old == c) { *fresh = false; return (int)h; }
    h = (h + 1) & (K5_CAP - 1);
  }
}
__device__ __forceinline__ int k5_find(const K5Row& r, int c) {
  unsigned h = k5_hash(c) & (K5_CAP - 1);
  for (;;) {
    const int k = r.key[h];
    if (k == c) return (int)h;
    if (k == K5_EMPTY) return -1;
    h = (h + 1) & (K5_CAP - 1);
  }
}
__device__ __forceinline__ void k5_reset(K5Row& r, int lane) {
  for (int i = lane; i < K5_CAP; i += 32) { r.key[i] = K5_EMPTY; r.val[i] = 0.0; }
  __syncwarp();
}

// In-warp bitonic sort of m <= K5_LIST 64-bit keys held in shared memory (ascending).
__device__ inline void k5_sort(unsigned long long* keys, int m, int lane) {
  int p = 1;
  while (p < m) p <<= 1;
  for (int i = m + lane; i < p; i += 32) keys[i] = ~0ull;
  __syncwarp();
  for (int k = 2; k <= p; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < p; i += 32) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], b = keys[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[l] = a; }
        }
      }
      __syncwarp();
    }
}

// ------------------------------------------------------------------ prefix sum (i32 -> i64 totals)
// out[i] = sum_{j<i} in[j] for i in [0, n]; three launches: block sums, scan of the block sums
// (one block), add. n up to 2^31.
constexpr int K5_SCAN_T = 256, K5_SCAN_E = 8, K5_SCAN_TILE = K5_SCAN_T * K5_SCAN_E;

__global__ void __launch_bounds__(K5_SCAN_T) k5_scan_tile_kernel(const int* __restrict__ in, int64_t n,
                                                                   long long* __restrict__ tile_sum) {
  __shared__ long long red[K5_SCAN_T / 32];
  const int64_t base = (int64_t)blockIdx.x * K5_SCAN_TILE;
  long long s = 0;
  for (int e = 0; e < K5_SCAN_E; ++e) {
    const int64_t i = base + (int64_t)e * K5_SCAN_T + threadIdx.x;
    if (i < n) s += in[i];
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int i = 0; i < K5_SCAN_T / 32; ++i) t += red[i];
    tile_sum[blockIdx.x] = t;
  }
}
__global__ void k5_scan_tops_kernel(long long* __restrict__ tile_sum, int64_t ntiles) {   // one block, serial over tiles
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long run = 0;
    for (int64_t t = 0; t < ntiles; ++t) { const long long x = tile_sum[t]; tile_sum[t] = run; run += x; }
    tile_sum[ntiles] = run;
  }
}
template <typename OutT>
__global__ void __launch_bounds__(K5_SCAN_T) k5_scan_apply_kernel(const int* __restrict__ in, int64_t n,
                                                                    const long long* __restrict__ tile_sum,
                                                                    int64_t ntiles, OutT* __restrict__ out) {
  __shared__ long long wsum[K5_SCAN_T / 32];
  const int64_t base = (int64_t)blockIdx.x * K5_SCAN_TILE + (int64_t)threadIdx.x * K5_SCAN_E;   // thread owns E consecutive items
  long long v[K5_SCAN_E], s = 0;
  for (int e = 0; e < K5_SCAN_E; ++e) { v[e] = (base + e < n) ? in[base + e] : 0; s += v[e]; }
  long long inc = s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) { const long long x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  long long wb = 0;
  for (int w = 0; w < warp; ++w) wb += wsum[w];
  long long run = tile_sum[blockIdx.x] + wb + inc - s;
  for (int e = 0; e < K5_SCAN_E; ++e) { if (base + e < n) out[base + e] = (OutT)run; run += v[e]; }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (OutT)tile_sum[ntiles];
}

// ------------------------------------------------------------------ COO -> CSR
__global__ void k5_coo_count_kernel(const int64_t* __restrict__ row, int64_t nnz, int64_t n, int* __restrict__ cnt,
                                    int* __restrict__ bad) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = row[p];
    if (r < 0 || r >= n) { atomicExch(bad, 1); continue; }
    atomicAdd(cnt + r, 1);
  }
}
// entries of a row land in arbitrary order (atomic cursor); the row sort below orders them by (col, entry id)
__global__ void k5_coo_scatter_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col, int64_t nnz,
                                      int64_t n, const int64_t* __restrict__ ptr, int* __restrict__ cursor,
                                      unsigned long long* __restrict__ keyed, int* __restrict__ bad) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = row[p], c = col[p];
    if (r < 0 || r >= n) continue;
    if (c < 0 || c >= n) { atomicExch(bad, 1); continue; }
    const int at = atomicAdd(cursor + r, 1);
    keyed[ptr[r] + at] = ((unsigned long long)c << 32) | (unsigned long long)(unsigned)p;    // nnz < 2^31
  }
}
// warp per row: sort the row's (col, entry id) keys, then one lane per distinct column sums its run in
// entry order (torch.sparse coalesce() semantics: repeated coordinates are added). Rows longer than
// K5_LIST are sorted by a lane-serial insertion pass over global memory (rare; exact all the same).
__global__ void __launch_bounds__(K5_WARPS * 32)
k5_coo_rowsort_kernel(int64_t n, const int64_t* __restrict__ ptr, unsigned long long* __restrict__ keyed,
                      int* __restrict__ ucnt) {
  __shared__ unsigned long long sk[K5_WARPS][K5_LIST];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * K5_WARPS + w;
  if (i >= n) return;
  const int64_t b = ptr[i];
  const int m = (int)(ptr[i + 1] - b);
  if (m <= K5_LIST) {
    for (int t = lane; t < m; t += 32) sk[w][t] = keyed[b + t];
    __syncwarp();
    k5_sort(sk[w], m, lane);
    int u = 0;
    for (int t = lane; t < m; t += 32) {
      keyed[b + t] = sk[w][t];
      if (t == 0 || (sk[w][t] >> 32) != (sk[w][t - 1] >> 32)) ++u;
    }
    for (int o = 16; o; o >>= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
    if (lane == 0) ucnt[i] = u;
  } else if (lane == 0) {                                   // long row: insertion sort in place
    for (int t = 1; t < m; ++t) {
      const unsigned long long x = keyed[b + t];
      int s = t - 1;
      while (s >= 0 && keyed[b + s] > x) { keyed[b + s + 1] = keyed[b + s]; --s; }
      keyed[b + s + 1] = x;
    }
    int u = 0;
    for (int t = 0; t < m; ++t) if (t == 0 || (keyed[b + t] >> 32) != (keyed[b + t - 1] >> 32)) ++u;
    ucnt[i] = u;
  }
}
__global__ void __launch_bounds__(K5_WARPS * 32)
k5_coo_emit_kernel(int64_t n, const int64_t* __restrict__ ptr, const unsigned long long* __restrict__ keyed,
                   const double* __restrict__ val, const int* __restrict__ optr, int* __restrict__ ocol,
                   double* __restrict__ oval) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * K5_WARPS + w;
  if (i >= n) return;
  const int64_t b = ptr[i];
  const int m = (int)(ptr[i + 1] - b);
  int out = optr[i];
  // lane-parallel over run heads: a head at position t owns [t, next head)
  for (int t0 = 0; t0 < m; t0 += 32) {
    const int t = t0 + lane;
    const bool head = t < m && (t == 0 || (keyed[b + t] >> 32) != (keyed[b + t - 1] >> 32));
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    if (head) {
      const int o = out + __popc(hm & ((1u << lane) - 1u));
      const unsigned c = (unsigned)(keyed[b + t] >> 32);
      double s = 0.0;
      for (int q = t; q < m && (unsigned)(keyed[b + q] >> 32) == c; ++q) s += val[(unsigned)keyed[b + q]];
      ocol[o] = (int)c;
      oval[o] = s;
    }
    out += __popc(hm);
  }
}

// ------------------------------------------------------------------ CSR SpGEMM  C = A * B  (two passes)
// pass 0 (cptr == nullptr): ccnt[i] = nnz(C[i,:]); pass 1: columns ascending + values.
template <bool FILL>
__global__ void __launch_bounds__(K5_WARPS * 32)
k5_spgemm_kernel(int64_t n, const int* __restrict__ aptr, const int* __restrict__ acol, const double* __restrict__ aval,
                 const int* __restrict__ bptr, const int* __restrict__ bcol, const double* __restrict__ bval,
                 int* __restrict__ ccnt, const int* __restrict__ cptr, int* __restrict__ ccol,
                 double* __restrict__ cval, int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char k5_raw[];
  K5Row& r = reinterpret_cast<K5Row*>(k5_raw)[threadIdx.x >> 5];
  __shared__ unsigned long long sk[K5_WARPS][K5_LIST];
  __shared__ int cnt_s[K5_WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * K5_WARPS + w;
  if (i >= n) return;
  k5_reset(r, lane);
  if (lane == 0) cnt_s[w] = 0;
  __syncwarp();
  for (int p = aptr[i]; p < aptr[i + 1]; ++p) {             // source rows in order: fixed summation order
    if (cnt_s[w] > K5_LIST) break;                          // overflow (reported below); the hash must not fill up
    const int k = acol[p];
    const double a = aval ? aval[p] : 1.0;
    for (int q = bptr[k] + lane; q < bptr[k + 1]; q += 32) {
      bool fresh;
      const int s = k5_claim(r, bcol[q], &fresh);
      if (fresh) {
        const int at = atomicAdd(&cnt_s[w], 1);
        if (at < K5_LIST) r.slot_of[at] = s;
      }
      // separate multiply and add (no FMA contraction): bit-identical to scipy's csr_matmat, which matters
      // because the product is pruned of EXACT zeros afterwards. Columns of one B row are distinct: no race.
      if (FILL) r.val[s] = __dadd_rn(r.val[s], __dmul_rn(a, bval ? bval[q] : 1.0));
    }
    __syncwarp();
  }
  const int m = cnt_s[w];
  if (m > K5_LIST) { if (lane == 0) atomicExch(overflow, 1); return; }
  if (!FILL) { if (lane == 0) ccnt[i] = m; return; }
  for (int t = lane; t < m; t += 32) sk[w][t] = ((unsigned long long)(unsigned)r.key[r.slot_of[t]] << 32) | (unsigned)r.slot_of[t];
  __syncwarp();
  k5_sort(sk[w], m, lane);
  const int o = cptr[i];
  for (int t = lane; t < m; t += 32) {
    ccol[o + t] = (int)(sk[w][t] >> 32);
    cval[o + t] = r.val[(unsigned)sk[w][t]];
  }
}

// scipy's csr_matmat (what the reference's `L @ U` runs) stores only results != 0: drop exact zeros.
__global__ void k5_count_nonzero_kernel(int64_t n, const int* __restrict__ ptr, const double* __restrict__ val,
                                        int* __restrict__ cnt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int c = 0;
    for (int p = ptr[i]; p < ptr[i + 1]; ++p) c += val[p] != 0.0;
    cnt[i] = c;
  }
}
__global__ void k5_copy_nonzero_kernel(int64_t n, const int* __restrict__ ptr, const int* __restrict__ col,
                                       const double* __restrict__ val, const int* __restrict__ optr,
                                       int* __restrict__ ocol, double* __restrict__ oval) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int o = optr[i];
    for (int p = ptr[i]; p < ptr[i + 1]; ++p)
      if (val[p] != 0.0) { ocol[o] = col[p]; oval[o] = val[p]; ++o; }
  }
}

// ------------------------------------------------------------------ candidate superset S
// order 0: S(i) = first k of pattern(A^0) U pattern(A^1) U ... U pattern(A^max_power) (i,:) by
//          (graph distance from i, column id); levels are added until k candidates exist.
// order 1: S(i) = first k of pattern(A)(i,:) by (|col - i|, col)   (cfg5: random graph, no distance order)
// Output: scnt[i] (<= k) and scol[i*k ..] ascending by column.
__global__ void __launch_bounds__(K5_WARPS * 32)
k5_superset_kernel(int64_t n, const int* __restrict__ aptr, const int* __restrict__ acol, int k, int max_power,
                   int order, int* __restrict__ scnt, int* __restrict__ scol, int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char k5_raw[];
  K5Row& r = reinterpret_cast<K5Row*>(k5_raw)[threadIdx.x >> 5];
  __shared__ unsigned long long sk[K5_WARPS][K5_LIST];
  __shared__ int cnt_s[K5_WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * K5_WARPS + w;
  if (i >= n) return;
  int m = 0;
  if (order == 1) {
    m = aptr[i + 1] - aptr[i];
    if (m > K5_LIST) { if (lane == 0) atomicExch(overflow, 1); return; }
    for (int t = lane; t < m; t += 32) {
      const int c = acol[aptr[i] + t];
      const unsigned d = (unsigned)(c > (int)i ? c - (int)i : (int)i - c);
      sk[w][t] = ((unsigned long long)d << 32) | (unsigned)c;
    }
  } else {
    k5_reset(r, lane);
    if (lane == 0) {
      bool f;
      r.slot_of[0] = k5_claim(r, (int)i, &f);
      r.aux[0] = 0;
      cnt_s[w] = 1;
    }
    __syncwarp();
    int lo = 0, hi = 1;                                     // frontier = list[lo, hi)
    for (int d = 1; d <= max_power && hi < k && hi > lo; ++d) {
      for (int t = lo; t < hi; ++t) {
        if (cnt_s[w] > K5_LIST) break;
        const int u = r.key[r.slot_of[t]];
        for (int q = aptr[u] + lane; q < aptr[u + 1]; q += 32) {
          bool fresh;
          const int s = k5_claim(r, acol[q], &fresh);
          if (fresh) {
            const int at = atomicAdd(&cnt_s[w], 1);
            if (at < K5_LIST) { r.slot_of[at] = s; r.aux[at] = d; }
          }
        }
        __syncwarp();
      }
      lo = hi;
      hi = cnt_s[w];
      if (hi > K5_LIST) { if (lane == 0) atomicExch(overflow, 1); return; }
    }
    m = hi;
    for (int t = lane; t < m; t += 32)
      sk[w][t] = ((unsigned long long)(unsigned)r.aux[t] << 32) | (unsigned)r.key[r.slot_of[t]];
  }
  __syncwarp();
  k5_sort(sk[w], m, lane);
  const int keep = min(m, k);
  __syncwarp();
  for (int t = lane; t < keep; t += 32) sk[w][t] = sk[w][t] & 0xffffffffull;     // keep the first k, re-sort by column
  __syncwarp();
  k5_sort(sk[w], keep, lane);
  for (int t = lane; t < keep; t += 32) scol[i * k + t] = (int)sk[w][t];
  if (lane == 0) scnt[i] = keep;
}
// [n][k] padded -> row-major COO (row i64, col i64) at sptr
__global__ void k5_pattern_compact_kernel(int64_t n, int k, const int* __restrict__ scnt, const int* __restrict__ scol,
                                          const int64_t* __restrict__ sptr, int64_t* __restrict__ orow,
                                          int64_t* __restrict__ ocol) {
  const int64_t total = n * k;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / k;
    const int t = (int)(idx % k);
    if (t < scnt[i]) { orow[sptr[i] + t] = i; ocol[sptr[i] + t] = scol[idx]; }
  }
}

// ------------------------------------------------------------------ initial values on S
// max_i sum_j |A_ij|  (omega = 1 / that)
__global__ void k5_row_abs_max_kernel(int64_t n, const int* __restrict__ aptr, const double* __restrict__ aval,
                                      unsigned long long* __restrict__ out_bits) {
  double best = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int p = aptr[i]; p < aptr[i + 1]; ++p) s += fabs(aval[p]);
    best = fmax(best, s);
  }
  for (int o = 16; o; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(best));   // non-negative doubles order like their bits
}
// val(i, c) = omega * sum_{j < terms} [(I - omega A)^j]_{ic} for (i, c) in S: row i of the power series is
// carried as a sparse vector x_j (hash of column -> list position, values by list position),
// x_{j+1} = x_j - omega * x_j A with the entries of x_j taken in list order (deterministic sums).
struct K5Vec { double cur[K5_LIST], nxt[K5_LIST], acc[K5_LIST]; };
inline size_t k5_row_smem() { return sizeof(K5Row) * K5_WARPS; }
inline size_t k5_neumann_smem() { return (sizeof(K5Row) + sizeof(K5Vec)) * K5_WARPS; }

__global__ void __launch_bounds__(K5_WARPS * 32)
k5_neumann_kernel(int64_t n, const int* __restrict__ aptr, const int* __restrict__ acol, const double* __restrict__ aval,
                  const int64_t* __restrict__ sptr, const int64_t* __restrict__ scol, int terms, double omega,
                  double* __restrict__ out, int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char k5_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  K5Row& r = reinterpret_cast<K5Row*>(k5_raw)[w];
  K5Vec& v = reinterpret_cast<K5Vec*>(k5_raw + sizeof(K5Row) * K5_WARPS)[w];
  __shared__ int cnt_s[K5_WARPS];
  const int64_t i = (int64_t)blockIdx.x * K5_WARPS + w;
  if (i >= n) return;
  k5_reset(r, lane);
  for (int t = lane; t < K5_LIST; t += 32) { v.cur[t] = 0.0; v.acc[t] = 0.0; }
  __syncwarp();
  if (lane == 0) {
    bool f;
    const int s = k5_claim(r, (int)i, &f);
    r.slot_of[0] = s;
    r.val[s] = 0.0;                                         // r.val[slot] = list position of the column
    v.cur[0] = 1.0;
    v.acc[0] = 1.0;
    cnt_s[w] = 1;
  }
  __syncwarp();
  int m = 1;
  for (int j = 1; j < terms; ++j) {
    for (int t = lane; t < K5_LIST; t += 32) v.nxt[t] = (t < m) ? v.cur[t] : 0.0;
    __syncwarp();
    for (int t = 0; t < m; ++t) {
      const double xu = v.cur[t];
      if (xu == 0.0) continue;                              // warp-uniform
      if (cnt_s[w] > K5_LIST) break;
      const int u = r.key[r.slot_of[t]];
      for (int q = aptr[u] + lane; q < aptr[u + 1]; q += 32) {    // columns of one row of A are distinct
        bool fresh;
        const int s = k5_claim(r, acol[q], &fresh);
        if (fresh) {
          const int at = atomicAdd(&cnt_s[w], 1);
          r.val[s] = (double)at;
          if (at < K5_LIST) r.slot_of[at] = s;
        }
        const int pos = (int)r.val[s];
        if (pos < K5_LIST) v.nxt[pos] -= omega * xu * aval[q];
      }
      __syncwarp();
    }
    m = cnt_s[w];
    if (m > K5_LIST) { if (lane == 0) atomicExch(overflow, 1); return; }
    for (int t = lane; t < m; t += 32) { v.cur[t] = v.nxt[t]; v.acc[t] += v.nxt[t]; }
    __syncwarp();
  }
  for (int64_t p = sptr[i] + lane; p < sptr[i + 1]; p += 32) {
    const int s = k5_find(r, (int)scol[p]);
    out[p] = (s >= 0) ? omega * v.acc[(int)r.val[s]] : 0.0;
  }
}

}  // namespace spai
