// K1 — warp-per-row CSR gather of A(J_i, :) and index-set merge.
//
// For every row i of the candidate pattern S the warp walks the CSR rows
// A[c_e, :] of its candidate columns c_e (coalesced loads of column ids and
// values, one A row per pass), merges them into the sorted union index set I_i
// and emits one 16-byte record per gathered entry, ordered by (output column,
// slot). The merge is rank-based: the position of entry (x, e) is the number of
// gathered entries (x', e') < (x, e), obtained with one binary search per
// candidate row, so it needs no scratch memory and has no size cap. Segment
// heads (first record of every distinct output column) are found with
// __ballot_sync / __popc over the placed records.
//
// Replaces, in the reference: the COO->CSR conversion and row gather inside
// torch.mm (preconditioner.py:88) and the support construction of M@A. Runs
// once per context (the pattern superset and A do not change between batches);
// the reward kernels then select from the gathered tile with the kept-mask.
#pragma once

#include "spai_internal.cuh"

namespace spai {

__global__ void k1_count_kernel(int64_t n, const int32_t* __restrict__ sptr,
                                const int32_t* __restrict__ slot_col,
                                const int32_t* __restrict__ a_ptr, int64_t* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t tot = 0;
  for (int s = sptr[i]; s < sptr[i + 1]; ++s) {
    int c = slot_col[s];
    tot += a_ptr[c + 1] - a_ptr[c];
  }
  counts[i] = tot;
}

// T = float : rec_copy is Rec32[] (w and a in one record), rec_ls unused.
// T = double: rec_copy is Rec64[] holding w, rec_ls is Rec64[] holding a.
template <typename T>
__global__ void __launch_bounds__(256)
k1_fill_kernel(int64_t n, const int32_t* __restrict__ sptr, const int32_t* __restrict__ slot_col,
               const T* __restrict__ slot_val, const int32_t* __restrict__ a_ptr,
               const int32_t* __restrict__ a_col, const T* __restrict__ a_val,
               const int64_t* __restrict__ cptr, int32_t* c_col, void* rec_copy_v, void* rec_ls_v,
               int32_t* __restrict__ r_q, int32_t* __restrict__ r_diag, RowHdr* __restrict__ rhdr) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  Rec32* rc32 = reinterpret_cast<Rec32*>(rec_copy_v);
  Rec64* rc64 = reinterpret_cast<Rec64*>(rec_copy_v);
  Rec64* rl64 = reinterpret_cast<Rec64*>(rec_ls_v);
  const unsigned full = 0xffffffffu;

  for (int64_t i = warp; i < n; i += nwarps) {
    const int sp = sptr[i];
    const int k = sptr[i + 1] - sp;
    const int64_t c0 = cptr[i];
    const int64_t nci = cptr[i + 1] - c0;
    if (nci == 0) {
      if (lane == 0) {
        r_q[i] = 0; r_diag[i] = -1;
        RowHdr h; h.cnt = 0; h.sp = sp; h.k = k; h.flags = 0;
        rhdr[i] = h;
      }
      continue;
    }
    // ---- pass 1: gather A rows, place every entry at its (x, e) rank
    for (int e = 0; e < k; ++e) {
      const int c = slot_col[sp + e];
      const T m = slot_val[sp + e];
      const int beg = a_ptr[c], end = a_ptr[c + 1];
      for (int p = beg + lane; p < end; p += 32) {
        const int x = a_col[p];
        const T av = a_val[p];
        int rank = 0;
        for (int e2 = 0; e2 < k; ++e2) {
          const int c2 = slot_col[sp + e2];
          int lo = a_ptr[c2];
          const int b2 = lo, hi0 = a_ptr[c2 + 1];
          int hi = hi0;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (a_col[mid] < x) lo = mid + 1; else hi = mid;
          }
          rank += lo - b2;
          if (e2 < e && lo < hi0 && a_col[lo] == x) rank += 1;
        }
        const int64_t pos = c0 + rank;
        c_col[pos] = x;
        const uint32_t ebit = (k <= 32) ? (1u << e) : 0u;
        const uint32_t flags = (uint32_t)e;          // segment fields filled in pass 2
        if constexpr (sizeof(T) == 4) {
          Rec32 r;
          r.ebit = ebit; r.flags = flags; r.w = m * av; r.a = av;
          rc32[pos] = r;
        } else {
          Rec64 r;
          r.ebit = ebit; r.flags = flags; r.v = m * av;
          rc64[pos] = r;
          if (rl64) { r.v = av; rl64[pos] = r; }
        }
      }
    }
    __syncwarp();
    // ---- pass 2: segment heads -> output slot index, END / DIAG flags
    int sbase = 0;
    int diag = -1;
    for (int64_t base = 0; base < nci; base += 32) {
      const int64_t p = base + lane;
      const bool valid = p < nci;
      const int x = valid ? __ldcg(c_col + c0 + p) : -1;
      const int prev = (valid && p > 0) ? __ldcg(c_col + c0 + p - 1) : -2;
      const int next = (valid && p + 1 < nci) ? __ldcg(c_col + c0 + p + 1) : -3;
      const bool head = valid && (x != prev);
      const unsigned hb = __ballot_sync(full, head);
      const int s = sbase + __popc(hb & (0xffffffffu >> (31 - lane))) - 1;
      if (valid) {
        uint32_t fl = ((uint32_t)s << 14) | ((next != x) ? F_END : 0u) |
                      ((next != x && next == (int)i) ? F_NEXT_DIAG : 0u);
        if constexpr (sizeof(T) == 4) {
          rc32[c0 + p].flags = __ldcg(&rc32[c0 + p].flags) | fl;
        } else {
          fl |= __ldcg(&rc64[c0 + p].flags);
          rc64[c0 + p].flags = fl;
          if (rl64) rl64[c0 + p].flags = fl;
        }
      }
      const unsigned db = __ballot_sync(full, valid && x == (int)i);
      if (db) diag = __shfl_sync(full, s, __ffs(db) - 1);
      sbase += __popc(hb);
    }
    if (lane == 0) {
      r_q[i] = sbase; r_diag[i] = diag;
      RowHdr h; h.cnt = (int32_t)nci; h.sp = sp; h.k = k; h.flags = (diag == 0) ? 1 : 0;
      rhdr[i] = h;
    }
  }
}

}  // namespace spai
