// libspai_b200.so — C-ABI host side of the B200-native SPAI reward path.
// Declarations and the reference interfaces they replace: include/spai_b200.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <atomic>
#include <chrono>
#include <numeric>
#include <thread>

#include "k0_masks.cuh"
#include "k0b_bucket.cuh"
#include "k0c_cluster.cuh"
#include "k1_plan.cuh"
#include "k2_ls.cuh"
#include "k2g_gram.cuh"
#include "k3_copy.cuh"
#include "k3s_sparse.cuh"
#include "k3t_lut.cuh"
#include "k3m_mma.cuh"
#include "k4_sample.cuh"
#include "spai_internal.cuh"

namespace spai {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct Arena {      // owns device allocations of one object
  std::vector<void*> ptrs;
  int64_t bytes = 0;
  template <typename T> int alloc(T** out, int64_t count) {
    void* p = nullptr;
    const size_t sz = (size_t)std::max<int64_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", sz, cudaGetErrorString(e));
      cudaGetLastError();
      return SPAI_ERR_NOMEM;
    }
    ptrs.push_back(p);
    bytes += (int64_t)sz;
    *out = reinterpret_cast<T*>(p);
    return SPAI_OK;
  }
  template <typename T> int upload(T** out, const std::vector<T>& h) {
    SPAI_TRY(alloc(out, (int64_t)h.size()));
    if (!h.empty()) SPAI_CUDA(cudaMemcpy(*out, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SPAI_OK;
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
    bytes = 0;
  }
  ~Arena() { release(); }
};

// ------------------------------------------------------------------ host side
struct HostPattern {
  std::vector<int32_t> sptr, slot_col, slot_edge, edge_slot, dup_start, dup_len;
  std::vector<float> val32;
  std::vector<double> val64;
};
struct HostCsr {
  std::vector<int32_t> ptr, col;
  std::vector<float> val32;
  std::vector<double> val64;
};

static int sort_coo(int64_t n, int64_t nnz, const int64_t* row, const int64_t* col,
                    std::vector<int64_t>& order, std::vector<int64_t>& key, const char* what) {
  if (nnz >= (int64_t)2147483000LL) { set_error("%s: too many entries (%lld)", what, (long long)nnz); return SPAI_ERR_UNSUPPORTED; }
  key.resize(nnz);
  order.resize(nnz);
  bool sorted = true;
  for (int64_t p = 0; p < nnz; ++p) {
    if (row[p] < 0 || row[p] >= n || col[p] < 0 || col[p] >= n) {
      set_error("%s: entry %lld has index (%lld, %lld) outside [0, %lld)", what, (long long)p,
                (long long)row[p], (long long)col[p], (long long)n);
      return SPAI_ERR_INVALID;
    }
    key[p] = row[p] * n + col[p];
    order[p] = p;
    if (p && key[p] < key[p - 1]) sorted = false;
  }
  if (!sorted)
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key[a] < key[b]; });
  return SPAI_OK;
}

static int build_host_pattern(int64_t n, int64_t E, const int64_t* row, const int64_t* col,
                              const double* val, HostPattern& h, Pattern& P) {
  std::vector<int64_t> order, key;
  SPAI_TRY(sort_coo(n, E, row, col, order, key, "initial matrix"));
  h.sptr.assign(n + 1, 0);
  h.slot_col.resize(E); h.slot_edge.resize(E); h.edge_slot.resize(E);
  h.val32.resize(E); h.val64.resize(E);
  bool ident = true;
  int64_t groups = 0;
  for (int64_t p = 0; p < E; ++p) {
    const int64_t e = order[p];
    if (e != p) ident = false;
    h.sptr[row[e] + 1]++;
    h.slot_col[p] = (int32_t)col[e];
    h.slot_edge[p] = (int32_t)e;
    h.edge_slot[e] = (int32_t)p;
    h.val64[p] = val[e];
    h.val32[p] = (float)val[e];
    if (p == 0 || key[order[p]] != key[order[p - 1]]) {
      ++groups;
    } else {
      if (!h.dup_start.empty() && h.dup_start.back() + h.dup_len.back() == p) h.dup_len.back()++;
      else { h.dup_start.push_back((int32_t)(p - 1)); h.dup_len.push_back(2); }
    }
  }
  int maxk = 0;
  for (int64_t i = 0; i < n; ++i) {
    maxk = std::max(maxk, h.sptr[i + 1]);
    h.sptr[i + 1] += h.sptr[i];
  }
  if (maxk > MAX_ROW_SLOTS) { set_error("a row of the initial matrix has %d entries (max %d)", maxk, MAX_ROW_SLOTS); return SPAI_ERR_UNSUPPORTED; }
  P.n = n; P.E = E; P.init_nnz = groups; P.max_k = maxk; P.identity_perm = ident;
  P.ndup = (int64_t)h.dup_start.size();
  return SPAI_OK;
}

static int upload_pattern(Arena& ar, const HostPattern& h, Pattern& P) {
  SPAI_TRY(ar.upload(&P.sptr, h.sptr));
  SPAI_TRY(ar.upload(&P.slot_col, h.slot_col));
  SPAI_TRY(ar.upload(&P.slot_edge, h.slot_edge));
  SPAI_TRY(ar.upload(&P.edge_slot, h.edge_slot));
  SPAI_TRY(ar.upload(&P.slot_val32, h.val32));
  SPAI_TRY(ar.upload(&P.slot_val64, h.val64));
  SPAI_TRY(ar.upload(&P.dup_start, h.dup_start));
  SPAI_TRY(ar.upload(&P.dup_len, h.dup_len));
  return SPAI_OK;
}

static int build_host_csr(int64_t n, int64_t nnz, const int64_t* row, const int64_t* col,
                          const double* val, HostCsr& h) {
  std::vector<int64_t> order, key;
  SPAI_TRY(sort_coo(n, nnz, row, col, order, key, "original matrix"));
  h.ptr.assign(n + 1, 0);
  h.col.clear(); h.val32.clear(); h.val64.clear();
  h.col.reserve(nnz); h.val32.reserve(nnz); h.val64.reserve(nnz);
  for (int64_t p = 0; p < nnz; ++p) {
    const int64_t e = order[p];
    if (p && key[order[p]] == key[order[p - 1]]) {        // coalesce: sum in stored order
      h.val64.back() += val[e];
      h.val32.back() += (float)val[e];
    } else {
      h.ptr[row[e] + 1]++;
      h.col.push_back((int32_t)col[e]);
      h.val64.push_back(val[e]);
      h.val32.push_back((float)val[e]);
    }
  }
  for (int64_t i = 0; i < n; ++i) h.ptr[i + 1] += h.ptr[i];
  return SPAI_OK;
}

static int upload_csr(Arena& ar, const HostCsr& h, int64_t n, int64_t stored, CsrA& A) {
  A.n = n; A.nnz = (int64_t)h.col.size(); A.nnz_stored = stored;
  SPAI_TRY(ar.upload(&A.ptr, h.ptr));
  SPAI_TRY(ar.upload(&A.col, h.col));
  SPAI_TRY(ar.upload(&A.val32, h.val32));
  SPAI_TRY(ar.upload(&A.val64, h.val64));
  return SPAI_OK;
}

// ls register-kernel classes.
//   kind 0: row-distributed tile (k2_ls_kernel<T,KMAX,G,QL>), rows up to G*QL
//   kind 1: column-per-lane tile (k2c_ls_kernel<T,W,QMAX>), W = KMAX lanes per problem
struct LsClass { int kind, kmax, g, ql; };
static const LsClass kLsClasses[LS_NCLASS - 1] = {
    {0, 8, 2, 9}, {0, 8, 4, 5}, {0, 8, 8, 5}, {1, 16, 16, 52}, {1, 16, 16, 64},
    {1, 32, 32, 52}, {1, 32, 32, 64}};
static inline int ls_class_rows_max(const LsClass& L) { return L.kind == 0 ? L.g * L.ql : L.ql; }
static inline int ls_class_lanes(const LsClass& L) { return L.g; }

static int classify_row(int k, int q, bool has_dup) {
  for (int c = 0; c < LS_NCLASS - 1; ++c) {
    const LsClass& L = kLsClasses[c];
    if (L.kmax == 0) continue;
    if (has_dup) continue;                        // register kernels assume full column rank
    if (k <= L.kmax && q <= ls_class_rows_max(L)) return c;
  }
  return LS_GENERIC;
}

// K1 driver: count -> scan (host) -> fill -> classify/tiling (host)
static int build_plan(Arena& ar, const Pattern& P, const HostPattern& hp, const CsrA& A,
                      const HostCsr& ha, int dtype, bool want_ls, Plan& plan, cudaStream_t st) {
  const int64_t n = P.n;
  plan.dtype = dtype; plan.n = n;
  int64_t* counts = nullptr;
  Arena tmp;
  SPAI_TRY(tmp.alloc(&counts, n));
  k1_count_kernel<<<(unsigned)ceil_div(std::max<int64_t>(n, 1), 256), 256, 0, st>>>(n, P.sptr, P.slot_col, A.ptr, counts);
  SPAI_CUDA(cudaGetLastError());
  std::vector<int64_t> hc(n);
  SPAI_CUDA(cudaMemcpyAsync(hc.data(), counts, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));
  plan.cptr_host.assign(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    if (hc[i] >= (int64_t)1 << 30) { set_error("row %lld gathers %lld entries (unsupported)", (long long)i, (long long)hc[i]); return SPAI_ERR_UNSUPPORTED; }
    plan.cptr_host[i + 1] = plan.cptr_host[i] + hc[i];
  }
  plan.nc = plan.cptr_host[n];
  SPAI_TRY(ar.upload(&plan.cptr, plan.cptr_host));
  SPAI_TRY(ar.alloc(&plan.c_col, plan.nc));
  SPAI_TRY(ar.alloc(&plan.r_q, n));
  SPAI_TRY(ar.alloc(&plan.r_diag, n));
  SPAI_TRY(ar.alloc(&plan.rhdr, n));
  if (dtype == SPAI_F32) {
    Rec32* r = nullptr;
    SPAI_TRY(ar.alloc(&r, plan.nc + 1));          // +1: the copy kernel prefetches one record ahead
    plan.rec_copy = r; plan.rec_ls = r;
  } else {
    Rec64* r = nullptr;
    SPAI_TRY(ar.alloc(&r, plan.nc + 1));
    plan.rec_copy = r;
    if (want_ls) { Rec64* l = nullptr; SPAI_TRY(ar.alloc(&l, plan.nc)); plan.rec_ls = l; }
  }
  if (n > 0) {
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 8), 148 * 64);
    if (dtype == SPAI_F32)
      k1_fill_kernel<float><<<blocks, 256, 0, st>>>(n, P.sptr, P.slot_col, P.slot_val32, A.ptr, A.col, A.val32,
                                                   plan.cptr, plan.c_col, plan.rec_copy, nullptr, plan.r_q, plan.r_diag, plan.rhdr);
    else
      k1_fill_kernel<double><<<blocks, 256, 0, st>>>(n, P.sptr, P.slot_col, P.slot_val64, A.ptr, A.col, A.val64,
                                                    plan.cptr, plan.c_col, plan.rec_copy, plan.rec_ls, plan.r_q, plan.r_diag, plan.rhdr);
    SPAI_CUDA(cudaGetLastError());
  }
  if (dtype == SPAI_F32) {
    float* rb = nullptr;
    SPAI_TRY(ar.alloc(&rb, n));
    k3_row_base_kernel<float><<<(unsigned)ceil_div(std::max<int64_t>(n, 1), 256), 256, 0, st>>>(
        reinterpret_cast<const Rec32*>(plan.rec_copy), plan.cptr, plan.rhdr, n, rb);
    plan.row_base = rb;
  } else {
    double* rb = nullptr;
    SPAI_TRY(ar.alloc(&rb, n));
    k3_row_base_kernel<double><<<(unsigned)ceil_div(std::max<int64_t>(n, 1), 256), 256, 0, st>>>(
        reinterpret_cast<const Rec64*>(plan.rec_copy), plan.cptr, plan.rhdr, n, rb);
    plan.row_base = rb;
  }
  SPAI_CUDA(cudaGetLastError());
  std::vector<int32_t> hq(n), hd(n);
  SPAI_CUDA(cudaMemcpyAsync(hq.data(), plan.r_q, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaMemcpyAsync(hd.data(), plan.r_diag, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));

  // tiles of <= K3_TILE_C records (an oversized row is a tile of its own)
  std::vector<int32_t> tiles;
  tiles.push_back(0);
  int64_t acc = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t c = hc[i];
    if (acc > 0 && acc + c > K3_TILE_C) { tiles.push_back((int32_t)i); acc = 0; }
    acc += c;
    if (i + 1 - tiles.back() >= K3_TILE_R) { tiles.push_back((int32_t)(i + 1)); acc = 0; }
  }
  if (tiles.back() != n) tiles.push_back((int32_t)n);
  plan.ntiles = (int)tiles.size() - 1;
  plan.tile_row_host = tiles;
  SPAI_TRY(ar.upload(&plan.tile_row, tiles));

  std::vector<int32_t> cls[LS_NCLASS];
  std::vector<char> row_dup(n, 0);
  for (size_t g = 0; g < hp.dup_start.size(); ++g) {
    const int32_t s0 = hp.dup_start[g];
    const int64_t row = std::upper_bound(hp.sptr.begin(), hp.sptr.end(), s0) - hp.sptr.begin() - 1;
    row_dup[row] = 1;
  }
  plan.rows_missing_diag = 0; plan.max_q = 0; plan.max_k = P.max_k;
  plan.generic_max_q = plan.generic_max_k = 0;
  const double w = (dtype == SPAI_F32) ? 4.0 : 8.0;
  double g = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const int k = hp.sptr[i + 1] - hp.sptr[i];
    const int q = hq[i];
    if (q >= MAX_ROW_UNION) { set_error("row %lld has a union index set of %d columns (max %d)", (long long)i, q, MAX_ROW_UNION - 1); return SPAI_ERR_UNSUPPORTED; }
    if (hd[i] < 0) plan.rows_missing_diag++;
    plan.max_q = std::max(plan.max_q, q);
    g += k * (4.0 + w + 8.0) + (double)hc[i] * (4.0 + w);
    if (q == 0) continue;
    const int c = classify_row(k, q, row_dup[i] != 0);
    cls[c].push_back((int32_t)i);
    if (c == LS_GENERIC) {
      plan.generic_max_q = std::max<int64_t>(plan.generic_max_q, q);
      plan.generic_max_k = std::max<int64_t>(plan.generic_max_k, k);
    }
  }
  plan.g_bytes_full = g;
  plan.missing_prefix.assign(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) plan.missing_prefix[i + 1] = plan.missing_prefix[i] + (hd[i] < 0 ? 1 : 0);
  for (int c = 0; c < LS_NCLASS; ++c) {
    plan.class_rows_host[c] = cls[c];
    plan.class_count[c] = (int64_t)cls[c].size();
    if (want_ls && !cls[c].empty()) SPAI_TRY(ar.upload(&plan.class_rows[c], cls[c]));
  }
  (void)ha;
  if ((want_ls || dtype == SPAI_F32) && plan.rec_ls) {
    // all-kept ls residual of every row (incremental path of the ls kernels): the generic
    // solver in per-row mode over all rows with a non-empty union, one all-ones trajectory
    SPAI_TRY(ar.alloc(&plan.row_base_ls, n));
    SPAI_CUDA(cudaMemsetAsync(plan.row_base_ls, 0, (size_t)std::max<int64_t>(n, 1) * 8, st));
    std::vector<int32_t> live;
    for (int64_t i = 0; i < n; ++i) if (hq[i] > 0) live.push_back((int32_t)i);
    if (!live.empty()) {
      Arena t2;
      int32_t* live_dev = nullptr;
      SPAI_TRY(t2.upload(&live_dev, live));
      const int64_t W = std::max<int64_t>(P.words(), 1), Bp = 32;
      uint32_t* ones = nullptr;
      SPAI_TRY(t2.alloc(&ones, W * Bp));
      SPAI_CUDA(cudaMemsetAsync(ones, 0xff, (size_t)W * Bp * 4, st));
      const int64_t warps = 148 * 16;
      const int64_t wstride = std::max<int64_t>(plan.max_q, 1) * ((int64_t)P.max_k + 1);
      int32_t* cmap = nullptr;
      SPAI_TRY(t2.alloc(&cmap, warps * std::max<int64_t>(P.max_k, 1)));
      if (dtype == SPAI_F32) {
        float* work = nullptr;
        SPAI_TRY(t2.alloc(&work, warps * wstride));
        k2_ls_generic_kernel<float><<<(unsigned)(warps / 4), 128, 0, st>>>(
            reinterpret_cast<const Rec32*>(plan.rec_ls), plan.cptr, P.sptr, plan.r_q, plan.r_diag, live_dev,
            (int64_t)live.size(), ones, Bp, 1, work, wstride, cmap, std::max<int64_t>(P.max_k, 1), nullptr, nullptr,
            nullptr, 0u, plan.row_base_ls);
      } else {
        double* work = nullptr;
        SPAI_TRY(t2.alloc(&work, warps * wstride));
        k2_ls_generic_kernel<double><<<(unsigned)(warps / 4), 128, 0, st>>>(
            reinterpret_cast<const Rec64*>(plan.rec_ls), plan.cptr, P.sptr, plan.r_q, plan.r_diag, live_dev,
            (int64_t)live.size(), ones, Bp, 1, work, wstride, cmap, std::max<int64_t>(P.max_k, 1), nullptr, nullptr,
            nullptr, 0u, plan.row_base_ls);
      }
      SPAI_CUDA(cudaGetLastError());
      SPAI_CUDA(cudaStreamSynchronize(st));      // scratch of t2 is released on return
    }
  }
  plan.bytes = ar.bytes;
  return SPAI_OK;
}

// ------------------------------------------------------------------ evaluation
struct Workspace {
  void* base = nullptr;
  int64_t bytes = 0;
  int ensure(int64_t need) {
    if (need <= bytes) return SPAI_OK;
    if (base) cudaFree(base);
    base = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&base, (size_t)need);
    if (e != cudaSuccess) { set_error("workspace cudaMalloc(%lld) failed: %s", (long long)need, cudaGetErrorString(e)); cudaGetLastError(); return SPAI_ERR_NOMEM; }
    bytes = need;
    return SPAI_OK;
  }
  ~Workspace() { if (base) cudaFree(base); }
};

struct Carver {
  char* p; char* end;
  template <typename T> T* take(int64_t count) {
    uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255);
    T* out = reinterpret_cast<T*>(a);
    p = reinterpret_cast<char*>(a) + count * (int64_t)sizeof(T);
    return out;
  }
};
static inline int64_t padded(int64_t bytes) { return round_up(bytes, 256) + 256; }

constexpr unsigned int LS_FAIL_CAP = 1u << 18;   // (row, trajectory) tiles redone by the generic kernel
constexpr int K3S_MAX_CHUNKS = 512;
constexpr int64_t K3S_MIN_RATIO = 40;            // K3s when a trajectory deletes <= 1/40 of the candidates (measured crossover ~3 % on cfg2)

struct EvalShape {       // launch geometry of one reward evaluation over Bc trajectories
  int64_t Bp = 32;                                  // padded trajectory count (columns of maskT)
  int nt = 1, gx = 1, gy = 1;                       // copy kernel
  int ls_gx[LS_NCLASS] = {}, ls_gy[LS_NCLASS] = {}, ls_ntg[LS_NCLASS] = {};
  int gram_gx[GRAM_NCLASS] = {}, gram_gy[GRAM_NCLASS] = {};   // semi-normal-equation classes (ls_gram mode)
  int parts = 1;
  int64_t generic_work = 0, generic_cmap = 0, generic_warps = 0;
  bool has_column_class = false;      // column-per-lane kernels may hand tiles to the generic kernel
};

// resident CTAs per SM of the copy kernel instantiation (occupancy API, cached):
// the grid is sized to ONE full wave so no half-empty second wave trails.
static int k3_blocks_per_sm(int dtype, int nt, bool compact = false) {
  static int cache[2][2][9] = {};
  int& c = cache[compact][dtype == SPAI_F64][nt];
  if (c) return c;
  int v = 0;
  const size_t smem = (size_t)(dtype == SPAI_F32 ? k3_smem_bytes<float>(compact) : k3_smem_bytes<double>(compact));
#define SPAI_OCC(T, NT)                                                                                          \
  (compact ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k3_copy_kernel<T, NT, true>, K3_THREADS, smem)    \
           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k3_copy_kernel<T, NT, false>, K3_THREADS, smem))
  if (dtype == SPAI_F32) {
    if (nt == 8) SPAI_OCC(float, 8); else if (nt == 4) SPAI_OCC(float, 4);
    else if (nt == 2) SPAI_OCC(float, 2); else SPAI_OCC(float, 1);
  } else {
    if (nt == 4) SPAI_OCC(double, 4); else if (nt == 2) SPAI_OCC(double, 2); else SPAI_OCC(double, 1);
  }
#undef SPAI_OCC
  if (cudaGetLastError() != cudaSuccess || v <= 0) v = 4;
  c = v;
  return c;
}

// Gram classes (ls_gram mode): which kernel solves rows with k <= kmax in which precision.
//   lanes = 1: k2g_solve_kernel<T,K> (packed triangle); lanes > 1: k2gd_solve_kernel<T,K,G> (full matrix)
struct GramVariant { int kmax, lanes; int64_t rb; int smem; };
static GramVariant gram_variant(int dtype, int g) {
  if (dtype == SPAI_F32) {
    if (g == 0) return {8, 1, GramGeom<float, 8, false>::RB, GramGeom<float, 8, false>::STAGES};
    if (g == 1) return {16, 1, GramGeom<float, 16, false>::RB, GramGeom<float, 16, false>::STAGES};
    return {32, 4, GramGeom<float, 32, true>::RB, GramDist<float, 32, 4>::SMEM};
  }
  if (g == 0) return {8, 1, GramGeom<double, 8, false>::RB, GramGeom<double, 8, false>::STAGES};
  // lanes per problem: measured on B200 (B = 256): k <= 16 fp64 2 lanes 13.7 ms vs 4 lanes 17.5 ms (cfg3);
  // k <= 32 fp64 8 lanes 106 ms vs 16 lanes 152 ms, fp32 4 lanes 52.7 ms vs 8 lanes 56.1 ms (cfg4)
  if (g == 1) return {16, 2, GramGeom<double, 16, true>::RB, GramDist<double, 16, 2>::SMEM};
  return {32, 8, GramGeom<double, 32, true>::RB, GramDist<double, 32, 8>::SMEM};
}

// ls_gram mode through the (row, kept-mask) table (K3t) instead of the Gram class-0 kernel
static inline bool ls_table_on(const Plan& p) {
  if (!p.lut_ls_ready || !p.tables_on) return false;
  const char* v = getenv("SPAI_K3_LUT");            // A/B switch shared with copy mode
  return !v || atoi(v) != 0;
}

// ls mode through the Householder-filled table (covers every row when rows have <= 8 candidates)
static inline bool qr_table_on(const Plan& p) {
  if (!p.lut_qr_ready || !p.tables_on) return false;
  const char* v = getenv("SPAI_K3_LUT");
  return !v || atoi(v) != 0;
}

// row list of QR class c: in ls_gram mode only the rows no Gram class takes
static inline int64_t ls_count(const Plan& p, int mode, int c) {
  return mode == SPAI_MODE_LS_GRAM ? p.rest_count[c] : p.class_count[c];
}
static inline const int32_t* ls_rows(const Plan& p, int mode, int c) {
  return mode == SPAI_MODE_LS_GRAM ? p.rest_rows[c] : p.class_rows[c];
}
static inline const std::vector<int32_t>& ls_rows_host(const Plan& p, int mode, int c) {
  return mode == SPAI_MODE_LS_GRAM ? p.rest_rows_host[c] : p.class_rows_host[c];
}

static EvalShape plan_shape(const Plan& plan, int mode, int dtype, int64_t Bc, int sm_count) {
  EvalShape s;
  if (mode == SPAI_MODE_COPY) {
    s.nt = (Bc >= 1024) ? 8 : (Bc >= 512 ? 4 : (Bc >= 256 ? 2 : 1));
    if (dtype == SPAI_F64 && s.nt > 4) s.nt = 4;
    s.Bp = round_up(Bc, (int64_t)K3_THREADS * s.nt);
    const int64_t Bp = s.Bp;
    s.gy = (int)ceil_div(Bp, (int64_t)K3_THREADS * s.nt);
    const int target = sm_count * k3_blocks_per_sm(dtype, s.nt);     // one full wave
    s.gx = (int)std::max<int64_t>(1, std::min<int64_t>(plan.ntiles, std::max(1, target / s.gy)));
    s.parts = std::max(s.gx, K3S_MAX_CHUNKS);        // the deletion-driven kernel writes one partial row per word chunk
  } else {
    s.Bp = round_up(Bc, 32);
    const bool ls_table = mode == SPAI_MODE_LS_GRAM && ls_table_on(plan);
    if (mode == SPAI_MODE_LS && qr_table_on(plan)) {          // one lookup kernel, nothing else
      s.nt = (Bc >= 512) ? 4 : (Bc >= 256 ? 2 : 1);
      s.Bp = round_up(Bc, (int64_t)K3_THREADS * s.nt);
      s.gy = (int)ceil_div(s.Bp, (int64_t)K3_THREADS * s.nt);
      s.gx = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(plan.n, 16), std::max(1, sm_count * 6 / s.gy)));
      s.parts = s.gx;
      return s;
    }
    if (ls_table) {                                  // the lookup kernel reads 128*NT mask columns per block
      s.nt = (Bc >= 1024) ? 8 : (Bc >= 512 ? 4 : (Bc >= 256 ? 2 : 1));
      if (dtype == SPAI_F64 && s.nt > 4) s.nt = 4;
      s.Bp = round_up(Bc, (int64_t)K3_THREADS * s.nt);
    }
    const int64_t Bp = s.Bp;
    s.parts = 0;
    if (mode == SPAI_MODE_LS_GRAM) {
      for (int g = 0; g < GRAM_NCLASS; ++g) {
        if (!plan.gram_count[g]) continue;
        if (g == 0 && ls_table) {
          s.has_column_class = true;
          s.gram_gy[0] = (int)ceil_div(Bp, (int64_t)K3_THREADS * s.nt);
          s.gram_gx[0] = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(plan.n, 16), std::max(1, sm_count * 6 / s.gram_gy[0])));
          s.parts += s.gram_gx[0];
          continue;
        }
        const GramVariant v = gram_variant(dtype, g);
        s.has_column_class = true;                   // ill-conditioned tiles go to the generic kernel
        s.gram_gy[g] = (int)ceil_div(Bp, (int64_t)(K3_THREADS / v.lanes));
        const int target = sm_count * (g == 0 ? 4 : 2) * 4;
        s.gram_gx[g] = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(plan.gram_count[g], 16),
                                                                   std::max(1, target / s.gram_gy[g])));
        s.parts += s.gram_gx[g];
      }
    }
    for (int c = 0; c < LS_NCLASS - 1; ++c) {
      if (!ls_count(plan, mode, c)) continue;
      const LsClass& L = kLsClasses[c];
      const int groups = K2_NW * (32 / ls_class_lanes(L));
      s.has_column_class = true;      // every register kernel assumes full rank and may hand tiles over
      s.ls_ntg[c] = (int)std::min<int64_t>(K2_MAX_NTG, ceil_div(Bp, groups));
      s.ls_gy[c] = (int)ceil_div(Bp, (int64_t)groups * s.ls_ntg[c]);
      const int target = sm_count * 12;
      s.ls_gx[c] = (int)std::max<int64_t>(1, std::min<int64_t>(ls_count(plan, mode, c), std::max(1, target / s.ls_gy[c])));
      s.parts += s.ls_gx[c];
    }
    if (ls_count(plan, mode, LS_GENERIC) || s.has_column_class) {
      s.generic_warps = (int64_t)sm_count * 16;
      const int64_t gq = s.has_column_class ? plan.max_q : plan.generic_max_q;
      const int64_t gk = s.has_column_class ? plan.max_k : plan.generic_max_k;
      s.generic_work = std::max<int64_t>(gq, 1) * (gk + 1);
      s.generic_cmap = std::max<int64_t>(gk, 1);
    }
    if (s.parts == 0) s.parts = 1;
  }
  return s;
}

static int64_t eval_bytes(const Plan& plan, const EvalShape& s, int64_t W, int dtype) {
  const int64_t Bp = s.Bp;
  int64_t b = padded(W * Bp * 4)              // maskT
              + padded(Bp * 8)                 // nnz
              + padded((int64_t)s.parts * Bp * 8)
              + padded(Bp * 8);                // res2 extra
  if (s.generic_warps) {
    b += padded(s.generic_warps * s.generic_work * 8);
    b += padded(s.generic_warps * s.generic_cmap * 4);
    b += padded((int64_t)LS_FAIL_CAP * 8) + padded(64);
  }
  (void)plan;
  return b;
}

template <typename T>
static int launch_ls_class(int c, const int32_t* crows, const Plan& plan, const Pattern& P, const EvalShape& s,
                           const uint32_t* maskT, int64_t Bp, int64_t Bc, double* partial, int2* fail_pairs,
                           unsigned int* fail_count, cudaStream_t st, int64_t roff, int64_t rcnt) {
  using Rec = typename RecOf<T>::type;
  const Rec* recs = reinterpret_cast<const Rec*>(plan.rec_ls);
  const dim3 grid(s.ls_gx[c], s.ls_gy[c]);
  const dim3 block(K2_NW * 32);
#define SPAI_LS_ROW(IDX, KMAX, G, QL)                                                          \
  case IDX:                                                                                    \
    k2_ls_kernel<T, KMAX, G, QL><<<grid, block, 0, st>>>(recs, plan.cptr, P.sptr, plan.r_diag,  \
        crows + roff, rcnt, maskT, Bp, s.ls_ntg[c], partial,                                    \
        plan.row_base_ls, Bc, fail_pairs, fail_count, LS_FAIL_CAP);                            \
    break;
#define SPAI_LS_COL(IDX, W, QMAX)                                                              \
  case IDX:                                                                                    \
    k2c_ls_kernel<T, W, QMAX><<<grid, block, 0, st>>>(recs, plan.cptr, P.sptr, plan.r_diag,     \
        crows + roff, rcnt, maskT, Bp, Bc, s.ls_ntg[c], partial, fail_pairs,                    \
        fail_count, LS_FAIL_CAP, plan.row_base_ls);                                            \
    break;
  switch (c) {
    SPAI_LS_ROW(0, 8, 2, 9)
    SPAI_LS_ROW(1, 8, 4, 5)
    SPAI_LS_ROW(2, 8, 8, 5)
    SPAI_LS_COL(3, 16, 52)
    SPAI_LS_COL(4, 16, 64)
    SPAI_LS_COL(5, 32, 52)
    SPAI_LS_COL(6, 32, 64)
    default: set_error("bad ls class %d", c); return SPAI_ERR_INVALID;
  }
#undef SPAI_LS_ROW
#undef SPAI_LS_COL
  SPAI_CUDA(cudaGetLastError());
  return SPAI_OK;
}

struct PhaseTimer {
  bool on = false;
  cudaEvent_t ev[5] = {};
  int init() {
    for (auto& e : ev) SPAI_CUDA(cudaEventCreate(&e));
    return SPAI_OK;
  }
  void destroy() { for (auto& e : ev) if (e) cudaEventDestroy(e); }
};

// Reward of Bc trajectories whose slot-order masks mask[Bc][W] are already built.
// `scratch` has at least eval_bytes(); outputs are device pointers (may be null).
static inline bool k3m_on() {
  const char* e = getenv("SPAI_K3_MMA");
  return !e || atoi(e) != 0;
}
static int eval_masks(const Pattern& P, const Plan& plan, int mode, int dtype, const uint32_t* mask,
                      int64_t Bc, char* scratch, int64_t scratch_bytes, int sm_count, double n_d,
                      double res0, double flops0, double alpha, double* reward, double* residual,
                      int64_t* nnz_out, cudaStream_t st, PhaseTimer* pt, int* launches,
                      const long long* nnz_ready = nullptr, int64_t row_lo = 0, int64_t row_hi = -1,
                      bool partial_only = false, int64_t t_hint = 0) {
  const int64_t W = P.words();
  if (row_hi < 0 || row_hi > plan.n) row_hi = plan.n;
  if (row_lo < 0) row_lo = 0;
  if (row_lo > row_hi) row_lo = row_hi;
  // sub-range of a sorted class row list
  auto sub = [&](int c, int64_t* off, int64_t* cnt) {
    const auto& v = c >= 100 ? plan.gram_rows_host[c - 100] : ls_rows_host(plan, mode, c);
    const int64_t a = std::lower_bound(v.begin(), v.end(), (int32_t)row_lo) - v.begin();
    const int64_t b = std::lower_bound(v.begin(), v.end(), (int32_t)std::min<int64_t>(row_hi, INT32_MAX)) - v.begin();
    *off = a; *cnt = b - a;
  };
  const EvalShape s = plan_shape(plan, mode, dtype, Bc, sm_count);
  const int64_t Bp = s.Bp;
  Carver cv{scratch, scratch + scratch_bytes};
  uint32_t* maskT = cv.take<uint32_t>(W * Bp);
  long long* nnz = cv.take<long long>(Bp);
  double* partial = cv.take<double>((int64_t)s.parts * Bp);
  double* res2x = cv.take<double>(Bp);
  int nl = 0;

  // few deletions per trajectory: the deletion-driven kernel (K3s) works on the trajectory-major
  // mask directly, no transposed copy needed
  bool sparse = mode == SPAI_MODE_COPY && plan.sparse_ready && t_hint > 0 && t_hint * K3S_MIN_RATIO <= P.E && W > 0;
  if (const char* v = getenv("SPAI_K3_SPARSE")) sparse = atoi(v) != 0 && mode == SPAI_MODE_COPY && plan.sparse_ready && W > 0;   // A/B switch
  bool use_lut = mode == SPAI_MODE_COPY && !sparse && plan.lut_ready && plan.tables_on && W > 0;
  if (const char* v = getenv("SPAI_K3_LUT")) use_lut = use_lut && atoi(v) != 0;               // A/B switch
  if (W > 0 && !sparse) {
    static const bool t32 = [] { const char* e = getenv("SPAI_K0_TRANSPOSE"); return e && atoi(e) == 32; }();    // A/B: the 32 x 32 version
    if (!t32 && (W & 3) == 0 && (Bp & 63) == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0 && (reinterpret_cast<uintptr_t>(maskT) & 15) == 0) {
      const dim3 tg((unsigned)ceil_div(W, 64), (unsigned)(Bp / 64));
      k0_transpose64_kernel<<<tg, 256, 0, st>>>(mask, Bc, W, maskT, Bp);
    } else {
      const dim3 tg((unsigned)ceil_div(W, 32), (unsigned)(Bp / 32));
      k0_transpose_kernel<<<tg, 256, 0, st>>>(mask, Bc, W, maskT, Bp);
    }
    SPAI_CUDA(cudaGetLastError()); ++nl;
  }
  // K3s reads every mask word anyway: over the full row range it also counts the kept slots
  const bool count_in_k3s = sparse && !nnz_ready && row_lo == 0 && row_hi == plan.n;
  if (nnz_ready) {
    SPAI_CUDA(cudaMemcpyAsync(nnz, nnz_ready, (size_t)Bc * 8, cudaMemcpyDeviceToDevice, st));   // popcount fused into K0
  } else if (count_in_k3s) {
    SPAI_CUDA(cudaMemsetAsync(nnz, 0, (size_t)Bp * 8, st));
  } else {
    k0_popcount_kernel<<<(unsigned)Bc, 256, 0, st>>>(mask, W, Bc, nnz);
    SPAI_CUDA(cudaGetLastError()); ++nl;
  }
  if (P.ndup) {
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(P.ndup * Bc, 256), 65535);
    k0_dup_correction_kernel<<<blocks, 256, 0, st>>>(mask, W, Bc, P.dup_start, P.dup_len, P.ndup, nnz);
    SPAI_CUDA(cudaGetLastError()); ++nl;
  }
  if (pt && pt->on) cudaEventRecord(pt->ev[2], st);

  bool use_extra = false;
  const unsigned int* fail_count_dev = nullptr;
  int parts = s.parts;
  if (sparse) {
    const int64_t w_lo = plan.sptr_host[row_lo] >> 5;
    const int64_t w_hi = std::min<int64_t>(W, ((int64_t)plan.sptr_host[row_hi] + 31) >> 5);
    const int64_t range = std::max<int64_t>(w_hi - w_lo, 1);
    // ~4096 mask words (131 072 slots) per block: the slot metadata and (f, w) pairs of the range
    // (~13 MB) stay in L2 while the batch sweeps it. Measured on cfg5 (W = 340 384 words, B = 4096):
    // 4 chunks 32.9 ms, 16: 25.0, 64: 19.5, 128: 20.5, 512: 23.8; cfg2 (W = 16 384): 4 chunks best.
    const int64_t want_chunks = std::max<int64_t>(1, std::min<int64_t>(K3S_MAX_CHUNKS, ceil_div(range, 4096)));
    const int64_t chunk_words = round_up(ceil_div(range, want_chunks), 128);
    const int chunks = (int)ceil_div(range, chunk_words);
    const double base_sum = plan.base_prefix[row_hi] - plan.base_prefix[row_lo];
    const dim3 grid((unsigned)Bc, (unsigned)chunks);
    const int slot_lo = plan.sptr_host[row_lo], slot_hi = plan.sptr_host[row_hi];
    const SlotMeta* meta = reinterpret_cast<const SlotMeta*>(plan.sl_meta);
    if (dtype == SPAI_F32)
      k3s_sparse_kernel<float><<<grid, K3S_THREADS, 0, st>>>(
          meta, P.slot_col, plan.a_ptr, plan.a_col, reinterpret_cast<const Pair<float>*>(plan.sl_rec), mask, W,
          Bc, Bp, w_lo, w_hi, chunk_words, slot_lo, slot_hi, base_sum, partial, count_in_k3s ? nnz : nullptr);
    else
      k3s_sparse_kernel<double><<<grid, K3S_THREADS, 0, st>>>(
          meta, P.slot_col, plan.a_ptr, plan.a_col, reinterpret_cast<const Pair<double>*>(plan.sl_rec), mask, W,
          Bc, Bp, w_lo, w_hi, chunk_words, slot_lo, slot_hi, base_sum, partial, count_in_k3s ? nnz : nullptr);
    SPAI_CUDA(cudaGetLastError()); ++nl;
    parts = chunks;
  } else if (mode == SPAI_MODE_COPY && use_lut) {
    // one full wave of blocks over (row ranges) x (128*NT trajectories)
    const int64_t rows = row_hi - row_lo;
    int gx = 1;
    if (rows > 0) {
      const int per_sm = std::max(1, (int)(200 * 1024 / (dtype == SPAI_F32 ? k3t_smem_bytes<float>() : k3t_smem_bytes<double>())));
      gx = (int)std::max<int64_t>(1, std::min<int64_t>({ceil_div(rows, 16), (int64_t)std::max(1, sm_count * std::min(per_sm, 16) / s.gy),
                                                       (int64_t)s.parts}));
    }
    const dim3 grid(gx, s.gy);
#define SPAI_K3T(T, NT)                                                                                   \
  k3t_lookup_kernel<T, NT><<<grid, K3_THREADS, k3t_smem_bytes<T>(), st>>>(reinterpret_cast<const T*>(plan.lut), \
      plan.rhdr, maskT, Bp, W, partial, (int)row_lo, (int)row_hi)
    if (rows <= 0) {
      SPAI_CUDA(cudaMemsetAsync(partial, 0, (size_t)Bp * 8, st));
    } else if (dtype == SPAI_F32) {
      if (s.nt == 8) SPAI_K3T(float, 8); else if (s.nt == 4) SPAI_K3T(float, 4);
      else if (s.nt == 2) SPAI_K3T(float, 2); else SPAI_K3T(float, 1);
    } else {
      if (s.nt == 4) SPAI_K3T(double, 4); else if (s.nt == 2) SPAI_K3T(double, 2); else SPAI_K3T(double, 1);
    }
#undef SPAI_K3T
    SPAI_CUDA(cudaGetLastError()); ++nl;
    parts = gx;
  } else if (mode == SPAI_MODE_COPY && dtype == SPAI_F32 && plan.mma_ready && plan.tables_on && k3m_on() && Bc >= 64) {
    // tensor-core row residuals (K3m): per row class (<= 16 candidates: N = 16; 17..32: N = 32) one wave of CTAs over
    // (ranges of the class's row list) x (128 * NTM trajectories)
    int off = 0;
    for (int cl = 0; cl < 2; ++cl) {
      const auto& rl = plan.mma_rows_host[cl];
      const int64_t i_lo = std::lower_bound(rl.begin(), rl.end(), (int32_t)row_lo) - rl.begin();
      const int64_t i_hi = std::lower_bound(rl.begin(), rl.end(), (int32_t)std::min<int64_t>(row_hi, INT32_MAX)) - rl.begin();
      const int64_t rows = i_hi - i_lo;
      if (rows <= 0) continue;
      const int n_cl = cl ? 32 : 16;
      int ntm = std::min(s.nt, 4);
      if (const char* v = getenv("SPAI_K3M_NTM")) ntm = std::max(1, std::min(ntm, atoi(v)));      // A/B switch
      if (ntm == 3) ntm = 2;
      if (n_cl == 32 && ntm > 2) ntm = 2;                     // TMEM columns / registers at N = 32
      const int gy = (int)(Bp / ((int64_t)128 * ntm));
      const void* fn = nullptr;
      size_t smem = 0;
      int tcols = 512;
#define SPAI_K3M_PICK(N_, S_, NTM_)                                                      \
  do { fn = (const void*)k3m_kernel<N_, S_, NTM_>; smem = (size_t)k3m_smem_bytes<N_, S_, NTM_>(); tcols = k3m_tmem_cols<N_, NTM_>(); } while (0)
#define SPAI_K3M_CLASS(N_)                                                               \
  do {                                                                                   \
    if (plan.mma_split == 2) { if (ntm == 2) SPAI_K3M_PICK(N_, 2, 2); else SPAI_K3M_PICK(N_, 2, 1); } \
    else { if (ntm == 2) SPAI_K3M_PICK(N_, 3, 2); else SPAI_K3M_PICK(N_, 3, 1); }                       \
  } while (0)
      if (n_cl == 16 && ntm >= 4) {
        if (plan.mma_split == 2) SPAI_K3M_PICK(16, 2, 4); else SPAI_K3M_PICK(16, 3, 4);
      } else if (n_cl == 16) SPAI_K3M_CLASS(16);
      else SPAI_K3M_CLASS(32);
#undef SPAI_K3M_CLASS
#undef SPAI_K3M_PICK
      SPAI_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SPAI_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
      int per_sm = 1;
      {                                           // resident CTAs from the kernel's own resources (the occupancy API answered 1
        cudaFuncAttributes fa;                    // for a 288-thread, 63 KB, 70-register kernel on this driver)
        if (cudaFuncGetAttributes(&fa, fn) == cudaSuccess && fa.numRegs > 0) {
          const int by_regs = 65536 / (((fa.numRegs + 7) / 8 * 8) * ((K3M_THREADS + 31) / 32 * 32));
          const int by_smem = (int)((227 * 1024) / (smem + 1024));
          per_sm = std::max(1, std::min({by_regs, by_smem, 2048 / K3M_THREADS}));
        } else cudaGetLastError();
      }
      per_sm = std::max(1, std::min(per_sm, 512 / tcols));                        // TMEM columns per SM
      const int room = std::max(1, s.parts - off);
      const int gx = (int)std::max<int64_t>(1, std::min<int64_t>({ceil_div(rows, 32), (int64_t)std::max(1, sm_count * per_sm / gy), (int64_t)room}));
      if (getenv("SPAI_K3M_VERBOSE"))
        fprintf(stderr, "[k3m] class N=%d rows=%lld split=%d ntm=%d smem=%zu tmem_cols=%d ctas/SM=%d grid=(%d,%d)\n", n_cl, (long long)rows,
                plan.mma_split, ntm, smem, tcols, per_sm, gx, gy);
      const unsigned char* recp = plan.mma_rec[cl];
      const int2* rh = plan.mma_hdr[cl];
      int rlo = (int)i_lo, rhi = (int)i_hi;
      int64_t bp_ = Bp, w_ = W;
      const uint32_t* mt = maskT;
      double* pp = partial + (int64_t)off * Bp;
      int dbg = 0;
      if (const char* v = getenv("SPAI_K3M_DEBUG")) dbg = atoi(v);                // timing experiments only (results are wrong)
      void* args[] = {(void*)&recp, (void*)&rh, (void*)&mt, (void*)&bp_, (void*)&w_, (void*)&pp, (void*)&rlo, (void*)&rhi, (void*)&dbg};
      SPAI_CUDA(cudaLaunchKernel(fn, dim3(gx, gy), dim3(K3M_THREADS), args, smem, st));
      SPAI_CUDA(cudaGetLastError()); ++nl;
      off += gx;
    }
    if (off == 0) { SPAI_CUDA(cudaMemsetAsync(partial, 0, (size_t)Bp * 8, st)); off = 1; }
    parts = off;
  } else if (mode == SPAI_MODE_COPY) {
    // tiles that intersect [row_lo, row_hi)
    const auto& tr = plan.tile_row_host;
    int t_lo = 0, t_n = plan.ntiles;
    if (row_lo > 0 || row_hi < plan.n) {
      t_lo = (int)(std::upper_bound(tr.begin(), tr.end(), (int32_t)row_lo) - tr.begin()) - 1;
      int t_hi = (int)(std::lower_bound(tr.begin(), tr.end(), (int32_t)row_hi) - tr.begin());
      t_lo = std::max(0, std::min(t_lo, plan.ntiles));
      t_hi = std::max(t_lo, std::min(t_hi, plan.ntiles));
      t_n = (row_hi > row_lo) ? t_hi - t_lo : 0;
    }
    // compaction pays when few of the 32*NT trajectories of a warp touch a given row: expected
    // touching trajectories per (warp, row) ~ 32*NT * deletions/n (hint = longest trajectory; 0 = unknown)
    bool compact = t_hint <= 0 || (double)t_hint * 32.0 * s.nt <= 128.0 * (double)plan.n;
    if (const char* v = getenv("SPAI_K3_COMPACT")) compact = atoi(v) != 0;      // A/B switch
    const size_t smem = (size_t)(dtype == SPAI_F32 ? k3_smem_bytes<float>(compact) : k3_smem_bytes<double>(compact));
    // the workspace holds s.gx partial rows (sized for the smaller-footprint variant); the
    // compact variant keeps fewer CTAs resident, so its one-wave grid is narrower
    int gx = s.gx;
    if (compact) gx = std::max(1, std::min(gx, sm_count * k3_blocks_per_sm(dtype, s.nt, true) / s.gy));
    parts = gx;
    const dim3 grid(gx, s.gy);
#define SPAI_K3(T, NT)                                                                          \
  do {                                                                                          \
    const typename RecOf<T>::type* rc_ = reinterpret_cast<const typename RecOf<T>::type*>(plan.rec_copy); \
    const T* rb_ = reinterpret_cast<const T*>(plan.row_base);                                   \
    if (compact)                                                                                \
      k3_copy_kernel<T, NT, true><<<grid, K3_THREADS, smem, st>>>(rc_, plan.cptr, plan.rhdr, rb_, \
          plan.tile_row + t_lo, t_n, maskT, Bp, W, partial, (int)row_lo, (int)row_hi);          \
    else                                                                                        \
      k3_copy_kernel<T, NT, false><<<grid, K3_THREADS, smem, st>>>(rc_, plan.cptr, plan.rhdr, rb_, \
          plan.tile_row + t_lo, t_n, maskT, Bp, W, partial, (int)row_lo, (int)row_hi);          \
  } while (0)
    if (t_n == 0) {
      SPAI_CUDA(cudaMemsetAsync(partial, 0, (size_t)Bp * 8, st)); parts = 1;
    } else if (dtype == SPAI_F32) {
      if (s.nt == 8) SPAI_K3(float, 8); else if (s.nt == 4) SPAI_K3(float, 4);
      else if (s.nt == 2) SPAI_K3(float, 2); else SPAI_K3(float, 1);
    } else {
      if (s.nt == 4) SPAI_K3(double, 4); else if (s.nt == 2) SPAI_K3(double, 2); else SPAI_K3(double, 1);
    }
#undef SPAI_K3
    SPAI_CUDA(cudaGetLastError()); ++nl;
  } else if (mode == SPAI_MODE_LS && qr_table_on(plan)) {
    const dim3 tgrid(s.gx, s.gy);
    if (row_hi <= row_lo) {
      SPAI_CUDA(cudaMemsetAsync(partial, 0, (size_t)Bp * 8, st));
      parts = 1;
    } else {
      if (s.nt == 4) k3t_lookup_kernel<double, 4><<<tgrid, K3_THREADS, k3t_smem_bytes<double>(), st>>>(plan.lut_qr, plan.rhdr, maskT, Bp, W, partial, (int)row_lo, (int)row_hi);
      else if (s.nt == 2) k3t_lookup_kernel<double, 2><<<tgrid, K3_THREADS, k3t_smem_bytes<double>(), st>>>(plan.lut_qr, plan.rhdr, maskT, Bp, W, partial, (int)row_lo, (int)row_hi);
      else k3t_lookup_kernel<double, 1><<<tgrid, K3_THREADS, k3t_smem_bytes<double>(), st>>>(plan.lut_qr, plan.rhdr, maskT, Bp, W, partial, (int)row_lo, (int)row_hi);
      SPAI_CUDA(cudaGetLastError()); ++nl;
      parts = s.gx;
    }
  } else {
    if (!plan.rec_ls) { set_error("plan was built without ls records"); return SPAI_ERR_INVALID; }
    int2* fail_pairs = nullptr;
    unsigned int* fail_count = nullptr;
    void* gwork = nullptr;
    int32_t* cmap = nullptr;
    if (s.generic_warps) {
      gwork = cv.take<double>(s.generic_warps * s.generic_work);       // sized for fp64, also holds fp32
      cmap = cv.take<int32_t>(s.generic_warps * s.generic_cmap);
      fail_pairs = cv.take<int2>(LS_FAIL_CAP);
      fail_count = cv.take<unsigned int>(16);
      SPAI_CUDA(cudaMemsetAsync(fail_count, 0, 64, st));
      SPAI_CUDA(cudaMemsetAsync(res2x, 0, (size_t)Bp * 8, st));
      use_extra = true;
      fail_count_dev = fail_count;
    }
    int off = 0;
    bool any = false;
    if (mode == SPAI_MODE_LS_GRAM) {
      if (!plan.gram_ready) { set_error("plan was built without Gram records"); return SPAI_ERR_INVALID; }
      for (int g = 0; g < GRAM_NCLASS; ++g) {
        if (!plan.gram_count[g]) continue;
        double* pp = partial + (int64_t)off * Bp;
        if (g == 0 && ls_table_on(plan)) {            // every (row, mask) residual is tabulated: look it up
          const dim3 tgrid(s.gram_gx[0], s.gram_gy[0]);
#define SPAI_K3TL(T, NT)                                                                                  \
  k3t_lookup_kernel<T, NT, true><<<tgrid, K3_THREADS, k3t_smem_bytes<T>(), st>>>(reinterpret_cast<const T*>(plan.lut_ls), \
      plan.rhdr, maskT, Bp, W, pp, (int)row_lo, (int)row_hi, Bc, fail_pairs, fail_count, LS_FAIL_CAP)
          if (row_hi <= row_lo) {
            SPAI_CUDA(cudaMemsetAsync(pp, 0, (size_t)s.gram_gx[0] * Bp * 8, st));
          } else if (dtype == SPAI_F32) {
            if (s.nt == 8) SPAI_K3TL(float, 8); else if (s.nt == 4) SPAI_K3TL(float, 4);
            else if (s.nt == 2) SPAI_K3TL(float, 2); else SPAI_K3TL(float, 1);
          } else {
            if (s.nt == 4) SPAI_K3TL(double, 4); else if (s.nt == 2) SPAI_K3TL(double, 2); else SPAI_K3TL(double, 1);
          }
#undef SPAI_K3TL
          SPAI_CUDA(cudaGetLastError());
          off += s.gram_gx[0]; ++nl; any = true;
          continue;
        }
        int64_t roff, rcnt;
        sub(100 + g, &roff, &rcnt);
        const dim3 grid(s.gram_gx[g], s.gram_gy[g]);
        const GramVariant gv = gram_variant(dtype, g);
        const unsigned char* gp = plan.gram[g] + roff * gv.rb;
#define SPAI_GRAM(T, K)                                                                          \
  k2g_solve_kernel<T, K><<<grid, K3_THREADS, gv.smem, st>>>(gp, rcnt, maskT, Bp, W, Bc, pp, fail_pairs, fail_count, LS_FAIL_CAP)
#define SPAI_GRAMD(T, K, G)                                                                      \
  k2gd_solve_kernel<T, K, G><<<grid, K3_THREADS, gv.smem, st>>>(gp, rcnt, maskT, Bp, W, Bc, pp, fail_pairs, fail_count, LS_FAIL_CAP)
        if (dtype == SPAI_F32) {
          if (g == 0) SPAI_GRAM(float, 8); else if (g == 1) SPAI_GRAM(float, 16);
          else SPAI_GRAMD(float, 32, 4);
        } else if (g == 0) SPAI_GRAM(double, 8);
        else if (g == 1) SPAI_GRAMD(double, 16, 2);
        else SPAI_GRAMD(double, 32, 8);
#undef SPAI_GRAM
#undef SPAI_GRAMD
        SPAI_CUDA(cudaGetLastError());
        off += s.gram_gx[g]; ++nl; any = true;
      }
    }
    for (int c = 0; c < LS_NCLASS - 1; ++c) {
      if (!ls_count(plan, mode, c)) continue;
      double* pp = partial + (int64_t)off * Bp;
      int64_t roff, rcnt;
      sub(c, &roff, &rcnt);
      const int32_t* crows = ls_rows(plan, mode, c);
      if (dtype == SPAI_F32)
        SPAI_TRY(launch_ls_class<float>(c, crows, plan, P, s, maskT, Bp, Bc, pp, fail_pairs, fail_count, st, roff, rcnt));
      else
        SPAI_TRY(launch_ls_class<double>(c, crows, plan, P, s, maskT, Bp, Bc, pp, fail_pairs, fail_count, st, roff, rcnt));
      off += s.ls_gx[c]; ++nl; any = true;
    }
    if (!any) { SPAI_CUDA(cudaMemsetAsync(partial, 0, (size_t)Bp * 8, st)); parts = 1; }
    else parts = off;
    if (s.generic_warps) {
      const unsigned blocks = (unsigned)(s.generic_warps / 4);
      for (int pass = 0; pass < 2; ++pass) {
        // pass 0: the rows classified as generic; pass 1: tiles handed over by the column kernels
        if (pass == 0 && !ls_count(plan, mode, LS_GENERIC)) continue;
        if (pass == 1 && !s.has_column_class) continue;
        const int2* pairs = pass ? fail_pairs : nullptr;
        int64_t goff = 0, gcnt = ls_count(plan, mode, LS_GENERIC);
        if (!pass) sub(LS_GENERIC, &goff, &gcnt);
        const int32_t* grows = ls_rows(plan, mode, LS_GENERIC) ? ls_rows(plan, mode, LS_GENERIC) + goff : nullptr;
        if (dtype == SPAI_F32)
          k2_ls_generic_kernel<float><<<blocks, 128, 0, st>>>(
              reinterpret_cast<const Rec32*>(plan.rec_ls), plan.cptr, P.sptr, plan.r_q, plan.r_diag,
              grows, gcnt, maskT, Bp, Bc,
              reinterpret_cast<float*>(gwork), s.generic_work, cmap, s.generic_cmap, res2x, pairs, fail_count,
              LS_FAIL_CAP, nullptr);
        else
          k2_ls_generic_kernel<double><<<blocks, 128, 0, st>>>(
              reinterpret_cast<const Rec64*>(plan.rec_ls), plan.cptr, P.sptr, plan.r_q, plan.r_diag,
              grows, gcnt, maskT, Bp, Bc,
              reinterpret_cast<double*>(gwork), s.generic_work, cmap, s.generic_cmap, res2x, pairs, fail_count,
              LS_FAIL_CAP, nullptr);
        SPAI_CUDA(cudaGetLastError()); ++nl;
      }
    }
  }
  if (pt && pt->on) cudaEventRecord(pt->ev[3], st);
  k3_finalize_kernel<<<(unsigned)ceil_div(Bc, 256), 256, 0, st>>>(
      partial, parts, Bp, Bc, use_extra ? res2x : nullptr,
      (double)(plan.missing_prefix.empty() ? plan.rows_missing_diag
                                           : plan.missing_prefix[row_hi] - plan.missing_prefix[row_lo]),
      nnz, n_d, res0, flops0, alpha, reward, residual, reinterpret_cast<long long*>(nnz_out), fail_count_dev,
      LS_FAIL_CAP, partial_only ? 1 : 0);
  SPAI_CUDA(cudaGetLastError()); ++nl;
  if (launches) *launches += nl;
  return SPAI_OK;
}

}  // namespace spai

// =================================================================== context
using namespace spai;

struct spai_ctx {
  int device = 0;
  int sm_count = 148;
  int64_t n = 0;
  Arena arena;            // pattern + CSR
  Arena plan_arena[2];
  Pattern P;
  CsrA A;
  HostPattern hp;
  HostCsr ha;
  Plan plan[2];
  bool plan_ready[2] = {false, false};
  bool plan_has_ls[2] = {false, false};
  double res0[2] = {0, 0};
  int64_t flops0 = 0;
  int64_t ws_limit = (int64_t)16 << 30;
  int64_t deletion_hint = 0;            // for the taken-bitmask entry point (no action list to measure)
  Workspace ws;
  PhaseTimer pt;
  spai_timing last = {};
  // one-time table/plan builds are enqueued on the stream of the call that needs them first;
  // `build_ev` orders every later call (any stream) after the last build
  cudaEvent_t build_ev = nullptr;
  bool build_pending = false;
  int64_t build_bytes_seen = 0;
  int wait_builds(cudaStream_t st) {
    if (build_pending) SPAI_CUDA(cudaStreamWaitEvent(st, build_ev, 0));
    return SPAI_OK;
  }
  int mark_builds(cudaStream_t st) {          // called after the ensure_* steps of a call
    const int64_t now = plan_arena[0].bytes + plan_arena[1].bytes;
    if (now != build_bytes_seen) {
      if (!build_ev) SPAI_CUDA(cudaEventCreateWithFlags(&build_ev, cudaEventDisableTiming));
      SPAI_CUDA(cudaEventRecord(build_ev, st));
      build_pending = true;
      build_bytes_seen = now;
    }
    return SPAI_OK;
  }
};

namespace spai {

static int ensure_plan(spai_ctx* c, int dtype, bool want_ls, cudaStream_t st) {
  if (c->plan_ready[dtype] && (!want_ls || c->plan_has_ls[dtype])) return SPAI_OK;
  c->plan_arena[dtype].release();
  c->plan[dtype] = Plan();
  // fp32 records always carry `a`; fp64 ls records are a second array
  SPAI_TRY(build_plan(c->plan_arena[dtype], c->P, c->hp, c->A, c->ha, dtype, want_ls || dtype == SPAI_F32,
                      c->plan[dtype], st));
  SPAI_CUDA(cudaStreamSynchronize(st));            // one-time build (K1): plans are shared by every later call and stream
  c->plan_ready[dtype] = true;
  c->plan_has_ls[dtype] = want_ls || dtype == SPAI_F32;
  return SPAI_OK;
}

// ls_gram mode: split the rows with a non-empty union into the Gram classes and the rest
// (which keep their QR class), then form every Gram row once (k2g_build_kernel).
static int ensure_gram(spai_ctx* c, int dtype, cudaStream_t st) {
  Plan& plan = c->plan[dtype];
  if (plan.gram_ready) return SPAI_OK;
  if (!plan.rec_ls || !plan.row_base_ls) { set_error("ls_gram: plan has no ls records"); return SPAI_ERR_INVALID; }
  Arena& ar = c->plan_arena[dtype];
  const HostPattern& hp = c->hp;
  std::vector<char> row_dup(plan.n, 0);
  for (size_t g = 0; g < hp.dup_start.size(); ++g) {
    const int32_t s0 = hp.dup_start[g];
    row_dup[std::upper_bound(hp.sptr.begin(), hp.sptr.end(), s0) - hp.sptr.begin() - 1] = 1;
  }
  const int klimit = 32;
  std::vector<int32_t> gr[GRAM_NCLASS];
  for (int cl = 0; cl < LS_NCLASS; ++cl) {
    plan.rest_rows_host[cl].clear();
    for (int32_t i : plan.class_rows_host[cl]) {
      const int k = hp.sptr[i + 1] - hp.sptr[i];
      if (!row_dup[i] && k <= klimit) gr[k <= 8 ? 0 : (k <= 16 ? 1 : 2)].push_back(i);
      else plan.rest_rows_host[cl].push_back(i);
    }
    plan.rest_count[cl] = (int64_t)plan.rest_rows_host[cl].size();
    plan.rest_rows[cl] = nullptr;
    if (plan.rest_count[cl]) SPAI_TRY(ar.upload(&plan.rest_rows[cl], plan.rest_rows_host[cl]));
  }
  for (int g = 0; g < GRAM_NCLASS; ++g) {
    std::sort(gr[g].begin(), gr[g].end());
    plan.gram_rows_host[g] = gr[g];
    plan.gram_count[g] = (int64_t)gr[g].size();
    plan.gram[g] = nullptr;
    if (gr[g].empty()) continue;
    SPAI_TRY(ar.alloc(&plan.gram[g], plan.gram_count[g] * gram_variant(dtype, g).rb));
    Arena tmp;
    int32_t* rows_dev = nullptr;
    SPAI_TRY(tmp.upload(&rows_dev, gr[g]));
    const unsigned blocks = (unsigned)ceil_div(plan.gram_count[g], 4);
#define SPAI_GB(T, K, FULL)                                                                             \
  k2g_build_kernel<T, K, FULL><<<blocks, 128, 0, st>>>(reinterpret_cast<const typename RecOf<T>::type*>(plan.rec_ls), \
      plan.cptr, c->P.sptr, plan.r_diag, plan.row_base_ls, rows_dev, plan.gram_count[g], plan.gram[g])
    if (dtype == SPAI_F32) {
      if (g == 0) SPAI_GB(float, 8, false); else if (g == 1) SPAI_GB(float, 16, false); else SPAI_GB(float, 32, true);
    } else if (g == 0) SPAI_GB(double, 8, false);
    else if (g == 1) SPAI_GB(double, 16, true);
    else SPAI_GB(double, 32, true);
#undef SPAI_GB
    SPAI_CUDA(cudaGetLastError());
    SPAI_CUDA(cudaStreamSynchronize(st));          // rows_dev is released on scope exit
  }
  plan.bytes = ar.bytes;
  plan.gram_ready = true;
  return SPAI_OK;
}

// K3t table: every (row, kept-mask) residual of a pattern whose rows have <= 8 candidates.
static int ensure_lut(spai_ctx* c, int dtype, cudaStream_t st) {
  Plan& plan = c->plan[dtype];
  if (plan.lut_ready) return SPAI_OK;
  const int64_t n = c->P.n;
  if (n == 0 || c->P.max_k > K3T_K) return SPAI_OK;            // not applicable: the row sweep stays
  if (n * K3T_ENTRIES * (dtype == SPAI_F32 ? 4 : 8) > ((int64_t)4 << 30)) return SPAI_OK;   // table cap: 4 GiB
  Arena& ar = c->plan_arena[dtype];
  if (dtype == SPAI_F32) {
    float* t = nullptr;
    SPAI_TRY(ar.alloc(&t, n * K3T_ENTRIES));
    k3t_build_kernel<float><<<(unsigned)n, K3T_ENTRIES, 0, st>>>(reinterpret_cast<const Rec32*>(plan.rec_copy), plan.cptr, plan.rhdr, t);
    plan.lut = t;
  } else {
    double* t = nullptr;
    SPAI_TRY(ar.alloc(&t, n * K3T_ENTRIES));
    k3t_build_kernel<double><<<(unsigned)n, K3T_ENTRIES, 0, st>>>(reinterpret_cast<const Rec64*>(plan.rec_copy), plan.cptr, plan.rhdr, t);
    plan.lut = t;
  }
  SPAI_CUDA(cudaGetLastError());
  SPAI_CUDA(cudaStreamSynchronize(st));            // one-time build: a later call on another stream sees a finished table
  plan.bytes = ar.bytes;
  plan.lut_ready = true;
  return SPAI_OK;
}

// K3m records: per-row Cholesky factor of the Gram matrix of the row's candidate contributions, bf16-split,
// in the UMMA canonical layout (k3m_mma.cuh). fp32 copy mode only.
static bool k3m_enabled() {                  // A/B + test switch: SPAI_K3_MMA=0 keeps the row sweep (K3)
  const char* e = getenv("SPAI_K3_MMA");
  return !e || atoi(e) != 0;
}
static int k3m_split() {
  const char* e = getenv("SPAI_K3M_SPLIT");
  const int v = e ? atoi(e) : 3;
  return v == 2 ? 2 : 3;
}
static int ensure_mma(spai_ctx* c, int dtype, cudaStream_t st) {
  Plan& plan = c->plan[dtype];
  if (plan.mma_ready || plan.mma_unavailable || dtype != SPAI_F32) return SPAI_OK;
  const int64_t n = c->P.n;
  const int split = k3m_split();
  if (n == 0 || c->P.ndup || c->P.max_k > 32 || !plan.c_col || !plan.rec_copy || (int64_t)c->hp.sptr.size() != n + 1) { plan.mma_unavailable = true; return SPAI_OK; }
  std::vector<int32_t> rows[2];
  for (int64_t i = 0; i < n; ++i) rows[(c->hp.sptr[i + 1] - c->hp.sptr[i]) > 16 ? 1 : 0].push_back((int32_t)i);
  const int64_t rb[2] = {split == 2 ? K3mGeom<16, 2>::RB : K3mGeom<16, 3>::RB, split == 2 ? K3mGeom<32, 2>::RB : K3mGeom<32, 3>::RB};
  if ((int64_t)rows[0].size() * rb[0] + (int64_t)rows[1].size() * rb[1] > ((int64_t)24 << 30)) { plan.mma_unavailable = true; return SPAI_OK; }
  Arena& ar = c->plan_arena[dtype];
  const Rec32* rc = reinterpret_cast<const Rec32*>(plan.rec_copy);
  for (int cl = 0; cl < 2; ++cl) {
    const int64_t cnt = (int64_t)rows[cl].size();
    plan.mma_count[cl] = cnt;
    if (!cnt) continue;
    unsigned char* rec = nullptr;
    int2* hdr = nullptr;
    int32_t* rdev = nullptr;
    SPAI_TRY(ar.alloc(&rec, cnt * rb[cl]));
    SPAI_TRY(ar.alloc(&hdr, cnt));
    SPAI_TRY(ar.alloc(&rdev, cnt));
    SPAI_CUDA(cudaMemcpyAsync(rdev, rows[cl].data(), (size_t)cnt * 4, cudaMemcpyHostToDevice, st));
    const unsigned blocks = (unsigned)ceil_div(cnt, 4);
    if (cl == 0 && split == 2) k3m_build_kernel<16, 2><<<blocks, 128, 0, st>>>(rc, plan.cptr, plan.c_col, plan.rhdr, rdev, cnt, rec, hdr);
    else if (cl == 0) k3m_build_kernel<16, 3><<<blocks, 128, 0, st>>>(rc, plan.cptr, plan.c_col, plan.rhdr, rdev, cnt, rec, hdr);
    else if (split == 2) k3m_build_kernel<32, 2><<<blocks, 128, 0, st>>>(rc, plan.cptr, plan.c_col, plan.rhdr, rdev, cnt, rec, hdr);
    else k3m_build_kernel<32, 3><<<blocks, 128, 0, st>>>(rc, plan.cptr, plan.c_col, plan.rhdr, rdev, cnt, rec, hdr);
    SPAI_CUDA(cudaGetLastError());
    SPAI_CUDA(cudaStreamSynchronize(st));          // one-time build (and rows[] is a host temporary)
    plan.mma_rec[cl] = rec;
    plan.mma_hdr[cl] = hdr;
    plan.mma_rows_host[cl] = std::move(rows[cl]);
  }
  plan.mma_split = split;
  plan.bytes = ar.bytes;
  plan.mma_ready = true;
  return SPAI_OK;
}

static int ensure_lut_ls(spai_ctx* c, int dtype, cudaStream_t st) {
  Plan& plan = c->plan[dtype];
  if (plan.lut_ls_ready || !plan.gram_ready) return SPAI_OK;
  const int64_t n = c->P.n;
  if (n == 0 || c->P.max_k > K3T_K || !plan.gram_count[0]) return SPAI_OK;
  if (n * K3T_ENTRIES * (dtype == SPAI_F32 ? 4 : 8) > ((int64_t)4 << 30)) return SPAI_OK;
  Arena& ar = c->plan_arena[dtype];
  const size_t w = dtype == SPAI_F32 ? 4 : 8;
  unsigned char* t = nullptr;
  SPAI_TRY(ar.alloc(&t, n * K3T_ENTRIES * (int64_t)w));
  SPAI_CUDA(cudaMemsetAsync(t, 0, (size_t)n * K3T_ENTRIES * w, st));     // rows outside Gram class 0 add nothing here
  if (dtype == SPAI_F32)
    k3t_build_ls_kernel<float><<<(unsigned)plan.gram_count[0], K3T_ENTRIES, 0, st>>>(plan.gram[0], reinterpret_cast<float*>(t));
  else
    k3t_build_ls_kernel<double><<<(unsigned)plan.gram_count[0], K3T_ENTRIES, 0, st>>>(plan.gram[0], reinterpret_cast<double*>(t));
  SPAI_CUDA(cudaGetLastError());
  SPAI_CUDA(cudaStreamSynchronize(st));            // one-time build: a later call on another stream sees a finished table
  plan.lut_ls = t;
  plan.bytes = ar.bytes;
  plan.lut_ls_ready = true;
  return SPAI_OK;
}

// ls (Householder) mode table: the generic QR kernel evaluated on the 256 uniform masks.
static int ensure_lut_qr(spai_ctx* c, int dtype, cudaStream_t st) {
  Plan& plan = c->plan[dtype];
  if (plan.lut_qr_ready || !plan.rec_ls) return SPAI_OK;
  const int64_t n = c->P.n;
  if (n == 0 || c->P.max_k > K3T_K || n * K3T_ENTRIES * 8 > ((int64_t)4 << 30)) return SPAI_OK;
  Arena& ar = c->plan_arena[dtype];
  SPAI_TRY(ar.alloc(&plan.lut_qr, n * K3T_ENTRIES));
  SPAI_CUDA(cudaMemsetAsync(plan.lut_qr, 0, (size_t)n * K3T_ENTRIES * 8, st));
  std::vector<int32_t> live;
  for (int cl = 0; cl < LS_NCLASS; ++cl) live.insert(live.end(), plan.class_rows_host[cl].begin(), plan.class_rows_host[cl].end());
  if (!live.empty()) {
    Arena tmp;
    int32_t* live_dev = nullptr;
    SPAI_TRY(tmp.upload(&live_dev, live));
    const int64_t W = std::max<int64_t>(c->P.words(), 1), Bp = K3T_ENTRIES;
    uint32_t* masks = nullptr;
    SPAI_TRY(tmp.alloc(&masks, W * Bp));
    SPAI_CUDA(cudaMemsetAsync(masks, 0, (size_t)W * Bp * 4, st));
    k3t_uniform_masks_kernel<<<(unsigned)ceil_div(n * K3T_ENTRIES, 256), 256, 0, st>>>(c->P.sptr, n, masks);
    SPAI_CUDA(cudaGetLastError());
    const int64_t warps = (int64_t)c->sm_count * 16;
    const int64_t wstride = std::max<int64_t>(plan.max_q, 1) * ((int64_t)c->P.max_k + 1);
    int32_t* cmap = nullptr;
    SPAI_TRY(tmp.alloc(&cmap, warps * std::max<int64_t>(c->P.max_k, 1)));
    if (dtype == SPAI_F32) {
      float* work = nullptr;
      SPAI_TRY(tmp.alloc(&work, warps * wstride));
      k2_ls_generic_kernel<float><<<(unsigned)(warps / 4), 128, 0, st>>>(
          reinterpret_cast<const Rec32*>(plan.rec_ls), plan.cptr, c->P.sptr, plan.r_q, plan.r_diag, live_dev,
          (int64_t)live.size(), masks, Bp, K3T_ENTRIES, work, wstride, cmap, std::max<int64_t>(c->P.max_k, 1), nullptr,
          nullptr, nullptr, 0u, plan.lut_qr, K3T_ENTRIES);
    } else {
      double* work = nullptr;
      SPAI_TRY(tmp.alloc(&work, warps * wstride));
      k2_ls_generic_kernel<double><<<(unsigned)(warps / 4), 128, 0, st>>>(
          reinterpret_cast<const Rec64*>(plan.rec_ls), plan.cptr, c->P.sptr, plan.r_q, plan.r_diag, live_dev,
          (int64_t)live.size(), masks, Bp, K3T_ENTRIES, work, wstride, cmap, std::max<int64_t>(c->P.max_k, 1), nullptr,
          nullptr, nullptr, 0u, plan.lut_qr, K3T_ENTRIES);
    }
    SPAI_CUDA(cudaGetLastError());
    SPAI_CUDA(cudaStreamSynchronize(st));            // scratch of tmp is released on return
  }
  plan.bytes = ar.bytes;
  plan.lut_qr_ready = true;
  return SPAI_OK;
}

// K3s data: slot -> row map, slot-major (segment sum, contribution) pairs, prefix sums of the
// all-kept row residuals. Built the first time a batch of short trajectories is scored in copy mode.
static int ensure_sparse(spai_ctx* c, int dtype, cudaStream_t st) {
  Plan& plan = c->plan[dtype];
  if (plan.sparse_ready || plan.sparse_unavailable) return SPAI_OK;
  const HostPattern& hp = c->hp;
  const int64_t n = c->P.n, E = c->P.E;
  Arena& ar = c->plan_arena[dtype];
  std::vector<SlotMeta> meta((size_t)std::max<int64_t>(E, 1));
  int64_t off = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t sp = hp.sptr[i];
    const int k = hp.sptr[i + 1] - sp;
    for (int32_t s = sp; s < sp + k; ++s) {
      const int32_t col = hp.slot_col[s];
      const int rc = c->ha.ptr[col + 1] - c->ha.ptr[col];
      if (rc > 65535 || k > 65535) { plan.sparse_unavailable = true; return SPAI_OK; }   // the row sweep handles it
      meta[s].sp = sp; meta[s].k = (uint16_t)k; meta[s].rc = (uint16_t)rc; meta[s].off = off;
      off += rc;
    }
  }
  SlotMeta* meta_dev = nullptr;
  SPAI_TRY(ar.upload(&meta_dev, meta));
  plan.sl_meta = meta_dev;
  const unsigned blocks = (unsigned)ceil_div(std::max<int64_t>(n, 1), 128);
  plan.base_prefix.assign((size_t)n + 1, 0.0);
  if (dtype == SPAI_F32) {
    Pair<float>* out = nullptr;
    SPAI_TRY(ar.alloc(&out, off));
    k3s_build_kernel<float><<<blocks, 128, 0, st>>>(reinterpret_cast<const Rec32*>(plan.rec_copy), plan.cptr, plan.c_col,
                                                    plan.rhdr, c->P.slot_col, c->A.ptr, c->A.col, meta_dev, n, out);
    plan.sl_rec = out;
    std::vector<float> rb((size_t)std::max<int64_t>(n, 1));
    SPAI_CUDA(cudaMemcpyAsync(rb.data(), plan.row_base, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    SPAI_CUDA(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < n; ++i) plan.base_prefix[i + 1] = plan.base_prefix[i] + (double)rb[i];
  } else {
    Pair<double>* out = nullptr;
    SPAI_TRY(ar.alloc(&out, off));
    k3s_build_kernel<double><<<blocks, 128, 0, st>>>(reinterpret_cast<const Rec64*>(plan.rec_copy), plan.cptr, plan.c_col,
                                                     plan.rhdr, c->P.slot_col, c->A.ptr, c->A.col, meta_dev, n, out);
    plan.sl_rec = out;
    std::vector<double> rb((size_t)std::max<int64_t>(n, 1));
    SPAI_CUDA(cudaMemcpyAsync(rb.data(), plan.row_base, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    SPAI_CUDA(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < n; ++i) plan.base_prefix[i + 1] = plan.base_prefix[i] + rb[i];
  }
  SPAI_CUDA(cudaGetLastError());
  SPAI_CUDA(cudaStreamSynchronize(st));
  plan.sptr_host = hp.sptr.data();
  plan.a_ptr = c->A.ptr; plan.a_col = c->A.col;
  plan.bytes = ar.bytes;
  plan.sparse_ready = true;
  return SPAI_OK;
}

// ||M A - I||_F with every slot of P kept (B = 1).
static int residual_all_kept(const Pattern& P, const Plan& plan, int dtype, int sm_count,
                             double* res_out, cudaStream_t st) {
  const int64_t W = P.words();
  Arena tmp;
  uint32_t* mask = nullptr;
  SPAI_TRY(tmp.alloc(&mask, W));
  double* out = nullptr;
  SPAI_TRY(tmp.alloc(&out, 1));
  if (W) {
    k0_mask_init_kernel<<<(unsigned)std::min<int64_t>(ceil_div(W, 256), 65535), 256, 0, st>>>(mask, W, P.E, 1);
    SPAI_CUDA(cudaGetLastError());
  }
  const EvalShape s = plan_shape(plan, SPAI_MODE_COPY, dtype, 1, sm_count);
  const int64_t need = eval_bytes(plan, s, W, dtype);
  char* scratch = nullptr;
  SPAI_TRY(tmp.alloc(&scratch, need));
  SPAI_TRY(eval_masks(P, plan, SPAI_MODE_COPY, dtype, mask, 1, scratch, need, sm_count, (double)P.n, 1.0,
                      1.0, 0.5, nullptr, out, nullptr, st, nullptr, nullptr));
  SPAI_CUDA(cudaMemcpyAsync(res_out, out, sizeof(double), cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));
  return SPAI_OK;
}

static int pair_residual(int sm_count, int64_t n, int64_t m_nnz, const int64_t* mr, const int64_t* mc,
                         const double* mv, const HostCsr& ha, const CsrA& A, int dtype, double* res,
                         int64_t* m_nnz_out, cudaStream_t st) {
  Arena ar;
  HostPattern hp;
  Pattern P;
  SPAI_TRY(build_host_pattern(n, m_nnz, mr, mc, mv, hp, P));
  // coalesce(): repeated coordinates are summed into one stored entry
  if (P.ndup) {
    HostCsr hm;
    SPAI_TRY(build_host_csr(n, m_nnz, mr, mc, mv, hm));
    std::vector<int64_t> r2(hm.col.size()), c2(hm.col.size());
    for (int64_t i = 0; i < n; ++i)
      for (int32_t p = hm.ptr[i]; p < hm.ptr[i + 1]; ++p) { r2[p] = i; c2[p] = hm.col[p]; }
    std::vector<double> v2 = hm.val64;
    if (dtype == SPAI_F32) for (size_t p = 0; p < v2.size(); ++p) v2[p] = hm.val32[p];
    hp = HostPattern(); P = Pattern();
    SPAI_TRY(build_host_pattern(n, (int64_t)r2.size(), r2.data(), c2.data(), v2.data(), hp, P));
  }
  SPAI_TRY(upload_pattern(ar, hp, P));
  Plan plan;
  SPAI_TRY(build_plan(ar, P, hp, A, ha, dtype, false, plan, st));
  SPAI_TRY(residual_all_kept(P, plan, dtype, sm_count, res, st));
  if (m_nnz_out) *m_nnz_out = P.init_nnz;
  return SPAI_OK;
}

}  // namespace spai

extern "C" {

int spai_abi_version(void) { return SPAI_ABI_VERSION; }
const char* spai_last_error(void) { return spai::g_err.c_str(); }

int spai_device_count(int* count) {
  if (!count) return SPAI_ERR_INVALID;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) { *count = 0; set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); cudaGetLastError(); return SPAI_ERR_CUDA; }
  return SPAI_OK;
}

int spai_ctx_create(int device, int64_t n, int64_t num_edges, const int64_t* edge_row,
                    const int64_t* edge_col, const double* edge_val, int64_t a_nnz,
                    const int64_t* a_row, const int64_t* a_col, const double* a_val, spai_ctx** out) {
  if (!out) return SPAI_ERR_INVALID;
  *out = nullptr;
  if (n <= 0 || num_edges < 0 || a_nnz < 0 || n >= ((int64_t)1 << 31) ||
      (num_edges && (!edge_row || !edge_col || !edge_val)) || (a_nnz && (!a_row || !a_col || !a_val))) {
    set_error("spai_ctx_create: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  int ndev = 0;
  SPAI_TRY(spai_device_count(&ndev));
  if (device < 0 || device >= ndev) { set_error("device %d not available (%d CUDA devices)", device, ndev); return SPAI_ERR_CUDA; }
  DeviceGuard guard(device);
  spai_ctx* c = new spai_ctx();
  c->device = device; c->n = n;
  auto fail = [&](int s) { delete c; return s; };
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  int s;
  if ((s = build_host_pattern(n, num_edges, edge_row, edge_col, edge_val, c->hp, c->P))) return fail(s);
  if ((s = upload_pattern(c->arena, c->hp, c->P))) return fail(s);
  if ((s = build_host_csr(n, a_nnz, a_row, a_col, a_val, c->ha))) return fail(s);
  if ((s = upload_csr(c->arena, c->ha, n, a_nnz, c->A))) return fail(s);
  c->flops0 = 2 * a_nnz * n;                       // preconditioner.py:71-72 on the stored values
  if ((s = c->pt.init())) return fail(s);

  // baselines ||A0 A0 - I||_F (preconditioner.py:28): pattern = A0 itself
  {
    std::vector<int64_t> r(c->ha.col.size()), cc(c->ha.col.size());
    for (int64_t i = 0; i < n; ++i)
      for (int32_t p = c->ha.ptr[i]; p < c->ha.ptr[i + 1]; ++p) { r[p] = i; cc[p] = c->ha.col[p]; }
    for (int dt = 0; dt < 2; ++dt) {
      std::vector<double> v = c->ha.val64;
      if (dt == SPAI_F32) for (size_t p = 0; p < v.size(); ++p) v[p] = c->ha.val32[p];
      if ((s = pair_residual(c->sm_count, n, (int64_t)r.size(), r.data(), cc.data(), v.data(), c->ha, c->A,
                             dt, &c->res0[dt], nullptr, 0)))
        return fail(s);
    }
  }
  // the fp32 plan is what the reference-comparable path needs; build it eagerly
  if ((s = ensure_plan(c, SPAI_F32, false, 0))) return fail(s);
  *out = c;
  return SPAI_OK;
}

void spai_ctx_destroy(spai_ctx* c) {
  if (!c) return;
  DeviceGuard guard(c->device);
  c->pt.destroy();
  if (c->build_ev) cudaEventDestroy(c->build_ev);
  delete c;
}

int spai_ctx_info(const spai_ctx* c, spai_info* o) {
  if (!c || !o) return SPAI_ERR_INVALID;
  memset(o, 0, sizeof(*o));
  o->n = c->n; o->num_edges = c->P.E; o->init_nnz = c->P.init_nnz; o->num_actions = c->P.init_nnz + 1;
  o->a_nnz_stored = c->A.nnz_stored; o->a_nnz = c->A.nnz; o->orig_flops = c->flops0;
  o->orig_residual_f32 = c->res0[SPAI_F32]; o->orig_residual_f64 = c->res0[SPAI_F64];
  const Plan* p = c->plan_ready[SPAI_F32] ? &c->plan[SPAI_F32] : (c->plan_ready[SPAI_F64] ? &c->plan[SPAI_F64] : nullptr);
  if (p) {
    o->contributions = p->nc; o->max_row_union = p->max_q; o->rows_missing_diag = p->rows_missing_diag;
    for (int i = 0; i < LS_NCLASS; ++i) o->ls_class_rows[i] = p->class_count[i];
  }
  o->max_row_slots = c->P.max_k; o->has_duplicates = c->P.ndup ? 1 : 0; o->device = c->device;
  o->device_bytes = c->arena.bytes + c->plan_arena[0].bytes + c->plan_arena[1].bytes + c->ws.bytes;
  return SPAI_OK;
}

int spai_ctx_set_workspace_limit(spai_ctx* c, int64_t bytes) {
  if (!c || bytes < ((int64_t)1 << 20)) return SPAI_ERR_INVALID;
  c->ws_limit = bytes;
  return SPAI_OK;
}

int spai_ctx_set_deletion_hint(spai_ctx* c, int64_t max_deletions) {
  if (!c || max_deletions < 0) return SPAI_ERR_INVALID;
  c->deletion_hint = max_deletions;
  return SPAI_OK;
}

int spai_ctx_k3m_rows(const spai_ctx* c, int64_t* rows16, int64_t* rows32) {
  if (!c || !rows16 || !rows32) { set_error("spai_ctx_k3m_rows: null argument"); return SPAI_ERR_INVALID; }
  const Plan& p = c->plan[SPAI_F32];
  *rows16 = p.mma_ready ? p.mma_count[0] : 0;
  *rows32 = p.mma_ready ? p.mma_count[1] : 0;
  return SPAI_OK;
}

int spai_ctx_enable_timing(spai_ctx* c, int enable) {
  if (!c) return SPAI_ERR_INVALID;
  c->pt.on = enable != 0;
  return SPAI_OK;
}
int spai_ctx_last_timing(const spai_ctx* c, spai_timing* out) {
  if (!c || !out) return SPAI_ERR_INVALID;
  *out = c->last;
  return SPAI_OK;
}

}  // extern "C"

namespace spai {

// Host side of spai_reward_batch_host: find, for every row of the -1 padded
// action matrix, the length of the prefix that still holds ids (scan back from
// the end over the padding), so only that prefix crosses PCIe. Exact for any
// input (-1 anywhere else is simply copied and ignored by the kernels).
static inline int64_t trimmed_len(const int64_t* row, int64_t T) {
  int64_t t = T;
  while (t >= 8) {
    const int64_t* q = row + t - 8;
    if ((q[0] & q[1] & q[2] & q[3] & q[4] & q[5] & q[6] & q[7]) != -1) break;
    t -= 8;
  }
  while (t > 0 && row[t - 1] == -1) --t;
  return t;
}

struct RowTrimmer {          // worker threads publish lengths group by group
  static constexpr int64_t GROUP = 16;
  std::vector<int32_t> len;
  std::vector<std::atomic<int>> ready;
  std::vector<std::thread> workers;
  RowTrimmer(const int64_t* base, int64_t B, int64_t T, int64_t ld)
      : len(B), ready((size_t)ceil_div(std::max<int64_t>(B, 1), GROUP)) {
    for (auto& r : ready) r.store(0, std::memory_order_relaxed);
    int nt = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) {        // one process per GPU shares the host cores
      const int lw = atoi(e);
      if (lw > 1) nt = std::max(2, nt / lw);
    }
    if (const char* e = getenv("SPAI_HOST_THREADS")) nt = atoi(e);
    nt = std::max(1, std::min(nt, 32));
    const int64_t ngroups = (int64_t)ready.size();
    nt = (int)std::min<int64_t>(nt, ngroups);
    auto next = std::make_shared<std::atomic<int64_t>>(0);
    for (int w = 0; w < nt; ++w)
      workers.emplace_back([this, base, B, T, ld, ngroups, next]() {
        for (;;) {
          const int64_t g = next->fetch_add(1);
          if (g >= ngroups) return;
          const int64_t hi = std::min(B, (g + 1) * GROUP);
          for (int64_t b = g * GROUP; b < hi; ++b) len[b] = (int32_t)trimmed_len(base + b * ld, T);
          ready[g].store(1, std::memory_order_release);
        }
      });
  }
  void wait(int64_t g) {
    while (!ready[g].load(std::memory_order_acquire)) std::this_thread::yield();
  }
  ~RowTrimmer() { for (auto& t : workers) t.join(); }
};

// Scratch of the two-pass mask build (K0b, k0b_bucket.cuh) for one group of trajectories.
struct K0bScratch {
  uint16_t* stage = nullptr;     // [group][ld_stage] sorted 16-bit local ids
  uint16_t* hdr = nullptr;       // [group][nchunks][C + 1]
  int64_t group = 0, ld_stage = 0, nchunks = 0;
  int C = 0;
  int threads = K0B_THREADS;     // sort CTA size
  int ids = K0B_IDS;             // ids per thread: 16, or 8 with 512-thread CTAs (SPAI_K0B_IDS=8: same 4096-id chunk, more warps)
  int64_t chunk() const { return (int64_t)threads * ids; }
};
// sort CTA size: SPAI_K0B_THREADS=256|512 (A/B); 256-thread CTAs need C <= 256
static inline int k0b_threads(int64_t C) {
  const char* e = getenv("SPAI_K0B_THREADS");
  int t = e ? atoi(e) : 256;               // measured (cfg3, B = 1024): sort 2.13 ms with 256 threads, 2.58 ms with 512
  if (t != 256 && t != 512) t = 256;
  if (C > t) t = K0B_THREADS;
  return t;
}
static inline bool k0_fits_smem(int64_t W) { return W * 4 <= K0S_MAX_SMEM; }
static inline int k0_variant() {             // A/B + test switch, read per call
  const char* e = getenv("SPAI_K0_VARIANT");
  return !e ? 0 : (!strcmp(e, "red") ? 1 : (!strcmp(e, "smem") ? 2 : (!strcmp(e, "bucket") ? 3 : (!strcmp(e, "cluster") ? 4 : 0))));
}
// K0c (k0c_cluster.cuh): one cluster per trajectory, the bitmask sliced over the CTAs' shared memory.
// Smallest cluster whose slice + exchange buffers fit 227 KB; SPAI_K0C_CS / SPAI_K0C_GROUPS override (A/B, tests).
struct K0cGeom { int cs = 0, spc = 0, cap = 0, nh = 2; size_t smem = 0; uint32_t inv = 0; };
static bool k0c_geometry(const Pattern& P, K0cGeom* g) {
  const int64_t C = ceil_div(P.E, (int64_t)1 << K0C_SEG_SHIFT);
  const char* e = getenv("SPAI_K0C_GROUPS");
  const int nh_forced = e ? atoi(e) : 0;
  const char* f = getenv("SPAI_K0C_CS");
  const int forced = f ? atoi(f) : 0;
  // most groups first (their phases overlap), then the smallest cluster, then the roomiest inbox slot
  for (int nh = 4; nh >= 1; --nh) {
    if (nh_forced >= 1 && nh_forced <= 4 && nh != nh_forced) continue;
    if (!(nh_forced >= 1 && nh_forced <= 4) && nh == 4) continue;          // default: at most 3 groups
    for (int cs = 2; cs <= K0C_MAX_CS; ++cs) {
      if (forced >= 2 && forced <= K0C_MAX_CS && cs != forced) continue;
      const int64_t spc = ceil_div(C, (int64_t)cs);
      for (double sig = 6.0; sig >= 4.0; sig -= 0.5) {
        const int cap = k0c_cap(cs, sig);
        const size_t smem = k0c_smem((int)spc, cs, cap, nh);
        if (smem > (size_t)K0C_MAX_SMEM) continue;
        g->cs = cs; g->spc = (int)spc; g->cap = cap; g->nh = nh; g->smem = smem;
        g->inv = (uint32_t)((65536 + spc - 1) / spc);
        return true;
      }
    }
  }
  return false;
}
static inline bool k0c_applies(const Pattern& P, int64_t T) {
  // opt-in (SPAI_K0_VARIANT=cluster): measured on B200 at 6.4-6.7 ms per 1024 cfg3 trajectories against 3.87 ms for K0b
  //  — ~90 instructions per id, mostly per-round bookkeeping amortised over 8 ids per thread
  if (k0_variant() != 4 || T <= 0) return false;
  K0cGeom g;
  return k0c_geometry(P, &g);
}
// K0b applies when the bitmask exceeds one CTA's shared memory and the segment count fits the sort kernel
static inline bool k0b_applies(const Pattern& P, int64_t T) {
  if (k0_variant() == 1 || T <= 0 || k0c_applies(P, T)) return false;
  const int64_t C = ceil_div(P.E, (int64_t)1 << K0B_SEG_SHIFT);
  if (C > K0B_MAX_SEGS) return false;
  return !k0_fits_smem(P.words()) || k0_variant() == 3;
}
// trajectories per sort+build launch pair. Measured on B200 (cfg3, B = 1024, tools/ab_k0.py): small groups whose
// staged ids would stay in L2 LOSE — 8: 12.3 ms, 16: 7.8, 32: 6.5, 64: 5.5, whole batch: 4.5 ms (launch tails
// dominate, the 2 + 2 staged bytes per id are cheap next to them) — so a group is as large as 8 GiB of staged
// ids allow; SPAI_K0B_GROUP overrides (A/B).
static int64_t k0b_group(int64_t bc, int64_t T, int64_t C) {
  const char* e = getenv("SPAI_K0B_GROUP");
  const int64_t forced = e ? atoll(e) : 0;
  if (forced > 0) return std::min(bc, forced);
  (void)C;
  const int64_t g = ((int64_t)8 << 30) / std::max<int64_t>(2 * round_up(std::max<int64_t>(T, 1), K0B_CHUNK_MAX), 1);
  return std::min(bc, std::max<int64_t>(g, 32));
}
static void k0b_plan(const Pattern& P, int64_t bc, int64_t T, K0bScratch* sc) {
  sc->C = (int)ceil_div(P.E, (int64_t)1 << K0B_SEG_SHIFT);
  sc->threads = k0b_threads(sc->C);
  // SPAI_K0B_IDS=8: 512-thread sort CTAs with 8 ids per thread (same 4096-id chunk, 1536 threads per SM): measured SLOWER
  // (cfg3 sort 2.54 -> 3.53 ms, cfg5 3.64 -> 5.08 ms: twice the per-warp counter arrays to zero and scan); kept for A/B
  { const char* e = getenv("SPAI_K0B_IDS"); sc->ids = (e && atoi(e) == 8) ? 8 : K0B_IDS; if (sc->ids == 8) sc->threads = 512; }
  sc->nchunks = ceil_div(T, sc->chunk());
  sc->ld_stage = sc->nchunks * sc->chunk();
  sc->group = k0b_group(bc, T, sc->C);
}
static int64_t k0b_bytes(const K0bScratch& sc) {
  return padded(sc.group * sc.ld_stage * 2) + padded(sc.group * sc.nchunks * (sc.C + 1) * 2);
}

// K0 for rows [0, bc) of `act` (device-accessible pointer, ids of `elem` bytes, leading
// dimension act_ld): shared-memory bitmask per trajectory when it fits, else the two-pass
// segment build (K0b); init + global RED only for patterns beyond K0b's envelope.
static int launch_k0(const Pattern& P, const void* act, int elem, int64_t bc, int64_t T, int64_t act_ld,
                     uint32_t* mask, long long* nnz0, const int32_t* row_len, const K0bScratch* sc,
                     cudaStream_t st, int* launches, bool* nnz_fused) {
  const int64_t W = P.words();
  const int32_t* eslot = P.identity_perm ? nullptr : P.edge_slot;
  const int64_t* a64 = reinterpret_cast<const int64_t*>(act);
  const int32_t* a32 = reinterpret_cast<const int32_t*>(act);
  if (k0c_applies(P, T)) {
    K0cGeom g;
    k0c_geometry(P, &g);
    SPAI_CUDA(cudaMemsetAsync(nnz0, 0, (size_t)bc * 8, st));
    if (bc * g.cs >= ((int64_t)1 << 31)) { set_error("batch too large for one launch"); return SPAI_ERR_UNSUPPORTED; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(bc * g.cs));
    cfg.blockDim = dim3((unsigned)(g.nh * K0C_HT));
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)g.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    unsigned long long* nnzu = reinterpret_cast<unsigned long long*>(nnz0);
#define SPAI_K0C_LAUNCH(IDT, PTR, MAP, NH)                                                                              \
  do {                                                                                                                  \
    SPAI_CUDA(cudaFuncSetAttribute(k0c_cluster_kernel<IDT, MAP, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem)); \
    SPAI_CUDA(cudaLaunchKernelEx(&cfg, k0c_cluster_kernel<IDT, MAP, NH>, PTR, T, act_ld, row_len, eslot, P.E, g.cs, g.spc, g.cap, \
                                 g.inv, mask, W, nnzu));                                                                \
  } while (0)
#define SPAI_K0C_LAUNCH2(IDT, PTR, MAP) \
  do { if (g.nh == 1) SPAI_K0C_LAUNCH(IDT, PTR, MAP, 1); else if (g.nh == 2) SPAI_K0C_LAUNCH(IDT, PTR, MAP, 2); else if (g.nh == 3) SPAI_K0C_LAUNCH(IDT, PTR, MAP, 3); else SPAI_K0C_LAUNCH(IDT, PTR, MAP, 4); } while (0)
    if (elem == 8) { if (eslot) SPAI_K0C_LAUNCH2(int64_t, a64, true); else SPAI_K0C_LAUNCH2(int64_t, a64, false); }
    else { if (eslot) SPAI_K0C_LAUNCH2(int32_t, a32, true); else SPAI_K0C_LAUNCH2(int32_t, a32, false); }
#undef SPAI_K0C_LAUNCH2
#undef SPAI_K0C_LAUNCH
    ++*launches;
    *nnz_fused = true;
  } else if (k0_fits_smem(W) && k0_variant() != 1 && k0_variant() != 3 && T > 0) {
    if (elem == 8) {
      SPAI_CUDA(cudaFuncSetAttribute(k0_mask_build_smem_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, K0S_MAX_SMEM));
      k0_mask_build_smem_kernel<int64_t><<<(unsigned)bc, K0S_THREADS, (size_t)W * 4, st>>>(a64, bc, T, act_ld, eslot, P.E, mask, W, nnz0, row_len);
    } else {
      SPAI_CUDA(cudaFuncSetAttribute(k0_mask_build_smem_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, K0S_MAX_SMEM));
      k0_mask_build_smem_kernel<int32_t><<<(unsigned)bc, K0S_THREADS, (size_t)W * 4, st>>>(a32, bc, T, act_ld, eslot, P.E, mask, W, nnz0, row_len);
    }
    SPAI_CUDA(cudaGetLastError()); ++*launches;
    *nnz_fused = true;
  } else if (sc && sc->stage && k0b_applies(P, T)) {
    const int C = sc->C;
    const int TH = sc->threads;
    const size_t ssm = k0b_sort_smem(C, TH, sc->ids);
    const size_t bsm = (size_t)std::min<int64_t>((int64_t)K0B_R, C) * K0B_SEG_WORDS * 4;
    const int tasks = (int)ceil_div(C, (int64_t)K0B_R);
    const int smax = (int)k0b_sort_smem(K0B_MAX_SEGS, K0B_THREADS);
#define SPAI_K0B_ATTR(IDT, MAP, THR) \
  SPAI_CUDA(cudaFuncSetAttribute(k0b_sort_kernel<IDT, MAP, THR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smax))
    SPAI_K0B_ATTR(int64_t, false, 512); SPAI_K0B_ATTR(int64_t, true, 512); SPAI_K0B_ATTR(int32_t, false, 512); SPAI_K0B_ATTR(int32_t, true, 512);
    SPAI_K0B_ATTR(int64_t, false, 256); SPAI_K0B_ATTR(int64_t, true, 256); SPAI_K0B_ATTR(int32_t, false, 256); SPAI_K0B_ATTR(int32_t, true, 256);
#undef SPAI_K0B_ATTR
#define SPAI_K0B_ATTR8(IDT, MAP) \
  SPAI_CUDA(cudaFuncSetAttribute(k0b_sort_kernel<IDT, MAP, 512, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smax))
    SPAI_K0B_ATTR8(int64_t, false); SPAI_K0B_ATTR8(int64_t, true); SPAI_K0B_ATTR8(int32_t, false); SPAI_K0B_ATTR8(int32_t, true);
#undef SPAI_K0B_ATTR8
    SPAI_CUDA(cudaFuncSetAttribute(k0b_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K0B_R * K0B_SEG_WORDS * 4));
    SPAI_CUDA(cudaFuncSetAttribute(k0b_build2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, K0B_R * K0B_SEG_WORDS * 4));
    SPAI_CUDA(cudaFuncSetAttribute(k0b_build2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, K0B_R * K0B_SEG_WORDS * 4));
    SPAI_CUDA(cudaFuncSetAttribute(k0b_build2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, K0B_R * K0B_SEG_WORDS * 4));
    SPAI_CUDA(cudaFuncSetAttribute(k0b_build3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K0B_R * K0B_SEG_WORDS * 4));
    SPAI_CUDA(cudaMemsetAsync(nnz0, 0, (size_t)bc * 8, st));
    static const bool timing = getenv("SPAI_K0B_TIMING") != nullptr;      // per-pass device times on stderr (diagnostic)
    cudaEvent_t ev[3] = {};
    float ms_sort = 0, ms_build = 0;
    if (timing) for (auto& e : ev) cudaEventCreate(&e);
    for (int64_t g0 = 0; g0 < bc; g0 += sc->group) {
      const int64_t gb = std::min(sc->group, bc - g0);
      const int64_t nblk = gb * sc->nchunks;
      if (nblk >= ((int64_t)1 << 31) || gb * tasks >= ((int64_t)1 << 31)) { set_error("batch too large for one launch"); return SPAI_ERR_UNSUPPORTED; }
      const int32_t* rl = row_len ? row_len + g0 : nullptr;
      if (timing) cudaEventRecord(ev[0], st);
#define SPAI_K0B_SORT(IDT, PTR, MAP, THR)                                                                         \
  k0b_sort_kernel<IDT, MAP, THR><<<(unsigned)nblk, THR, ssm, st>>>(PTR + g0 * act_ld, T, act_ld, rl, eslot, P.E, C, \
                                                                   sc->stage, sc->ld_stage, sc->hdr, sc->nchunks)
#define SPAI_K0B_SORT8(IDT, PTR, MAP)                                                                             \
  k0b_sort_kernel<IDT, MAP, 512, 8><<<(unsigned)nblk, 512, ssm, st>>>(PTR + g0 * act_ld, T, act_ld, rl, eslot, P.E, C, \
                                                                      sc->stage, sc->ld_stage, sc->hdr, sc->nchunks)
#define SPAI_K0B_SORT2(IDT, PTR, MAP) do { if (sc->ids == 8) SPAI_K0B_SORT8(IDT, PTR, MAP); else if (TH == 256) SPAI_K0B_SORT(IDT, PTR, MAP, 256); else SPAI_K0B_SORT(IDT, PTR, MAP, 512); } while (0)
      // SPAI_K0B_SORT=2: persistent CTAs with the next chunk prefetched by cp.async (256-thread CTAs, C <= 256). Measured SLOWER
      // than the one-CTA-per-chunk kernel (cfg3 2.55 -> 2.93 ms, cfg5 3.64 -> 4.11 ms); kept for A/B only
      const int sort_v = getenv("SPAI_K0B_SORT") ? atoi(getenv("SPAI_K0B_SORT")) : 1;
      if (sort_v == 2 && TH == 256 && sc->ids == K0B_IDS) {
        static int sms = 0;
        if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
        const size_t ssm2 = k0b_sort2_smem(C, 256, elem);
        const unsigned grid2 = (unsigned)std::min<int64_t>(nblk, (int64_t)sms * 4);
#define SPAI_K0B_SORTP(IDT, PTR, MAP)                                                                                         \
  do {                                                                                                                        \
    SPAI_CUDA(cudaFuncSetAttribute(k0b_sort2_kernel<IDT, MAP, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm2)); \
    k0b_sort2_kernel<IDT, MAP, 256><<<grid2, 256, ssm2, st>>>(PTR + g0 * act_ld, T, act_ld, rl, eslot, P.E, C, sc->stage,     \
                                                              sc->ld_stage, sc->hdr, sc->nchunks, nblk);                       \
  } while (0)
        if (elem == 8) { if (eslot) SPAI_K0B_SORTP(int64_t, a64, true); else SPAI_K0B_SORTP(int64_t, a64, false); }
        else { if (eslot) SPAI_K0B_SORTP(int32_t, a32, true); else SPAI_K0B_SORTP(int32_t, a32, false); }
#undef SPAI_K0B_SORTP
      } else if (elem == 8) { if (eslot) SPAI_K0B_SORT2(int64_t, a64, true); else SPAI_K0B_SORT2(int64_t, a64, false); }
      else { if (eslot) SPAI_K0B_SORT2(int32_t, a32, true); else SPAI_K0B_SORT2(int32_t, a32, false); }
#undef SPAI_K0B_SORT2
#undef SPAI_K0B_SORT8
#undef SPAI_K0B_SORT
      SPAI_CUDA(cudaGetLastError()); ++*launches;
      if (timing) cudaEventRecord(ev[1], st);
      const int build_v = getenv("SPAI_K0B_BUILD") ? atoi(getenv("SPAI_K0B_BUILD")) : 2;   // 1, 3: other versions (A/B, tests); read per call
      if (build_v == 1)
        k0b_build_kernel<<<(unsigned)(gb * tasks), K0B_THREADS, bsm, st>>>(
            sc->stage, sc->ld_stage, sc->hdr, sc->nchunks, C, rl, T, P.E, mask + g0 * W, W,
            reinterpret_cast<unsigned long long*>(nnz0 + g0), tasks, (int)sc->chunk());
      else if (build_v == 3)
        k0b_build3_kernel<<<(unsigned)(gb * tasks), K0B_THREADS, bsm, st>>>(
            sc->stage, sc->ld_stage, sc->hdr, sc->nchunks, C, rl, T, P.E, mask + g0 * W, W,
            reinterpret_cast<unsigned long long*>(nnz0 + g0), tasks, (int)sc->chunk());
      else {
        // lanes per (chunk, segment) run from the mean run length chunk / C (SPAI_K0B_LANES overrides, A/B)
        const int lanes_forced = getenv("SPAI_K0B_LANES") ? atoi(getenv("SPAI_K0B_LANES")) : 0;
        const int64_t run = sc->chunk() / std::max(C, 1);
        const int lanes = (lanes_forced == 1 || lanes_forced == 2 || lanes_forced == 4) ? lanes_forced : (run >= 48 ? 4 : 2);
#define SPAI_K0B_BUILD2(LN)                                                                          \
  k0b_build2_kernel<LN><<<(unsigned)(gb * tasks), K0B_THREADS, bsm, st>>>(                           \
      sc->stage, sc->ld_stage, sc->hdr, sc->nchunks, C, rl, T, P.E, mask + g0 * W, W,                \
      reinterpret_cast<unsigned long long*>(nnz0 + g0), tasks, (int)sc->chunk())
        if (lanes == 4) SPAI_K0B_BUILD2(4); else if (lanes == 2) SPAI_K0B_BUILD2(2); else SPAI_K0B_BUILD2(1);
#undef SPAI_K0B_BUILD2
      }
      SPAI_CUDA(cudaGetLastError()); ++*launches;
      if (timing) {
        cudaEventRecord(ev[2], st);
        cudaEventSynchronize(ev[2]);
        float t;
        cudaEventElapsedTime(&t, ev[0], ev[1]); ms_sort += t;
        cudaEventElapsedTime(&t, ev[1], ev[2]); ms_build += t;
      }
    }
    if (timing) {
      fprintf(stderr, "[k0b] B=%lld T=%lld C=%d threads=%d group=%lld: sort %.3f ms, build %.3f ms\n", (long long)bc,
              (long long)T, C, TH, (long long)sc->group, ms_sort, ms_build);
      for (auto& e : ev) cudaEventDestroy(e);
    }
    *nnz_fused = true;
  } else {
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(W * bc, 256), 148 * 32);
    k0_mask_init_kernel<<<blocks, 256, 0, st>>>(mask, W, P.E, bc);
    SPAI_CUDA(cudaGetLastError()); ++*launches;
    if (T > 0) {
      constexpr int U = 4;
      const int64_t chunks_t = ceil_div(T, 256 * U);
      const int64_t nblk = chunks_t * bc;
      if (nblk >= ((int64_t)1 << 31)) { set_error("batch too large for one launch"); return SPAI_ERR_UNSUPPORTED; }
      if (elem == 8)
        k0_mask_clear_kernel<U, int64_t><<<(unsigned)nblk, 256, 0, st>>>(a64, bc, T, act_ld, eslot, P.E, mask, W, chunks_t, row_len);
      else
        k0_mask_clear_kernel<U, int32_t><<<(unsigned)nblk, 256, 0, st>>>(a32, bc, T, act_ld, eslot, P.E, mask, W, chunks_t, row_len);
      SPAI_CUDA(cudaGetLastError()); ++*launches;
    }
    *nnz_fused = false;
  }
  return SPAI_OK;
}

enum MaskSource { FROM_ACTIONS_DEV, FROM_ACTIONS_HOST, FROM_TAKEN_DEV };

// One reward call. Everything a call depends on travels here (nothing per-call is
// kept in the context, so a context can serve calls with different parameters back to back).
struct RewardCall {
  MaskSource src = FROM_ACTIONS_DEV;
  const void* input = nullptr;          // actions (elem bytes per id) or taken bitmask
  int elem = 8;                         // 8: int64 ids (the reference's format), 4: int32 ids
  const int32_t* row_len = nullptr;     // optional valid length of every row (host array for the host entry,
                                        // device array for the device entry); entries beyond it are not read
  int64_t B = 0, T = 0, ld = 0;
  double alpha = 0.5;
  int mode = SPAI_MODE_COPY, dtype = SPAI_F32;
  double* reward = nullptr; double* residual = nullptr; int64_t* nnz_m = nullptr;
  bool out_host = false;
  uint8_t* kept_bytes_dev = nullptr;    // mask-only call (spai_kept_mask_dev)
  int64_t row_lo = 0, row_hi = -1;      // row-range evaluation (spai_reward_rows_dev)
  bool partial_only = false;
  int64_t t_hint = 0;                   // longest trajectory (0 = unknown); selects a kernel only
  void* stream = nullptr;
};

// Shared driver of the reward entry points: chunk the batch so the scratch
// stays under the workspace limit, build slot-order masks, evaluate.
static int reward_driver(spai_ctx* c, const RewardCall& rc) {
  const MaskSource src = rc.src;
  const int64_t B = rc.B, T = rc.T, ld = rc.ld;
  const int mode = rc.mode, dtype = rc.dtype;
  if (!c || B < 0 || (B && !rc.input && (src == FROM_TAKEN_DEV || T > 0)) || (mode != SPAI_MODE_COPY && mode != SPAI_MODE_LS && mode != SPAI_MODE_LS_GRAM) ||
      (dtype != SPAI_F32 && dtype != SPAI_F64) || (src != FROM_TAKEN_DEV && (T < 0 || ld < T)) || (rc.elem != 8 && rc.elem != 4)) {
    set_error("reward: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  if (B == 0) return SPAI_OK;
  DeviceGuard guard(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(rc.stream);
  const bool mask_only = rc.kept_bytes_dev != nullptr;
  const bool out_host = rc.out_host;
  const int64_t esz = rc.elem;
  // one-time builds are enqueued on the stream of the call that triggers them; later calls on
  // other streams wait for the recorded event before they read the tables
  SPAI_TRY(c->wait_builds(st));
  if (!mask_only) SPAI_TRY(ensure_plan(c, dtype, mode != SPAI_MODE_COPY, st));
  if (!mask_only && mode == SPAI_MODE_LS_GRAM) SPAI_TRY(ensure_gram(c, dtype, st));
  if (!mask_only && mode == SPAI_MODE_LS_GRAM && B >= 64) SPAI_TRY(ensure_lut_ls(c, dtype, st));
  if (!mask_only && mode == SPAI_MODE_LS && B >= 64) SPAI_TRY(ensure_lut_qr(c, dtype, st));
  if (!mask_only && mode == SPAI_MODE_COPY && B >= 64) SPAI_TRY(ensure_lut(c, dtype, st));
  if (!mask_only && mode == SPAI_MODE_COPY && dtype == SPAI_F32 && B >= 64 && !c->plan[dtype].lut_ready && k3m_enabled())
    SPAI_TRY(ensure_mma(c, dtype, st));
  if (!mask_only) c->plan[dtype].tables_on = B >= 64;
  int64_t t_len = rc.t_hint;                                   // longest trajectory (0 = unknown)
  if (src != FROM_TAKEN_DEV) {
    t_len = T;
    if (rc.row_len && src == FROM_ACTIONS_HOST) {              // host lengths are free to inspect
      int64_t m = 0;
      for (int64_t b = 0; b < B; ++b) {
        if (rc.row_len[b] < 0) { set_error("reward: negative row length"); return SPAI_ERR_INVALID; }
        m = std::max<int64_t>(m, std::min<int64_t>(rc.row_len[b], T));
      }
      t_len = m;
    }
  }
  {
    const char* force = getenv("SPAI_K3_SPARSE");
    const bool want = force ? atoi(force) != 0 : (t_len > 0 && t_len * K3S_MIN_RATIO <= c->P.E);
    if (!mask_only && mode == SPAI_MODE_COPY && want) SPAI_TRY(ensure_sparse(c, dtype, st));
  }
  SPAI_TRY(c->mark_builds(st));
  const Plan& plan = c->plan[mask_only ? SPAI_F32 : dtype];
  const Pattern& P = c->P;
  const int64_t W = P.words();
  const bool use_k0b = src != FROM_TAKEN_DEV && k0b_applies(P, T);

  // chunk size
  int64_t Bc = B;
  auto need_for = [&](int64_t bc) {
    const int64_t bp = round_up(bc, 32);
    int64_t need = padded(bc * std::max<int64_t>(W, 1) * 4) + padded(bp * 8);   // mask + fused popcount
    if (src == FROM_ACTIONS_HOST) need += padded(bc * std::max<int64_t>(T, 1) * esz) + padded(bc * 4);
    if (use_k0b) { K0bScratch sc; k0b_plan(P, bc, T, &sc); need += k0b_bytes(sc); }
    if (out_host) need += 3 * padded(bp * 8);
    if (!mask_only) need += eval_bytes(plan, plan_shape(plan, mode, dtype, bc, c->sm_count), W, dtype);
    return need;
  };
  while (Bc > 32 && need_for(Bc) > c->ws_limit) Bc = std::max<int64_t>(32, round_up(Bc / 2, 32));
  const int64_t need = need_for(Bc);
  SPAI_TRY(c->ws.ensure(need));

  std::unique_ptr<RowTrimmer> trimmer;      // keeps the host length array alive until the stream is drained
  int64_t trimmer_b0 = -1;
  double h2d_bytes = 0;
  PhaseTimer* pt = &c->pt;
  float ms_masks = 0, ms_tr = 0, ms_rw = 0, ms_fin = 0;
  int launches = 0, chunks = 0;
  const double res0 = c->res0[dtype];
  for (int64_t b0 = 0; b0 < B; b0 += Bc) {
    const int64_t bc = std::min(Bc, B - b0);
    const int64_t bp = round_up(bc, 32);
    Carver cv{reinterpret_cast<char*>(c->ws.base), reinterpret_cast<char*>(c->ws.base) + c->ws.bytes};
    uint32_t* mask = cv.take<uint32_t>(bc * std::max<int64_t>(W, 1));
    const char* act_dev = nullptr;
    int64_t act_ld = ld;
    const int32_t* row_len_dev = nullptr;
    long long* nnz0 = cv.take<long long>(bp);
    const long long* nnz_ready = nullptr;
    bool k0_done = false;
    K0bScratch sc;
    if (use_k0b) {
      k0b_plan(P, bc, T, &sc);
      sc.stage = cv.take<uint16_t>(sc.group * sc.ld_stage);
      sc.hdr = cv.take<uint16_t>(sc.group * sc.nchunks * (sc.C + 1));
    }
    if (src == FROM_ACTIONS_HOST) {
      char* buf = cv.take<char>(bc * std::max<int64_t>(T, 1) * esz);
      int32_t* len_dev = cv.take<int32_t>(bc);
      const char* hbase = reinterpret_cast<const char*>(rc.input) + b0 * ld * esz;
      act_ld = T;
      if (T >= 4096 || rc.row_len) {
        // valid length of every row: the caller's (no host pass over the data at all), else worker
        // threads scan each row back over its -1 padding while this thread feeds the GPU
        const int32_t* hlen = rc.row_len ? rc.row_len + b0 : nullptr;
        if (!hlen && (!trimmer || trimmer_b0 != b0)) {
          if (esz != 8) { set_error("reward: int32 host actions need row lengths"); return SPAI_ERR_INVALID; }
          trimmer.reset(new RowTrimmer(reinterpret_cast<const int64_t*>(hbase), bc, T, ld));
          trimmer_b0 = b0;
        }
        auto len_of = [&](int64_t b) -> int64_t { return hlen ? std::min<int64_t>(hlen[b], T) : trimmer->len[b]; };
        // zero-copy: pinned host memory is device-accessible (UVA), so K0 can stream the valid
        // prefix of every row straight over PCIe into its shared-memory bitmask — no staging
        // buffer and no per-row copy call. Pageable input falls back to per-row cudaMemcpyAsync.
        const char* hdev = nullptr;
        static const bool force_memcpy = [] { const char* v = getenv("SPAI_H2D"); return v && !strcmp(v, "memcpy"); }();
        if (!force_memcpy && W > 0) {
          cudaPointerAttributes at;
          if (cudaPointerGetAttributes(&at, hbase) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            hdev = reinterpret_cast<const char*>(at.devicePointer);
          else
            cudaGetLastError();
        }
        if (hdev) {
          // with caller lengths one launch covers the chunk; with the trimmer K0 follows the workers
          const int64_t SG = hlen ? bc : 256;                          // rows per K0 launch
          if (pt->on) cudaEventRecord(pt->ev[0], st);
          if (hlen) SPAI_CUDA(cudaMemcpyAsync(len_dev, hlen, (size_t)bc * 4, cudaMemcpyHostToDevice, st));
          for (int64_t sg = 0; sg < bc; sg += SG) {
            const int64_t nr = std::min(SG, bc - sg);
            if (!hlen) {
              for (int64_t g = sg / RowTrimmer::GROUP; g * RowTrimmer::GROUP < sg + nr; ++g) trimmer->wait(g);
              SPAI_CUDA(cudaMemcpyAsync(len_dev + sg, trimmer->len.data() + sg, (size_t)nr * 4, cudaMemcpyHostToDevice, st));
            }
            bool fused = false;
            SPAI_TRY(launch_k0(P, hdev + sg * ld * esz, (int)esz, nr, T, ld, mask + sg * W, nnz0 + sg, len_dev + sg,
                               use_k0b ? &sc : nullptr, st, &launches, &fused));
            if (fused) nnz_ready = nnz0;
          }
          k0_done = true;
        } else {
          // staged: per-row prefix copies (pageable memory)
          for (int64_t b = 0; b < bc; ++b) {
            if (!hlen && b % RowTrimmer::GROUP == 0) trimmer->wait(b / RowTrimmer::GROUP);
            const int64_t n = len_of(b);
            if (n) SPAI_CUDA(cudaMemcpyAsync(buf + b * T * esz, hbase + b * ld * esz, (size_t)n * esz, cudaMemcpyHostToDevice, st));
          }
          if (hlen) SPAI_CUDA(cudaMemcpyAsync(len_dev, hlen, (size_t)bc * 4, cudaMemcpyHostToDevice, st));
          else SPAI_CUDA(cudaMemcpyAsync(len_dev, trimmer->len.data(), (size_t)bc * 4, cudaMemcpyHostToDevice, st));
        }
        row_len_dev = len_dev;
        h2d_bytes += (double)bc * 4;
        for (int64_t b = 0; b < bc; ++b) h2d_bytes += (double)esz * len_of(b);
      } else if (ld == T) {
        SPAI_CUDA(cudaMemcpyAsync(buf, hbase, (size_t)bc * T * esz, cudaMemcpyHostToDevice, st));
        h2d_bytes += (double)bc * T * esz;
      } else {
        SPAI_CUDA(cudaMemcpy2DAsync(buf, (size_t)T * esz, hbase, (size_t)ld * esz, (size_t)T * esz, (size_t)bc,
                                    cudaMemcpyHostToDevice, st));
        h2d_bytes += (double)bc * T * esz;
      }
      act_dev = buf;
    } else if (src == FROM_ACTIONS_DEV) {
      act_dev = reinterpret_cast<const char*>(rc.input) + b0 * ld * esz;
      row_len_dev = rc.row_len ? rc.row_len + b0 : nullptr;
    }
    double* o_rw = nullptr; double* o_rs = nullptr; int64_t* o_nz = nullptr;
    if (out_host) {
      o_rw = cv.take<double>(bp); o_rs = cv.take<double>(bp); o_nz = cv.take<int64_t>(bp);
    } else {
      o_rw = rc.reward ? rc.reward + b0 : nullptr; o_rs = rc.residual ? rc.residual + b0 : nullptr;
      o_nz = rc.nnz_m ? rc.nnz_m + b0 : nullptr;
    }
    if (pt->on && !k0_done) cudaEventRecord(pt->ev[0], st);
    if (W > 0) {
      if (src == FROM_TAKEN_DEV) {
        const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(W * bc, 256), 148 * 32);
        k0_mask_from_taken_kernel<<<blocks, 256, 0, st>>>(
            reinterpret_cast<const uint32_t*>(rc.input) + b0 * ld, ld, P.identity_perm ? nullptr : P.slot_edge,
            P.E, mask, W, bc);
        SPAI_CUDA(cudaGetLastError()); ++launches;
      } else {
        if (!k0_done) {
          bool fused = false;
          SPAI_TRY(launch_k0(P, act_dev, (int)esz, bc, T, act_ld, mask, nnz0, row_len_dev, use_k0b ? &sc : nullptr, st,
                             &launches, &fused));
          if (fused) nnz_ready = nnz0;
        }
      }
    }
    if (pt->on) cudaEventRecord(pt->ev[1], st);
    if (mask_only) {
      if (P.E > 0) {
        const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(P.E * bc, 256), 148 * 32);
        k0_kept_bytes_kernel<<<blocks, 256, 0, st>>>(mask, W, bc, P.identity_perm ? nullptr : P.edge_slot, P.E,
                                                    rc.kept_bytes_dev + b0 * P.E);
        SPAI_CUDA(cudaGetLastError()); ++launches;
      }
      ++chunks;
      continue;
    }
    char* scratch = cv.take<char>(0);
    const int64_t left = (reinterpret_cast<char*>(c->ws.base) + c->ws.bytes) - scratch;
    SPAI_TRY(eval_masks(P, plan, mode, dtype, mask, bc, scratch, left, c->sm_count, (double)c->n, res0,
                        (double)c->flops0, rc.alpha, o_rw, o_rs, o_nz, st, pt, &launches, nnz_ready, rc.row_lo,
                        rc.row_hi, rc.partial_only, t_len));
    if (pt->on) cudaEventRecord(pt->ev[4], st);
    if (out_host) {
      if (rc.reward) SPAI_CUDA(cudaMemcpyAsync(rc.reward + b0, o_rw, (size_t)bc * 8, cudaMemcpyDeviceToHost, st));
      if (rc.residual) SPAI_CUDA(cudaMemcpyAsync(rc.residual + b0, o_rs, (size_t)bc * 8, cudaMemcpyDeviceToHost, st));
      if (rc.nnz_m) SPAI_CUDA(cudaMemcpyAsync(rc.nnz_m + b0, o_nz, (size_t)bc * 8, cudaMemcpyDeviceToHost, st));
    }
    ++chunks;
    if (pt->on) {
      SPAI_CUDA(cudaEventSynchronize(pt->ev[4]));
      float t;
      cudaEventElapsedTime(&t, pt->ev[0], pt->ev[1]); ms_masks += t;
      cudaEventElapsedTime(&t, pt->ev[1], pt->ev[2]); ms_tr += t;
      cudaEventElapsedTime(&t, pt->ev[2], pt->ev[3]); ms_rw += t;
      cudaEventElapsedTime(&t, pt->ev[3], pt->ev[4]); ms_fin += t;
    }
    // no sync between chunks: the scratch is reused in stream order
  }
  if (out_host) SPAI_CUDA(cudaStreamSynchronize(st));
  c->last = spai_timing();
  c->last.ms_masks = ms_masks; c->last.ms_transpose = ms_tr; c->last.ms_reward = ms_rw;
  c->last.ms_finalize = ms_fin; c->last.ms_total = ms_masks + ms_tr + ms_rw + ms_fin;
  c->last.launches = launches; c->last.chunks = chunks;
  c->last.algorithmic_bytes = plan.g_bytes_full;     // per fully-kept pattern; callers scale by nnz_m / E
  c->last.compulsory_bytes = (double)B * ((double)W * 4.0 + 8.0);
  c->last.h2d_bytes = h2d_bytes;
  return SPAI_OK;
}

// host entry points must not return while kernels may still read the caller's pinned buffer
static int reward_driver_host(spai_ctx* c, const RewardCall& rc) {
  const int s = reward_driver(c, rc);
  if (s != SPAI_OK && c) {
    DeviceGuard guard(c->device);
    cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(rc.stream));
    cudaGetLastError();
  }
  return s;
}

}  // namespace spai

extern "C" {

int spai_reward_batch_host(spai_ctx* c, const int64_t* actions, int64_t B, int64_t T, int64_t ld,
                           double alpha, int mode, int dtype, double* reward, double* residual,
                           int64_t* nnz_m, void* stream) {
  RewardCall rc;
  rc.src = FROM_ACTIONS_HOST; rc.input = actions; rc.B = B; rc.T = T; rc.ld = ld; rc.alpha = alpha;
  rc.mode = mode; rc.dtype = dtype; rc.reward = reward; rc.residual = residual; rc.nnz_m = nnz_m;
  rc.out_host = true; rc.stream = stream;
  return reward_driver_host(c, rc);
}

int spai_reward_batch_dev(spai_ctx* c, const int64_t* actions, int64_t B, int64_t T, int64_t ld,
                          double alpha, int mode, int dtype, double* reward, double* residual,
                          int64_t* nnz_m, void* stream) {
  RewardCall rc;
  rc.src = FROM_ACTIONS_DEV; rc.input = actions; rc.B = B; rc.T = T; rc.ld = ld; rc.alpha = alpha;
  rc.mode = mode; rc.dtype = dtype; rc.reward = reward; rc.residual = residual; rc.nnz_m = nnz_m;
  rc.stream = stream;
  return reward_driver(c, rc);
}

int spai_reward_batch_host_len(spai_ctx* c, const void* actions, int id_bytes, const int32_t* row_len,
                               int64_t B, int64_t T, int64_t ld, double alpha, int mode, int dtype,
                               double* reward, double* residual, int64_t* nnz_m, void* stream) {
  RewardCall rc;
  rc.src = FROM_ACTIONS_HOST; rc.input = actions; rc.elem = id_bytes; rc.row_len = row_len;
  rc.B = B; rc.T = T; rc.ld = ld; rc.alpha = alpha; rc.mode = mode; rc.dtype = dtype;
  rc.reward = reward; rc.residual = residual; rc.nnz_m = nnz_m; rc.out_host = true; rc.stream = stream;
  return reward_driver_host(c, rc);
}

int spai_reward_batch_dev_len(spai_ctx* c, const void* actions, int id_bytes, const int32_t* row_len,
                              int64_t B, int64_t T, int64_t ld, double alpha, int mode, int dtype,
                              double* reward, double* residual, int64_t* nnz_m, void* stream) {
  RewardCall rc;
  rc.src = FROM_ACTIONS_DEV; rc.input = actions; rc.elem = id_bytes; rc.row_len = row_len;
  rc.B = B; rc.T = T; rc.ld = ld; rc.alpha = alpha; rc.mode = mode; rc.dtype = dtype;
  rc.reward = reward; rc.residual = residual; rc.nnz_m = nnz_m; rc.stream = stream;
  return reward_driver(c, rc);
}

int spai_reward_from_taken_dev(spai_ctx* c, const uint32_t* taken, int64_t B, int64_t words_ld,
                               double alpha, int mode, int dtype, double* reward, double* residual,
                               int64_t* nnz_m, void* stream) {
  if (c && words_ld < c->P.words()) { set_error("taken mask has %lld words per sample, need >= %lld", (long long)words_ld, (long long)c->P.words()); return SPAI_ERR_INVALID; }
  RewardCall rc;
  rc.src = FROM_TAKEN_DEV; rc.input = taken; rc.B = B; rc.T = 0; rc.ld = words_ld; rc.alpha = alpha;
  rc.mode = mode; rc.dtype = dtype; rc.reward = reward; rc.residual = residual; rc.nnz_m = nnz_m;
  rc.stream = stream;
  if (c) { rc.t_hint = c->deletion_hint; c->deletion_hint = 0; }      // the hint serves ONE call (no stale hints)
  return reward_driver(c, rc);
}

int spai_kept_mask_dev(spai_ctx* c, const int64_t* actions, int64_t B, int64_t T, int64_t ld,
                       uint8_t* out, void* stream) {
  if (!out) return SPAI_ERR_INVALID;
  RewardCall rc;
  rc.src = FROM_ACTIONS_DEV; rc.input = actions; rc.B = B; rc.T = T; rc.ld = ld;
  rc.kept_bytes_dev = out; rc.stream = stream;
  return reward_driver(c, rc);
}

int spai_row_index_sets(spai_ctx* c, int64_t row, int64_t* num_j, int64_t* j_host, int64_t* num_i,
                        int64_t* i_host) {
  if (!c || row < 0 || row >= c->n || !num_j || !num_i) return SPAI_ERR_INVALID;
  DeviceGuard guard(c->device);
  SPAI_TRY(ensure_plan(c, SPAI_F32, false, 0));
  const Plan& plan = c->plan[SPAI_F32];
  std::vector<int64_t> J;
  for (int32_t p = c->hp.sptr[row]; p < c->hp.sptr[row + 1]; ++p)
    if (J.empty() || J.back() != c->hp.slot_col[p]) J.push_back(c->hp.slot_col[p]);
  const int64_t cb = plan.cptr_host[row], ce = plan.cptr_host[row + 1];
  std::vector<int32_t> cols(ce - cb);
  if (ce > cb) SPAI_CUDA(cudaMemcpy(cols.data(), plan.c_col + cb, (ce - cb) * 4, cudaMemcpyDeviceToHost));
  std::vector<int64_t> I;
  for (int32_t x : cols) if (I.empty() || I.back() != x) I.push_back(x);
  for (size_t t = 1; t < cols.size(); ++t)
    if (cols[t] < cols[t - 1]) { set_error("plan records of row %lld are not sorted", (long long)row); return SPAI_ERR_CUDA; }
  if (j_host) { if (*num_j < (int64_t)J.size()) return SPAI_ERR_INVALID; std::copy(J.begin(), J.end(), j_host); }
  if (i_host) { if (*num_i < (int64_t)I.size()) return SPAI_ERR_INVALID; std::copy(I.begin(), I.end(), i_host); }
  *num_j = (int64_t)J.size();
  *num_i = (int64_t)I.size();
  return SPAI_OK;
}

int spai_ls_solve_values_host(spai_ctx* c, const int64_t* actions, int64_t T, int dtype, double* m_val,
                              void* stream) {
  if (!c || !m_val || T < 0 || (T && !actions) || (dtype != SPAI_F32 && dtype != SPAI_F64)) {
    set_error("spai_ls_solve_values_host: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  DeviceGuard guard(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SPAI_TRY(ensure_plan(c, dtype, true, st));
  const Plan& plan = c->plan[dtype];
  const Pattern& P = c->P;
  const int64_t W = std::max<int64_t>(P.words(), 1), Bp = 32;
  Arena tmp;
  int64_t* act_dev = nullptr;
  uint32_t *mask = nullptr, *maskT = nullptr;
  double* out = nullptr;
  SPAI_TRY(tmp.alloc(&act_dev, T));
  SPAI_TRY(tmp.alloc(&mask, W));
  SPAI_TRY(tmp.alloc(&maskT, W * Bp));
  SPAI_TRY(tmp.alloc(&out, P.E));
  SPAI_CUDA(cudaMemsetAsync(out, 0, (size_t)std::max<int64_t>(P.E, 1) * 8, st));
  if (T) SPAI_CUDA(cudaMemcpyAsync(act_dev, actions, (size_t)T * 8, cudaMemcpyHostToDevice, st));
  if (P.E > 0) {
    k0_mask_init_kernel<<<(unsigned)std::min<int64_t>(ceil_div(W, 256), 65535), 256, 0, st>>>(mask, W, P.E, 1);
    if (T) k0_mask_clear_kernel<4><<<(unsigned)ceil_div(T, 1024), 256, 0, st>>>(
        act_dev, 1, T, T, P.identity_perm ? nullptr : P.edge_slot, P.E, mask, W, ceil_div(T, 1024), nullptr);
    k0_transpose_kernel<<<dim3((unsigned)ceil_div(W, 32), 1), 256, 0, st>>>(mask, 1, W, maskT, Bp);
    SPAI_CUDA(cudaGetLastError());
    const int64_t warps = (int64_t)c->sm_count * 16;
    const int64_t kmax = P.max_k, qmax = plan.max_q;
    const int64_t wstride = qmax * (kmax + 1) + 2 * kmax + 8;
    const unsigned blocks = (unsigned)(warps / 4);
    int32_t* imap = nullptr;
    SPAI_TRY(tmp.alloc(&imap, warps * 2 * std::max<int64_t>(kmax, 1)));
    if (dtype == SPAI_F32) {
      float* work = nullptr;
      SPAI_TRY(tmp.alloc(&work, warps * wstride));
      k2_ls_solve_kernel<float><<<blocks, 128, 0, st>>>(reinterpret_cast<const Rec32*>(plan.rec_ls), plan.cptr,
          P.sptr, P.slot_edge, plan.r_q, plan.r_diag, P.n, maskT, Bp, work, wstride, imap, 2 * kmax, out);
    } else {
      double* work = nullptr;
      SPAI_TRY(tmp.alloc(&work, warps * wstride));
      k2_ls_solve_kernel<double><<<blocks, 128, 0, st>>>(reinterpret_cast<const Rec64*>(plan.rec_ls), plan.cptr,
          P.sptr, P.slot_edge, plan.r_q, plan.r_diag, P.n, maskT, Bp, work, wstride, imap, 2 * kmax, out);
    }
    SPAI_CUDA(cudaGetLastError());
  }
  SPAI_CUDA(cudaMemcpyAsync(m_val, out, (size_t)P.E * 8, cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));
  return SPAI_OK;
}

int spai_residual_pair_host(int device, int64_t n, int64_t m_nnz, const int64_t* m_row, const int64_t* m_col,
                            const double* m_val, int64_t a_nnz, const int64_t* a_row, const int64_t* a_col,
                            const double* a_val, int dtype, double* residual_out, int64_t* m_nnz_out) {
  if (n <= 0 || m_nnz < 0 || a_nnz < 0 || !residual_out || (dtype != SPAI_F32 && dtype != SPAI_F64)) {
    set_error("spai_residual_pair_host: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  int ndev = 0;
  SPAI_TRY(spai_device_count(&ndev));
  if (device < 0 || device >= ndev) { set_error("device %d not available", device); return SPAI_ERR_CUDA; }
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  int sm = 148;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) sm = prop.multiProcessorCount;
  Arena ar;
  HostCsr ha;
  CsrA A;
  SPAI_TRY(build_host_csr(n, a_nnz, a_row, a_col, a_val, ha));
  SPAI_TRY(upload_csr(ar, ha, n, a_nnz, A));
  return pair_residual(sm, n, m_nnz, m_row, m_col, m_val, ha, A, dtype, residual_out, m_nnz_out, 0);
}

int spai_sample_step_dev(spai_ctx* c, const float* logits, int64_t logits_ld, int64_t A, uint32_t* taken,
                         int64_t words_ld, const float* uniforms, uint8_t* done, int64_t B,
                         int64_t* action, float* prob, void* stream) {
  if (!c || !logits || !taken || !uniforms || !done || !action || !prob || B < 0 || A <= 0 ||
      words_ld < (A + 31) / 32 || (logits_ld != 0 && logits_ld < A)) {
    set_error("spai_sample_step_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  if (B == 0) return SPAI_OK;
  DeviceGuard guard(c->device);
  k4_sample_kernel<<<(unsigned)B, K4_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, logits_ld, A, taken, words_ld, uniforms, done, action, prob);
  SPAI_CUDA(cudaGetLastError());
  return SPAI_OK;
}

int spai_reward_rows_dev(spai_ctx* c, const int64_t* actions, int64_t B, int64_t T, int64_t ld, int mode,
                         int dtype, int64_t row_begin, int64_t row_end, double* res2_partial, int64_t* nnz_m,
                         void* stream) {
  if (!c || row_begin < 0 || row_end < row_begin || row_end > c->n || !res2_partial) {
    set_error("spai_reward_rows_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  RewardCall rc;
  rc.src = FROM_ACTIONS_DEV; rc.input = actions; rc.B = B; rc.T = T; rc.ld = ld; rc.mode = mode; rc.dtype = dtype;
  rc.residual = res2_partial; rc.nnz_m = nnz_m; rc.row_lo = row_begin; rc.row_hi = row_end; rc.partial_only = true;
  rc.stream = stream;
  return reward_driver(c, rc);
}

int spai_finalize_rewards_dev(spai_ctx* c, const double* res2, const int64_t* nnz_m, int64_t B, double alpha,
                              int dtype, double* reward, double* residual, void* stream) {
  if (!c || B < 0 || (B && (!res2 || !nnz_m)) || (dtype != SPAI_F32 && dtype != SPAI_F64)) {
    set_error("spai_finalize_rewards_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  if (B == 0) return SPAI_OK;
  DeviceGuard guard(c->device);
  k3_finalize_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      res2, 1, B, B, nullptr, 0.0, reinterpret_cast<const long long*>(nnz_m), (double)c->n, c->res0[dtype],
      (double)c->flops0, alpha, reward, residual, nullptr, nullptr, 0u, 0);
  SPAI_CUDA(cudaGetLastError());
  return SPAI_OK;
}

int spai_pack_taken_dev(spai_ctx* c, const float* keys, int64_t keys_ld, int64_t A, int64_t B, uint32_t* taken,
                        int64_t words_ld, int32_t* length, void* stream) {
  if (!c || !keys || !taken || !length || A <= 0 || B < 0 || keys_ld < A || words_ld < (A + 31) / 32) {
    set_error("spai_pack_taken_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  if (B == 0) return SPAI_OK;
  DeviceGuard guard(c->device);
  k4_pack_taken_kernel<<<(unsigned)B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(keys, keys_ld, A, taken,
                                                                                       words_ld, length);
  SPAI_CUDA(cudaGetLastError());
  return SPAI_OK;
}

}  // extern "C"
