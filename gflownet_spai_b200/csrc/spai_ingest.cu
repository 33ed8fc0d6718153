// C-ABI host side of the device ingest kernels (k5_ingest.cuh); declarations in include/spai_b200.h.
#include <algorithm>
#include <vector>

#include "k5_ingest.cuh"
#include "spai_internal.cuh"

using namespace spai;

namespace {

struct Guard {
  int prev = -1;
  explicit Guard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
  ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};
struct Tmp {                       // stream-ordered scratch, released after a sync
  std::vector<void*> p;
  template <typename T> int get(T** out, int64_t count) {
    void* q = nullptr;
    const cudaError_t e = cudaMalloc(&q, (size_t)std::max<int64_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) { set_error("ingest: cudaMalloc failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return SPAI_ERR_NOMEM; }
    p.push_back(q);
    *out = reinterpret_cast<T*>(q);
    return SPAI_OK;
  }
  ~Tmp() { for (void* q : p) cudaFree(q); }
};

// out[0..n] = exclusive prefix sums of in[0..n)
template <typename OutT>
int scan(const int* in, int64_t n, OutT* out, Tmp& tmp, cudaStream_t st) {
  const int64_t tiles = std::max<int64_t>(1, ceil_div(n, (int64_t)K5_SCAN_TILE));
  long long* sums = nullptr;
  SPAI_TRY(tmp.get(&sums, tiles + 1));
  k5_scan_tile_kernel<<<(unsigned)tiles, K5_SCAN_T, 0, st>>>(in, n, sums);
  k5_scan_tops_kernel<<<1, 32, 0, st>>>(sums, tiles);
  k5_scan_apply_kernel<OutT><<<(unsigned)tiles, K5_SCAN_T, 0, st>>>(in, n, sums, tiles, out);
  SPAI_CUDA(cudaGetLastError());
  return SPAI_OK;
}

int check_flag(const int* flag_dev, cudaStream_t st, const char* what, int status) {
  int h = 0;
  SPAI_CUDA(cudaMemcpyAsync(&h, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));
  if (h) { set_error("%s", what); return status; }
  return SPAI_OK;
}

int max_row_len(const int32_t* ptr_dev, int64_t n, cudaStream_t st, int* out) {     // small helper: host max of diff(ptr)
  std::vector<int32_t> h((size_t)n + 1);
  SPAI_CUDA(cudaMemcpyAsync(h.data(), ptr_dev, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));
  int m = 0;
  for (int64_t i = 0; i < n; ++i) m = std::max(m, h[i + 1] - h[i]);
  *out = m;
  return SPAI_OK;
}

}  // namespace

extern "C" {

int spai_ingest_coo_to_csr_dev(int device, int64_t n, int64_t nnz, const int64_t* row_dev, const int64_t* col_dev,
                               const double* val_dev, int32_t* ptr_dev, int32_t* col_out_dev, double* val_out_dev,
                               int64_t* nnz_out_host, void* stream) {
  if (n <= 0 || nnz < 0 || nnz >= ((int64_t)1 << 31) || n >= ((int64_t)1 << 31) || !ptr_dev || !nnz_out_host ||
      (nnz && (!row_dev || !col_dev || !val_dev || !col_out_dev || !val_out_dev))) {
    set_error("spai_ingest_coo_to_csr_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  Guard g(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Tmp tmp;
  int *cnt = nullptr, *cursor = nullptr, *ucnt = nullptr, *bad = nullptr;
  int64_t* ptr64 = nullptr;
  unsigned long long* keyed = nullptr;
  SPAI_TRY(tmp.get(&cnt, n)); SPAI_TRY(tmp.get(&cursor, n)); SPAI_TRY(tmp.get(&ucnt, n)); SPAI_TRY(tmp.get(&bad, 1));
  SPAI_TRY(tmp.get(&ptr64, n + 1)); SPAI_TRY(tmp.get(&keyed, nnz));
  SPAI_CUDA(cudaMemsetAsync(cnt, 0, (size_t)n * 4, st));
  SPAI_CUDA(cudaMemsetAsync(cursor, 0, (size_t)n * 4, st));
  SPAI_CUDA(cudaMemsetAsync(ucnt, 0, (size_t)n * 4, st));
  SPAI_CUDA(cudaMemsetAsync(bad, 0, 4, st));
  const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nnz, 256), 148 * 16));
  k5_coo_count_kernel<<<blocks, 256, 0, st>>>(row_dev, nnz, n, cnt, bad);
  SPAI_TRY(scan<int64_t>(cnt, n, ptr64, tmp, st));
  k5_coo_scatter_kernel<<<blocks, 256, 0, st>>>(row_dev, col_dev, nnz, n, ptr64, cursor, keyed, bad);
  const unsigned rb = (unsigned)ceil_div(n, (int64_t)K5_WARPS);
  k5_coo_rowsort_kernel<<<rb, K5_WARPS * 32, 0, st>>>(n, ptr64, keyed, ucnt);
  SPAI_TRY(scan<int>(ucnt, n, ptr_dev, tmp, st));
  k5_coo_emit_kernel<<<rb, K5_WARPS * 32, 0, st>>>(n, ptr64, keyed, val_dev, ptr_dev, col_out_dev, val_out_dev);
  SPAI_CUDA(cudaGetLastError());
  SPAI_TRY(check_flag(bad, st, "spai_ingest_coo_to_csr_dev: an entry lies outside [0, n) x [0, n)", SPAI_ERR_INVALID));
  int32_t total = 0;
  SPAI_CUDA(cudaMemcpyAsync(&total, ptr_dev + n, 4, cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));
  *nnz_out_host = total;
  return SPAI_OK;
}

int spai_ingest_spgemm_count_dev(int device, int64_t n, const int32_t* a_ptr, const int32_t* a_col,
                                 const int32_t* b_ptr, const int32_t* b_col, int32_t* c_ptr_dev,
                                 int64_t* c_nnz_host, void* stream) {
  if (n <= 0 || n >= ((int64_t)1 << 31) || !a_ptr || !a_col || !b_ptr || !b_col || !c_ptr_dev || !c_nnz_host) {
    set_error("spai_ingest_spgemm_count_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  Guard g(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int blen = 0;
  SPAI_TRY(max_row_len(b_ptr, n, st, &blen));
  if (blen > K5_LIST) { set_error("spgemm: a row of B has %d entries (max %d)", blen, K5_LIST); return SPAI_ERR_UNSUPPORTED; }
  Tmp tmp;
  int *cnt = nullptr, *ovf = nullptr;
  SPAI_TRY(tmp.get(&cnt, n)); SPAI_TRY(tmp.get(&ovf, 1));
  SPAI_CUDA(cudaMemsetAsync(ovf, 0, 4, st));
  SPAI_CUDA(cudaFuncSetAttribute(k5_spgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5_row_smem()));
  k5_spgemm_kernel<false><<<(unsigned)ceil_div(n, (int64_t)K5_WARPS), K5_WARPS * 32, k5_row_smem(), st>>>(
      n, a_ptr, a_col, nullptr, b_ptr, b_col, nullptr, cnt, nullptr, nullptr, nullptr, ovf);
  SPAI_CUDA(cudaGetLastError());
  SPAI_TRY(scan<int>(cnt, n, c_ptr_dev, tmp, st));
  SPAI_TRY(check_flag(ovf, st, "spgemm: a row of the product has more than 512 entries", SPAI_ERR_UNSUPPORTED));
  int32_t total = 0;
  SPAI_CUDA(cudaMemcpyAsync(&total, c_ptr_dev + n, 4, cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));
  *c_nnz_host = total;
  return SPAI_OK;
}

int spai_ingest_spgemm_fill_dev(int device, int64_t n, const int32_t* a_ptr, const int32_t* a_col, const double* a_val,
                                const int32_t* b_ptr, const int32_t* b_col, const double* b_val,
                                const int32_t* c_ptr_dev, int32_t* c_col_dev, double* c_val_dev, void* stream) {
  if (n <= 0 || !a_ptr || !a_col || !b_ptr || !b_col || !c_ptr_dev || !c_col_dev || !c_val_dev) {
    set_error("spai_ingest_spgemm_fill_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  Guard g(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Tmp tmp;
  int* ovf = nullptr;
  SPAI_TRY(tmp.get(&ovf, 1));
  SPAI_CUDA(cudaMemsetAsync(ovf, 0, 4, st));
  SPAI_CUDA(cudaFuncSetAttribute(k5_spgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5_row_smem()));
  k5_spgemm_kernel<true><<<(unsigned)ceil_div(n, (int64_t)K5_WARPS), K5_WARPS * 32, k5_row_smem(), st>>>(
      n, a_ptr, a_col, a_val, b_ptr, b_col, b_val, nullptr, c_ptr_dev, c_col_dev, c_val_dev, ovf);
  SPAI_CUDA(cudaGetLastError());
  return check_flag(ovf, st, "spgemm: a row of the product has more than 512 entries", SPAI_ERR_UNSUPPORTED);
}

int spai_ingest_csr_drop_zeros_dev(int device, int64_t n, const int32_t* ptr, const int32_t* col, const double* val,
                                   int32_t* out_ptr_dev, int32_t* out_col_dev, double* out_val_dev,
                                   int64_t* nnz_out_host, void* stream) {
  if (n <= 0 || !ptr || !out_ptr_dev || !nnz_out_host) {
    set_error("spai_ingest_csr_drop_zeros_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  Guard g(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Tmp tmp;
  int* cnt = nullptr;
  SPAI_TRY(tmp.get(&cnt, n));
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
  k5_count_nonzero_kernel<<<blocks, 256, 0, st>>>(n, ptr, val, cnt);
  SPAI_TRY(scan<int>(cnt, n, out_ptr_dev, tmp, st));
  if (out_col_dev && out_val_dev) k5_copy_nonzero_kernel<<<blocks, 256, 0, st>>>(n, ptr, col, val, out_ptr_dev, out_col_dev, out_val_dev);
  SPAI_CUDA(cudaGetLastError());
  int32_t total = 0;
  SPAI_CUDA(cudaMemcpyAsync(&total, out_ptr_dev + n, 4, cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));
  *nnz_out_host = total;
  return SPAI_OK;
}

int spai_ingest_superset_dev(int device, int64_t n, const int32_t* a_ptr, const int32_t* a_col, int k, int max_power,
                             int order, int64_t* s_ptr_dev, int64_t* s_row_dev, int64_t* s_col_dev,
                             int64_t* s_nnz_host, void* stream) {
  if (n <= 0 || n >= ((int64_t)1 << 31) || !a_ptr || !a_col || k <= 0 || k > K5_LIST || max_power < 0 ||
      (order != 0 && order != 1) || !s_ptr_dev || !s_nnz_host) {
    set_error("spai_ingest_superset_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  Guard g(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int alen = 0;
  SPAI_TRY(max_row_len(a_ptr, n, st, &alen));
  if (alen > K5_LIST) { set_error("superset: a row of A has %d entries (max %d)", alen, K5_LIST); return SPAI_ERR_UNSUPPORTED; }
  Tmp tmp;
  int *scnt = nullptr, *scol = nullptr, *ovf = nullptr;
  SPAI_TRY(tmp.get(&scnt, n)); SPAI_TRY(tmp.get(&scol, n * k)); SPAI_TRY(tmp.get(&ovf, 1));
  SPAI_CUDA(cudaMemsetAsync(ovf, 0, 4, st));
  SPAI_CUDA(cudaFuncSetAttribute(k5_superset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5_row_smem()));
  k5_superset_kernel<<<(unsigned)ceil_div(n, (int64_t)K5_WARPS), K5_WARPS * 32, k5_row_smem(), st>>>(
      n, a_ptr, a_col, k, max_power, order, scnt, scol, ovf);
  SPAI_CUDA(cudaGetLastError());
  SPAI_TRY(scan<int64_t>(scnt, n, s_ptr_dev, tmp, st));
  SPAI_TRY(check_flag(ovf, st, "superset: a row reaches more than 512 columns within max_power steps", SPAI_ERR_UNSUPPORTED));
  int64_t total = 0;
  SPAI_CUDA(cudaMemcpyAsync(&total, s_ptr_dev + n, 8, cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));
  *s_nnz_host = total;
  if (s_row_dev && s_col_dev) {              // the caller sized them for n * k entries (upper bound)
    k5_pattern_compact_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n * k, 256), 148 * 32), 256, 0, st>>>(
        n, k, scnt, scol, s_ptr_dev, s_row_dev, s_col_dev);
    SPAI_CUDA(cudaGetLastError());
    SPAI_CUDA(cudaStreamSynchronize(st));
  }
  return SPAI_OK;
}

int spai_ingest_neumann_dev(int device, int64_t n, const int32_t* a_ptr, const int32_t* a_col, const double* a_val,
                            const int64_t* s_ptr_dev, const int64_t* s_col_dev, int terms, double* omega_host,
                            double* val_out_dev, void* stream) {
  if (n <= 0 || !a_ptr || !a_col || !a_val || !s_ptr_dev || !s_col_dev || terms < 1 || !val_out_dev) {
    set_error("spai_ingest_neumann_dev: invalid arguments");
    return SPAI_ERR_INVALID;
  }
  Guard g(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int alen = 0;
  SPAI_TRY(max_row_len(a_ptr, n, st, &alen));
  if (alen > K5_LIST) { set_error("neumann: a row of A has %d entries (max %d)", alen, K5_LIST); return SPAI_ERR_UNSUPPORTED; }
  Tmp tmp;
  unsigned long long* mx = nullptr;
  int* ovf = nullptr;
  SPAI_TRY(tmp.get(&mx, 1)); SPAI_TRY(tmp.get(&ovf, 1));
  SPAI_CUDA(cudaMemsetAsync(mx, 0, 8, st));
  SPAI_CUDA(cudaMemsetAsync(ovf, 0, 4, st));
  k5_row_abs_max_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 8), 256, 0, st>>>(n, a_ptr, a_val, mx);
  unsigned long long bits = 0;
  SPAI_CUDA(cudaMemcpyAsync(&bits, mx, 8, cudaMemcpyDeviceToHost, st));
  SPAI_CUDA(cudaStreamSynchronize(st));
  double mrow;
  memcpy(&mrow, &bits, 8);
  if (!(mrow > 0.0)) { set_error("neumann: the matrix has no non-zero row"); return SPAI_ERR_INVALID; }
  const double omega = 1.0 / mrow;
  if (omega_host) *omega_host = omega;
  SPAI_CUDA(cudaFuncSetAttribute(k5_neumann_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5_neumann_smem()));
  k5_neumann_kernel<<<(unsigned)ceil_div(n, (int64_t)K5_WARPS), K5_WARPS * 32, k5_neumann_smem(), st>>>(
      n, a_ptr, a_col, a_val, s_ptr_dev, s_col_dev, terms, omega, val_out_dev, ovf);
  SPAI_CUDA(cudaGetLastError());
  return check_flag(ovf, st, "neumann: a row of the power series has more than 512 entries", SPAI_ERR_UNSUPPORTED);
}

}  // extern "C"
