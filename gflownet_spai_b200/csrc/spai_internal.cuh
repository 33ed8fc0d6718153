// Internal declarations shared by the kernels and the C-ABI host code.
// sm_100a only. See DESIGN.md for the data layout.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "spai_b200.h"

namespace spai {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);

#define SPAI_CUDA(call)                                                               \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      ::spai::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                  \
                        cudaGetErrorString(e_));                                      \
      return SPAI_ERR_CUDA;                                                           \
    }                                                                                 \
  } while (0)

#define SPAI_TRY(call)                 \
  do {                                 \
    int s_ = (call);                   \
    if (s_ != SPAI_OK) return s_;      \
  } while (0)

// ---------------------------------------------------------------- plan records
// One record per gathered entry A[c_e, x] of row i ("contribution"): slot e of
// the row's candidate list reaches output column x. Records of a row are sorted
// by (x, e); records with the same x form one output segment (one entry of the
// row of M*A). 16 bytes, so a run of records is a legal cp.async.bulk source.
//   flags: bit31 END  last record of its output segment (sign bit: one ISETP)
//          bit30 NEXT_DIAG (END records only) the FOLLOWING segment's column is
//                the row index, i.e. the entry that carries the "-1" of "-I"
//          bits 14..29 s  output segment index inside the row (< 65536)
//          bits 0..13  e  slot index inside the row (< 16384)
constexpr uint32_t F_END = 0x80000000u;
constexpr uint32_t F_NEXT_DIAG = 0x40000000u;
constexpr int MAX_ROW_SLOTS = 16384;
constexpr int MAX_ROW_UNION = 65536;

__host__ __device__ __forceinline__ uint32_t rec_e(uint32_t flags) { return flags & 0x3fffu; }
__host__ __device__ __forceinline__ uint32_t rec_s(uint32_t flags) { return (flags >> 14) & 0xffffu; }

struct alignas(16) Rec32 {
  uint32_t ebit;   // 1u << e when the row has <= 32 slots, else 0
  uint32_t flags;
  float w;         // fl32(m_e * a): copy mode adds this when slot e is kept
  float a;         // A[c_e, x]: ls mode tile entry
};
struct alignas(16) Rec64 {
  uint32_t ebit;
  uint32_t flags;
  double v;        // w (copy plan) or a (ls plan)
};
// Row header (16 bytes, bulk-copied next to the records of its tile).
struct alignas(16) RowHdr {
  int32_t cnt;     // records of the row
  int32_t sp;      // first slot of the row == its bit offset in the kept-mask
  int32_t k;       // candidate slots of the row
  int32_t flags;   // bit0: the row's first output segment is the diagonal
};
template <typename T> struct RecOf;
template <> struct RecOf<float> { using type = Rec32; };
template <> struct RecOf<double> { using type = Rec64; };

__device__ __forceinline__ float rec_w(const Rec32& r) { return r.w; }
__device__ __forceinline__ double rec_w(const Rec64& r) { return r.v; }
__device__ __forceinline__ float rec_a(const Rec32& r) { return r.a; }
__device__ __forceinline__ double rec_a(const Rec64& r) { return r.v; }

// ---------------------------------------------------------------- device data
struct Pattern {            // initial matrix: candidate superset S in slot order
  int64_t n = 0, E = 0;
  int32_t* sptr = nullptr;        // [n+1] slot offsets; slots sorted by (row, col, edge id)
  int32_t* slot_col = nullptr;    // [E]
  int32_t* slot_edge = nullptr;   // [E] slot -> edge id (caller order)
  int32_t* edge_slot = nullptr;   // [E] edge id -> slot
  float* slot_val32 = nullptr;    // [E]
  double* slot_val64 = nullptr;   // [E]
  int32_t* dup_start = nullptr;   // [ndup] first slot of every coordinate group with > 1 slot
  int32_t* dup_len = nullptr;     // [ndup]
  int64_t ndup = 0;
  int64_t init_nnz = 0;           // distinct coordinates
  int max_k = 0;
  bool identity_perm = false;
  int64_t words() const { return (E + 31) / 32; }
};

struct CsrA {               // original matrix, coalesced
  int64_t n = 0, nnz = 0, nnz_stored = 0;
  int32_t* ptr = nullptr;   // [n+1]
  int32_t* col = nullptr;   // [nnz]
  float* val32 = nullptr;
  double* val64 = nullptr;
};

constexpr int LS_NCLASS = 8;     // last class = generic fallback
constexpr int LS_GENERIC = LS_NCLASS - 1;
constexpr int GRAM_NCLASS = 3;

struct Plan {
  int dtype = -1;
  int64_t n = 0;
  int64_t nc = 0;                 // total contributions
  int64_t* cptr = nullptr;        // [n+1] device
  std::vector<int64_t> cptr_host;
  int32_t* c_col = nullptr;       // [nc] output column of every record
  void* rec_copy = nullptr;       // Rec32 / Rec64(w)
  void* rec_ls = nullptr;         // Rec32 (same array) / Rec64(a)
  int32_t* r_q = nullptr;         // [n] |I_i|
  int32_t* r_diag = nullptr;      // [n] segment index of column i, or -1
  RowHdr* rhdr = nullptr;         // [n]
  void* row_base = nullptr;       // [n] T: row residual^2 with every candidate kept (copy mode)
  double* row_base_ls = nullptr;  // [n] the same for ls mode (re-solved row)
  int64_t rows_missing_diag = 0;  // rows with r_diag < 0 (each adds 1 to ||.||^2)
  int max_q = 0;
  int max_k = 0;
  // copy-kernel tiling: tile t = rows [tile_row[t], tile_row[t+1])
  int32_t* tile_row = nullptr;    // [ntiles+1] device
  std::vector<int32_t> tile_row_host;
  int ntiles = 0;
  // ls-kernel row classes
  int32_t* class_rows[LS_NCLASS] = {};
  int64_t class_count[LS_NCLASS] = {};
  std::vector<int32_t> class_rows_host[LS_NCLASS];   // sorted ascending (row-range evaluation)
  std::vector<int64_t> missing_prefix;               // [n+1] prefix count of rows without a diagonal slot
  int64_t generic_max_q = 0, generic_max_k = 0;
  // ls_gram mode (built on first use): rows solved through the semi-normal equations
  // (gram class 0: k <= 8; 1: 8 < k <= 16; 2: 16 < k <= 32, fp64 only); every other row keeps
  // its QR class
  bool gram_ready = false;
  unsigned char* gram[GRAM_NCLASS] = {};             // [gram_count][GramGeom::RB]
  int64_t gram_count[GRAM_NCLASS] = {};
  std::vector<int32_t> gram_rows_host[GRAM_NCLASS];  // sorted ascending
  int32_t* rest_rows[LS_NCLASS] = {};
  int64_t rest_count[LS_NCLASS] = {};
  std::vector<int32_t> rest_rows_host[LS_NCLASS];
  // table-lookup copy kernel (K3t, rows with <= 8 candidates only; built on first use)
  bool lut_ready = false;
  mutable bool tables_on = false;                    // set per call: tables serve calls with >= 64 trajectories, so the
                                                     // kernel (and the rounding) a trajectory gets depends on the call alone
  void* lut = nullptr;                               // T[n][256]
  bool lut_ls_ready = false;                         // the same for ls_gram mode (least-squares residuals)
  void* lut_ls = nullptr;
  bool lut_qr_ready = false;                         // ls mode: filled by the Householder kernel, always f64[n][256]
  double* lut_qr = nullptr;
  // tensor-core copy kernel (K3m, fp32, rows with <= 32 candidates, no repeated coordinates; built on first use)
  bool mma_ready = false;
  bool mma_unavailable = false;
  // two row classes: [0] rows with <= 16 candidates (N = K = 16), [1] rows with 17..32 (N = K = 32); each class has its own
  // sorted row list, contiguous records and compact headers {sp, k}
  unsigned char* mma_rec[2] = {nullptr, nullptr};    // [count][K3mGeom::RB]
  int2* mma_hdr[2] = {nullptr, nullptr};             // [count] {first slot, candidates}
  int64_t mma_count[2] = {0, 0};
  std::vector<int32_t> mma_rows_host[2];             // ascending row ids of the class
  int mma_split = 0;                                 // bf16 terms per entry of L
  // deletion-driven copy kernel (K3s, built on first use)
  bool sparse_ready = false;
  bool sparse_unavailable = false;                   // a row of A or of the pattern exceeds the SlotMeta fields
  void* sl_meta = nullptr;                           // SlotMeta[E]: row slot range + the slot's (f, w) list
  void* sl_rec = nullptr;                            // Pair<T>[nc] slot-major
  std::vector<double> base_prefix;                   // [n+1] prefix sums of row_base
  const int32_t* sptr_host = nullptr;                // [n+1] host copy of the slot offsets (owned by the context)
  const int32_t* a_ptr = nullptr;                    // CSR of A (owned by the context)
  const int32_t* a_col = nullptr;
  double g_bytes_full = 0;        // SURVEY §8d G with every candidate kept (bytes / pattern)
  int64_t bytes = 0;
};

// ---------------------------------------------------------------- small helpers
inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

}  // namespace spai
