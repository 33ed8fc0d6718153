/*
 * spai_b200.h — C ABI of the B200-native SPAI reward path (libspai_b200.so).
 *
 * The reference (tonylizza/gflownet-spai) is pure Python and has no FFI; its
 * seam is the Python `Env` protocol. Each entry point below cites the reference
 * interface it replaces (file:line relative to the reference checkout). The
 * host-side mirror of that protocol is gflownet_spai_b200/env.py; the ctypes
 * binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross the boundary;
 *   - every function returns an int status (SPAI_OK == 0); no exceptions;
 *     spai_last_error() returns a thread-local message for the last failure;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - pointers named *_dev are device pointers on the context's device,
 *     *_host are host pointers (pinned host memory makes copies asynchronous);
 *   - calls enqueue on `stream`; only functions that fill HOST outputs
 *     synchronise that stream before returning (host entry points also synchronise
 *     it before returning an error, so the caller's buffers are never read afterwards);
 *   - one thread per context at a time; calls on different streams are ordered after the
 *     context's one-time table builds by an internal event.
 */
#ifndef SPAI_B200_H_
#define SPAI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPAI_ABI_VERSION 2

enum spai_status {
  SPAI_OK = 0,
  SPAI_ERR_INVALID = 1,      /* bad argument (the reference raises ValueError) */
  SPAI_ERR_CUDA = 2,         /* CUDA runtime error, see spai_last_error()      */
  SPAI_ERR_UNSUPPORTED = 3,  /* shape outside the supported envelope           */
  SPAI_ERR_NOMEM = 4
};

enum spai_mode {
  SPAI_MODE_COPY = 0, /* surviving entries keep their values: what the reference
                         computes (gflownet/utils.py:331-337)                  */
  SPAI_MODE_LS = 1,   /* re-solve every row's least-squares problem on its
                         pattern (BASELINE.json north star; no reference code):
                         Householder QR on the gathered tile                    */
  SPAI_MODE_LS_GRAM = 2 /* the same least-squares residual through the semi-normal
                         equations: the row's Gram matrix is formed once per
                         context, a (row, pattern) solve is a masked k x k LDL^T
                         (rows with <= 32 candidates; when every row has <= 8 the
                         256 possible solves per row are tabulated once); other
                         rows and ill-conditioned tiles go to the Householder
                         kernels                                                 */
};

enum spai_dtype { SPAI_F32 = 0, SPAI_F64 = 1 };

typedef struct spai_ctx spai_ctx; /* opaque */

typedef struct spai_info {
  int64_t n;             /* matrix size                       preconditioner.py:13 */
  int64_t num_edges;     /* E: entries of the initial matrix in caller order (:23) */
  int64_t init_nnz;      /* distinct coordinates (coalesced count)           (:14) */
  int64_t num_actions;   /* init_nnz + 1                                     (:16) */
  int64_t a_nnz_stored;  /* stored values of the original matrix             (:71) */
  int64_t a_nnz;         /* coalesced nnz of the original matrix                   */
  int64_t orig_flops;    /* 2 * a_nnz_stored * n                          (:29,:72) */
  double orig_residual_f32; /* ||A0 A0 - I||_F, fp32 products (reference)    (:28) */
  double orig_residual_f64; /* same, fp64 products                                 */
  int64_t contributions; /* sum over rows of gathered A entries (plan size)        */
  int32_t max_row_slots; /* max candidates per row (k)                             */
  int32_t max_row_union; /* max |I_i| (q)                                          */
  int32_t has_duplicates;/* initial matrix has repeated coordinates                */
  int32_t device;
  int64_t rows_missing_diag; /* rows whose union index set misses the diagonal     */
  int64_t ls_class_rows[8];  /* rows per Householder kernel class of SPAI_MODE_LS, by
                                (max candidates, max union size): [0..2] k<=8 with
                                |I|<=18 / 20 / 40, [3..4] k<=16 with |I|<=52 / 64,
                                [5..6] k<=32 with |I|<=52 / 64, [7] generic kernel
                                (anything larger, rows with repeated coordinates)    */
  int64_t device_bytes;      /* bytes of device memory held by the context         */
} spai_info;

int spai_abi_version(void);
const char* spai_last_error(void);
int spai_device_count(int* count);

/* PreconditionerEnv.__init__ (preconditioner.py:12-29): captures the edge table
 * in the caller's COO order (action k == k-th entry as given, explicit zeros and
 * repeated coordinates included), coalesces the original matrix to CSR, builds
 * the per-row gather plans on the device (kernel K1) and the baseline constants
 * res0 = ||A0 A0 - I||_F and flops0. All inputs are HOST arrays; values are fp64
 * and are rounded to fp32 for the fp32 path exactly as the reference's
 * `.float()` does (preconditioner.py:25). */
int spai_ctx_create(int device, int64_t n,
                    int64_t num_edges, const int64_t* edge_row_host,
                    const int64_t* edge_col_host, const double* edge_val_host,
                    int64_t a_nnz, const int64_t* a_row_host, const int64_t* a_col_host,
                    const double* a_val_host, spai_ctx** out);
void spai_ctx_destroy(spai_ctx* ctx);
int spai_ctx_info(const spai_ctx* ctx, spai_info* out);

/* Cap on the scratch memory (masks + partial sums) one reward call may use;
 * larger batches are processed in trajectory chunks. Default 16 GiB. */
int spai_ctx_set_workspace_limit(spai_ctx* ctx, int64_t bytes);

/* Upper bound on the number of edges one trajectory removes, for the entry point that
 * carries no action list (spai_reward_from_taken_dev): lets the library pick the
 * deletion-driven kernel for short trajectories. 0 (default) = unknown. The bound only
 * selects a kernel; results do not depend on it. It applies to the NEXT
 * spai_reward_from_taken_dev call only and is cleared by it (no stale hints). */
int spai_ctx_set_deletion_hint(spai_ctx* ctx, int64_t max_deletions);

/* PreconditionerEnv.update (preconditioner.py:32-52) for a whole batch:
 * actions int64[B, T] with leading dimension `ld` (>= T), -1 padded; ids outside
 * [0, E) — the terminal id included — match no edge (gflownet/utils.py:323).
 * Outputs (each may be NULL): reward f64[B] (preconditioner.py:64), residual
 * f64[B] (:90), nnz_m i64[B] (:71).
 *   reward = 1000 * (alpha*(1 - res/res0) + (1-alpha)*(1 - flops/flops0)).
 * `res0` is the fp32-product baseline for SPAI_F32 and the fp64 one for SPAI_F64.
 * *_host variant: host in / host out, copies on `stream`, synchronises it. */
int spai_reward_batch_host(spai_ctx* ctx, const int64_t* actions_host, int64_t B, int64_t T,
                           int64_t ld, double alpha, int mode, int dtype,
                           double* reward_host, double* residual_host, int64_t* nnz_m_host,
                           void* stream);
int spai_reward_batch_dev(spai_ctx* ctx, const int64_t* actions_dev, int64_t B, int64_t T,
                          int64_t ld, double alpha, int mode, int dtype,
                          double* reward_dev, double* residual_dev, int64_t* nnz_m_dev,
                          void* stream);

/* The same update (preconditioner.py:32-52) for callers that KNOW the valid length of
 * every row — the reference's Log does: a sample stops being logged once `done`
 * (gflownet/log.py:84-87), so row b of `complete_actions` (gflownet.py:181) holds
 * row_len[b] ids followed by -1 padding only.
 *   actions   ids of `id_bytes` bytes each: 8 = int64 (the reference's tensor), 4 = int32
 *             (device-resident samplers; halves the bytes the mask kernels stream);
 *   row_len   i32[B], may be NULL (= T for every row). Entries at positions >= row_len[b] are
 *             NOT read: the caller vouches that they are padding. With it no thread, host or
 *             device, ever touches the padding (49 % of the cfg2 batch), and the host entry
 *             needs no scan of the input at all.
 *             _host_len: row_len is a HOST array; _dev_len: a DEVICE array.
 * Ids inside the valid prefix keep the exact semantics of the plain entry points (-1 and ids
 * outside [0, E) match no edge, duplicates collapse). int32 host actions require row_len. */
int spai_reward_batch_host_len(spai_ctx* ctx, const void* actions_host, int id_bytes,
                               const int32_t* row_len_host, int64_t B, int64_t T, int64_t ld,
                               double alpha, int mode, int dtype, double* reward_host,
                               double* residual_host, int64_t* nnz_m_host, void* stream);
int spai_reward_batch_dev_len(spai_ctx* ctx, const void* actions_dev, int id_bytes,
                              const int32_t* row_len_dev, int64_t B, int64_t T, int64_t ld,
                              double alpha, int mode, int dtype, double* reward_dev,
                              double* residual_dev, int64_t* nnz_m_dev, void* stream);

/* gflownet/utils.py:315-323 on its own: kept-edge mask of every trajectory in
 * EDGE (caller) order, one byte per edge: out_dev u8[B, E]. The bit-exact
 * artefact compared with the reference's `remaining_edges_mask`. */
int spai_kept_mask_dev(spai_ctx* ctx, const int64_t* actions_dev, int64_t B, int64_t T,
                       int64_t ld, uint8_t* out_dev, void* stream);

/* Reward from an already-built "taken" bitmask in EDGE order (the sampler's
 * state, bit e of word e/32 set == edge e removed): taken_dev u32[B, words],
 * words = ceil(num_actions / 32). Skips the actions->mask step. */
int spai_reward_from_taken_dev(spai_ctx* ctx, const uint32_t* taken_dev, int64_t B,
                               int64_t words_ld, double alpha, int mode, int dtype,
                               double* reward_dev, double* residual_dev, int64_t* nnz_m_dev,
                               void* stream);

/* Gathered index sets of the plan, for parity tests: for row i,
 * J_i = candidate columns (sorted), I_i = union of cols(A[c,:]), c in J_i
 * (sorted). Two-call protocol: pass NULL arrays to get the counts. HOST out. */
int spai_row_index_sets(spai_ctx* ctx, int64_t row, int64_t* num_j, int64_t* j_host,
                        int64_t* num_i, int64_t* i_host);

/* ls mode: values of M re-solved on the pattern of ONE trajectory, in EDGE
 * order (0 for removed edges): m_val_host f64[E]. */
int spai_ls_solve_values_host(spai_ctx* ctx, const int64_t* actions_host, int64_t T,
                              int dtype, double* m_val_host, void* stream);

/* PreconditionerEnv.calculate_residual (preconditioner.py:79-93) for an
 * arbitrary pair of sparse matrices given as HOST COO: ||M @ A - I||_F on the
 * device; also returns the stored-entry count of M after coalescing. */
int spai_residual_pair_host(int device, int64_t n, int64_t m_nnz, const int64_t* m_row_host,
                            const int64_t* m_col_host, const double* m_val_host, int64_t a_nnz,
                            const int64_t* a_row_host, const int64_t* a_col_host,
                            const double* a_val_host, int dtype, double* residual_out,
                            int64_t* m_coalesced_nnz_out);

/* One masked-categorical environment step (policy.py:64-73, gflownet.py:116-119,
 * :148, :177-179, log.py:67-87) for B samples:
 *   logits_dev f32[A] (logits_ld == 0: one vector shared by all samples) or
 *   f32[B, A] with leading dimension logits_ld; taken_dev u32[B, words_ld] is
 *   read and, for the sampled id, updated in place; uniforms_dev f32[B] in
 *   [0,1) drive the inverse-CDF draw; done_dev u8[B] in/out (set when the
 *   terminal id A-1 is drawn); action_dev i64[B] (-1 for finished samples);
 *   prob_dev f32[B] (1.0 for finished samples) = masked softmax probability of
 *   the drawn id. A = number of actions (length of a logits row; the terminal
 *   id is A-1, gflownet.py:177). */
int spai_sample_step_dev(spai_ctx* ctx, const float* logits_dev, int64_t logits_ld, int64_t A,
                         uint32_t* taken_dev, int64_t words_ld, const float* uniforms_dev,
                         uint8_t* done_dev, int64_t B, int64_t* action_dev, float* prob_dev,
                         void* stream);

/* Secondary partitioning across GPUs: evaluate only rows
 * [row_begin, row_end) of M for every trajectory. res2_partial_dev f64[B] receives the
 * sum of squared row residuals of that range (no sqrt, no mix; rows of the range
 * whose union misses the diagonal included), nnz_m_dev i64[B] the full nnz(M).
 * Sum res2 over the ranks (all-reduce), then spai_finalize_rewards_dev applies
 * sqrt and the mix formula (preconditioner.py:154-163, :64). */
int spai_reward_rows_dev(spai_ctx* ctx, const int64_t* actions_dev, int64_t B, int64_t T, int64_t ld,
                         int mode, int dtype, int64_t row_begin, int64_t row_end,
                         double* res2_partial_dev, int64_t* nnz_m_dev, void* stream);
int spai_finalize_rewards_dev(spai_ctx* ctx, const double* res2_dev, const int64_t* nnz_m_dev, int64_t B,
                              double alpha, int dtype, double* reward_dev, double* residual_dev,
                              void* stream);

/* Whole-trajectory sampling support (Gumbel-top-k; equal in distribution to the
 * step-by-step masked categorical draws of gflownet.py:135-179): given perturbed
 * keys f32[B, keys_ld] (logit + Gumbel noise, column A-1 = terminal), writes the
 * taken-bitmask u32[B, words_ld] of {i : key[b,i] > key[b,A-1]} plus the terminal
 * bit, and length i32[B] = number of ids drawn (terminal included). */
int spai_pack_taken_dev(spai_ctx* ctx, const float* keys_dev, int64_t keys_ld, int64_t A, int64_t B,
                        uint32_t* taken_dev, int64_t words_ld, int32_t* length_dev, void* stream);

/* ---- Whole trajectories of the environment step on the device. ----
 * gflownet/gflownet.py:135-179 draws one id per step from the masked, re-normalised policy
 * (policy.py:64-73) until the terminal id A-1 comes up. With one logits vector per epoch that
 * process is an exponential race (id i arrives at E_i / exp(logit_i), E_i ~ Exp(1)): ids come out
 * in order of arrival and the trajectory ends at the terminal's arrival — the same distribution,
 * O(A) per sample instead of O(A * T). E_i = Philox4x32-10(key = seed, counter = (i / 4, sample0 + b)),
 * so a key depends on (seed, global sample index, id) only: no B x A tensor exists at any point and a
 * shard [sample0, sample0 + B) of a batch equals the same rows of the full batch.
 *
 * spai_sample_taken_dev: taken_dev u32[B, words_ld] = bitmask of the ids drawn before the terminal
 *   plus the terminal's bit (the input of spai_reward_from_taken_dev), length_dev i32[B] = ids drawn,
 *   terminal included. keys_out_dev (tests; may be NULL) f32[B, keys_ld] receives every key
 *   lq = log(t_i / t_terminal) (0 for the terminal): id i is drawn iff lq < 0, in ascending (lq, id) order.
 * spai_sample_order_dev: actions_dev (id_bytes 4: i32, 8: i64) [B, ld] = the drawn ids in draw order,
 *   the terminal id last, then -1 padding (the layout of Log._actions transposed, log.py:89); length_dev
 *   from spai_sample_taken_dev with the same logits / seed / sample0; ld >= max length. Synchronises
 *   `stream` (a trajectory longer than ld is SPAI_ERR_INVALID). */
int spai_sample_taken_dev(int device, const float* logits_dev, int64_t A, int64_t B, uint64_t seed,
                          int64_t sample0, uint32_t* taken_dev, int64_t words_ld, int32_t* length_dev,
                          float* keys_out_dev, int64_t keys_ld, void* stream);
int spai_sample_order_dev(int device, const float* logits_dev, int64_t A, int64_t B, uint64_t seed,
                          int64_t sample0, const int32_t* length_dev, void* actions_dev, int id_bytes,
                          int64_t ld, void* stream);

/* Many masked-categorical environment steps per launch (k4p_steps.cuh): the step of
 * spai_sample_step_dev (policy.py:64-73, gflownet.py:148,:177-179, log.py:67-87) repeated up to
 * `nsteps` times per sample with per-block running sums of exp(logit - max) kept on chip, so a step
 * costs O(sqrt(A)) instead of three passes over the A logits. One logits vector f32[A] for all samples
 * (the policy is constant inside sample_states). taken_dev u32[B, words_ld] and done_dev u8[B] are
 * in/out exactly as in spai_sample_step_dev. uniforms_dev f32[nsteps, B] in [0,1) drives step s of
 * sample b with uniforms[s * B + b]; NULL = Philox4x32-10(seed; counter = ((step0 + s) / 4, sample0 + b)).
 * Outputs, row b, column step0 + s: actions_dev (id_bytes 4: i32, 8: i64) [B, ld] = drawn id, -1 once
 * the sample has finished; probs_dev f32[B, ld] (may be NULL) = its masked-softmax probability, 1.0 once
 * finished; steps_taken_dev i32[B] (may be NULL) = ids drawn in this call. A sample still not done
 * after the call has simply used all nsteps (call again with step0 += nsteps). */
int spai_sample_steps_dev(int device, const float* logits_dev, int64_t A, int64_t B, uint32_t* taken_dev,
                          int64_t words_ld, uint8_t* done_dev, const float* uniforms_dev, uint64_t seed,
                          int64_t sample0, int64_t step0, int64_t nsteps, void* actions_dev, int id_bytes,
                          float* probs_dev, int64_t ld, int32_t* steps_taken_dev, void* stream);

/* ---- Device-side ingest: the step immediately before the reward path. ----
 * All array arguments are DEVICE pointers on `device`; outputs are caller-allocated; totals that
 * size a later allocation come back through a HOST pointer (the call synchronises `stream`).
 * Rows are limited to 512 entries (working sets live in shared memory): SPAI_ERR_UNSUPPORTED beyond.
 *
 * gflownet/utils.py:54-63 (`market_matrix_to_sparse_tensor`: scipy COO -> torch sparse) followed by
 * the coalesce the reference's sparse ops apply: COO (any order, repeated coordinates summed in
 * entry order) -> CSR with ascending columns. ptr_dev i32[n+1], col/val sized for nnz entries. */
int spai_ingest_coo_to_csr_dev(int device, int64_t n, int64_t nnz, const int64_t* row_dev,
                               const int64_t* col_dev, const double* val_dev, int32_t* ptr_dev,
                               int32_t* col_out_dev, double* val_out_dev, int64_t* nnz_out_host,
                               void* stream);
/* GFlowNet100.py:139-141 (`LU = L @ U`, the initial matrix from sparse factors): CSR SpGEMM C = A*B in
 * two calls: _count fills c_ptr_dev i32[n+1] and returns nnz(C); _fill writes columns (ascending) and
 * values (products summed in the order of A's row entries: deterministic). a_val / b_val may be NULL
 * (pattern product, all ones). */
int spai_ingest_spgemm_count_dev(int device, int64_t n, const int32_t* a_ptr, const int32_t* a_col,
                                 const int32_t* b_ptr, const int32_t* b_col, int32_t* c_ptr_dev,
                                 int64_t* c_nnz_host, void* stream);
int spai_ingest_spgemm_fill_dev(int device, int64_t n, const int32_t* a_ptr, const int32_t* a_col,
                                const double* a_val, const int32_t* b_ptr, const int32_t* b_col,
                                const double* b_val, const int32_t* c_ptr_dev, int32_t* c_col_dev,
                                double* c_val_dev, void* stream);
/* scipy's product (the reference's `L @ U`) stores only results != 0: drop exact zeros from a CSR.
 * out_col_dev / out_val_dev sized for the input nnz (in-place is NOT allowed). */
int spai_ingest_csr_drop_zeros_dev(int device, int64_t n, const int32_t* ptr_dev, const int32_t* col_dev,
                                   const double* val_dev, int32_t* out_ptr_dev, int32_t* out_col_dev,
                                   double* out_val_dev, int64_t* nnz_out_host, void* stream);
/* Candidate superset S of SURVEY.md 8d (what the drivers' "initial matrix" pattern stands for):
 *   order 0: S(i) = first k entries of pattern(I) U pattern(A) U ... U pattern(A^max_power) (i,:) ordered
 *            by (graph distance from i, column id);
 *   order 1: S(i) = first k entries of pattern(A)(i,:) ordered by (|col - i|, col).
 * Output row-major COO with ascending columns per row: s_ptr_dev i64[n+1], s_row_dev / s_col_dev i64
 * sized for n*k entries (may be NULL to get only the counts), total through s_nnz_host. */
int spai_ingest_superset_dev(int device, int64_t n, const int32_t* a_ptr, const int32_t* a_col, int k,
                             int max_power, int order, int64_t* s_ptr_dev, int64_t* s_row_dev,
                             int64_t* s_col_dev, int64_t* s_nnz_host, void* stream);
/* Initial values on S: omega * sum_{j < terms} (I - omega*A)^j restricted to S, omega = 1 / max_i sum_j |A_ij|
 * (a truncated Neumann series of A^-1; stands in for the drivers' spilu factors, GFlowNet100.py:126-153).
 * fp64; val_out_dev f64[nnz(S)] in the order of s_col_dev; omega_host (optional) receives omega. */
int spai_ingest_neumann_dev(int device, int64_t n, const int32_t* a_ptr, const int32_t* a_col,
                            const double* a_val, const int64_t* s_ptr_dev, const int64_t* s_col_dev,
                            int terms, double* omega_host, double* val_out_dev, void* stream);

/* Per-kernel device time of the LAST reward call on this context, in ms,
 * measured with CUDA events on the caller's stream (masks, transpose, reward
 * kernel, finalize) and the number of kernels launched by it. */
typedef struct spai_timing {
  float ms_masks, ms_transpose, ms_reward, ms_finalize, ms_total;
  int32_t launches;
  int32_t chunks;
  double algorithmic_bytes; /* SURVEY.md §8d G summed over the batch            */
  double compulsory_bytes;  /* bytes that must cross HBM for the batch           */
  double h2d_bytes;         /* host->device bytes actually copied (host entry)   */
} spai_timing;
/* Rows served by the tensor-core copy kernel (K3m, k3m_mma.cuh) once its records exist (built by the first
 * copy/fp32 call with >= 64 trajectories on a pattern without repeated coordinates and <= 32 candidates per row;
 * preconditioner.py:79-93 is what it computes): rows with <= 16 candidates, rows with 17..32. Both 0 = the CUDA-core
 * row sweep serves those calls. */
int spai_ctx_k3m_rows(const spai_ctx* ctx, int64_t* rows16, int64_t* rows32);
int spai_ctx_enable_timing(spai_ctx* ctx, int enable);
int spai_ctx_last_timing(const spai_ctx* ctx, spai_timing* out);

#ifdef __cplusplus
}
#endif
#endif /* SPAI_B200_H_ */
