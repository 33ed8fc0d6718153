// K3m — copy-mode row residuals on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// for rows with 9..32 candidates where the (row, kept-mask) table of K3t does not exist and the row sweep
// (K3) spends 170-550 warp instructions per (row, trajectory).
//
// Same quantity as K3 (preconditioner.py:79-93 on the pattern of gflownet/utils.py:315-353): row i of
// M*A - I with kept indicator kappa in {0,1}^k is sum_j kappa_j w_j - e_i (w_j = m_j * A[c_j, :]), so
//     r_i(kappa) = kappa^T G kappa - 2 kappa.g + [i in union],   G = W W^T (k x k),  g_j = w_j(i).
// With the Cholesky factor G = L L^T and L v = g (both once per context, fp64, semi-definite safe)
//     r_i(kappa) = || L^T kappa - v ||^2 + c,   c = [i in union] - |v|^2 >= 0 (the row's least-squares optimum)
// — a sum of squares, no cancellation. For 128 trajectories at once, Z = kappa * L is ONE small dense
// contraction with an exact operand: A = 2*kappa as bf16 (0x4000 per kept bit, written straight into TENSOR
// MEMORY by the thread that owns the trajectory's lane: tcgen05.st, no shared-memory traffic for A),
// B = L/2 split into NSPLIT bf16 terms (every product exact, fp32 accumulation in TMEM), M = 128 trajectories,
// N = K = 16 or 32 (tcgen05.mma with the A operand in TMEM). The epilogue reads its trajectory's N
// accumulators (tcgen05.ld 32x32b), subtracts v and adds N squares.
//
// Warp-specialised CTA (288 threads), rows flow through three mbarrier rings without a block barrier:
//   (row_lo / row_hi index the CLASS's row list: rows with <= 16 candidates and rows with 17..32 are separate launches)
//   warps 4-7  producers: kept-mask words (cp.async into a private shared-memory window, two chunks of 8
//              words in flight) -> the row's k bits -> bf16 pairs (one 64-bit multiply per 4 bits) -> tcgen05.st
//   warp 8     one thread: cp.async.bulk of the record stages, tcgen05.mma issue, tcgen05.commit
//   warps 0-3  epilogue: tcgen05.ld of their TMEM lane quarter, (z - v)^2 sums, fixed-order partial sums
// fp32 only (fp64 stays on K3); rows with repeated coordinates keep K3.
#pragma once

#include <cuda_bf16.h>

#include "k3_copy.cuh"
#include "spai_internal.cuh"

namespace spai {

constexpr int K3M_THREADS = 288;                               // 4 epilogue + 4 producer warps + 1 issue warp
constexpr int K3M_NSTAGE = 4;                                  // record stages in flight
constexpr int K3M_MW = 8;                                      // mask words per producer chunk (+1 overlap word)

template <int N, int NSPLIT>
struct K3mGeom {
  static constexpr int KP = N;                                 // K = N: the candidates of one row
  static constexpr int TILE_B = N * KP * 2;                    // one bf16 term of L, canonical K-major core-matrix layout
  static constexpr int TAIL = 16 + 4 * N;                      // {c, pad[3], v[N]}
  static constexpr int RB = NSPLIT * TILE_B + TAIL;
  static constexpr int RPS = N == 16 ? 4 : 1;                  // rows per bulk-copied stage (~6.4 KB)
  static constexpr int STAGE = (RPS * RB + 127) / 128 * 128;
  static constexpr int SBO = (KP / 8) * 128;                   // stride between 8-row groups of B
  static constexpr int ACOLS = KP / 2;                         // TMEM columns of one A tile (two bf16 per column)
  static_assert(N == 16 || N == 32, "N");
};
struct K3mTail { float c; int32_t sp; int32_t k; int32_t pad; };   // followed by float v[N]

// byte offset of element (row, kk) in the no-swizzle K-major canonical layout: 8 x 16-byte core matrices,
// the two (or more) K chunks of a row group 128 bytes apart, row groups SBO apart
__host__ __device__ constexpr int k3m_off(int row, int kk, int KP) {
  return (row >> 3) * ((KP / 8) * 128) + (kk >> 3) * 128 + (row & 7) * 16 + (kk & 7) * 2;
}
template <int N, int NSPLIT, int NTM>
__host__ __device__ constexpr int k3m_smem_bytes() {
  using G = K3mGeom<N, NSPLIT>;
  return 512 + K3M_NSTAGE * G::STAGE + 2 * (K3M_MW + 1) * 128 * NTM * 4;
}
// ring depth (rows in flight) of the A tiles and of the accumulators: as deep as 512 TMEM columns allow, at most 8
// ring depth (rows in flight) of the A tiles and of the accumulators
template <int N, int NTM>
__host__ __device__ constexpr int k3m_depth() {
  return 2;                                                   // measured: depths 4 and 8 are no faster; 2 keeps the TMEM
                                                              // allocation small enough for 2-4 CTAs per SM
}
template <int N, int NTM>
__host__ __device__ constexpr int k3m_tmem_cols() {           // accumulator ring + A ring, rounded to an allocation size
  const int raw = k3m_depth<N, NTM>() * NTM * (N + N / 2);
  static_assert(k3m_depth<N, NTM>() * NTM * (N + N / 2) <= 512, "TMEM");
  return raw <= 32 ? 32 : raw <= 64 ? 64 : raw <= 128 ? 128 : raw <= 256 ? 256 : 512;
}

// ------------------------------------------------------------------------------------------------
// once per context: G, g -> Cholesky (fp64, shared memory, one warp per row) -> record
// ------------------------------------------------------------------------------------------------
template <int N, int NSPLIT>
__global__ void __launch_bounds__(128)
k3m_build_kernel(const Rec32* __restrict__ recs, const int64_t* __restrict__ cptr, const int32_t* __restrict__ c_col,
                 const RowHdr* __restrict__ rhdr, const int32_t* __restrict__ rows, int64_t n, unsigned char* __restrict__ out,
                 int2* __restrict__ hdr_out) {
  using Geo = K3mGeom<N, NSPLIT>;
  constexpr int KP = Geo::KP;
  constexpr int LD = 33, AUG = 32;
  __shared__ double Gs[4][LD * LD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t idx_row = (int64_t)blockIdx.x * 4 + warp;      // position in the class's row list
  if (idx_row >= n) return;
  const int64_t row = rows[idx_row];
  double* G = Gs[warp];
  const RowHdr h = rhdr[row];
  const int k = h.k < N ? h.k : N;                             // the host only builds when max_k <= N
  for (int i = lane; i < LD * LD; i += 32) G[i] = 0.0;
  __syncwarp();
  const Rec32* rr = recs + cptr[row];
  const int32_t* cc = c_col + cptr[row];
  bool has_diag = false;
  for (int r = lane; r < h.cnt; r += 32) {
    const int x = cc[r];
    const int e = (int)rec_e(rr[r].flags);
    const double w = (double)rr[r].w;
    if (x == (int)row) { G[AUG * LD + e] = w; has_diag = true; }
    for (int r2 = r; r2 < h.cnt && cc[r2] == x; ++r2)          // records of a segment ascend in e: lower triangle
      atomicAdd(&G[(int)rec_e(rr[r2].flags) * LD + e], w * (double)rr[r2].w);
  }
  has_diag = __any_sync(0xffffffffu, has_diag);
  __syncwarp();
  const double d0 = lane < k ? G[lane * LD + lane] : 0.0;
  for (int j = 0; j < k; ++j) {
    const double d = G[j * LD + j];
    const double dj0 = __shfl_sync(0xffffffffu, d0, j);
    const bool ok = d > 1e-12 * dj0 && d > 0.0;
    const double inv = ok ? rsqrt(d) : 0.0;
    __syncwarp();
    if (lane == j) G[j * LD + j] = ok ? d * inv : 0.0;
    const int myrow = (lane == 0) ? AUG : ((lane > j && lane < k) ? lane : -1);   // lane 0 also carries the g row
    if (myrow >= 0) G[myrow * LD + j] *= inv;
    __syncwarp();
    if (myrow >= 0) {
      const double lij = G[myrow * LD + j];
      const int top = myrow == AUG ? k - 1 : myrow;
      for (int j2 = j + 1; j2 <= top; ++j2) G[myrow * LD + j2] -= lij * G[j2 * LD + j];
    }
    __syncwarp();
  }
  double vv = (lane < k) ? G[AUG * LD + lane] : 0.0;
  vv *= vv;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vv += __shfl_xor_sync(0xffffffffu, vv, o);
  unsigned char* o = out + idx_row * (int64_t)Geo::RB;
  for (int idx = lane; idx < N * KP; idx += 32) {
    const int nn = idx / KP, kk = idx % KP;                    // B[nn][kk] = L[kk][nn]
    double x = 0.0;
    if (kk < k && nn <= kk) x = 0.5 * G[kk * LD + nn];          // the A operand carries 2 * kappa
    const int off = k3m_off(nn, kk, KP);
#pragma unroll
    for (int s = 0; s < NSPLIT; ++s) {
      const __nv_bfloat16 t = __float2bfloat16_rn((float)x);
      x -= (double)__bfloat162float(t);
      *reinterpret_cast<__nv_bfloat16*>(o + s * Geo::TILE_B + off) = t;
    }
  }
  if (lane == 0) {
    K3mTail t;
    const double c = (has_diag ? 1.0 : 0.0) - vv;
    t.c = (float)(c > 0.0 ? c : 0.0);
    t.sp = h.sp; t.k = k; t.pad = 0;
    *reinterpret_cast<K3mTail*>(o + NSPLIT * Geo::TILE_B) = t;
    hdr_out[idx_row] = make_int2(h.sp, k);
  }
  float* vout = reinterpret_cast<float*>(o + NSPLIT * Geo::TILE_B + 16);
  if (lane < N) vout[lane] = lane < k ? (float)G[AUG * LD + lane] : 0.f;
}

// ------------------------------------------------------------------------------------------------
// tcgen05 plumbing (PTX as CUTLASS emits it: cute/arch/mma_sm100_umma.hpp, copy_sm100.hpp, tmem_allocator_sm100.hpp)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t k3m_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  // SWIZZLE_NONE K-major: start address, leading (K-chunk) and stride (8-row group) byte offsets in 16-byte
  // units, descriptor version 1 (Blackwell)
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128
__host__ __device__ constexpr uint32_t k3m_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem descriptor given as its two 32-bit halves]; ACC = 0 overwrites D
template <int ACC>
__device__ __forceinline__ void k3m_mma(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, uint32_t bdesc_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, {%6, %6, %6, %6}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "n"(ACC), "r"(0u)
      : "memory");
}
// 8 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void k3m_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// 16 kept bits -> 16 bf16 (0x4000 = 2.0 per set bit), two elements per word: one 64-bit multiply spreads 4 bits
__device__ __forceinline__ void k3m_expand16(uint32_t bits, uint32_t (&w)[8]) {
  constexpr unsigned long long M = (1ull << 14) | (1ull << 29) | (1ull << 44) | (1ull << 59);
  constexpr unsigned long long MASK = 0x4000400040004000ull;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const unsigned long long p = ((unsigned long long)((bits >> (4 * q)) & 0xfu) * M) & MASK;
    w[2 * q] = (uint32_t)p;
    w[2 * q + 1] = (uint32_t)(p >> 32);
  }
}
__device__ __forceinline__ void k3m_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void k3m_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void k3m_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 16 consecutive fp32 columns of this thread's TMEM lane; the caller issues tcgen05.wait::ld before use
__device__ __forceinline__ void k3m_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void k3m_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void k3m_cp4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

template <int N, int NSPLIT, int NTM>
__global__ void __launch_bounds__(K3M_THREADS, 2)
k3m_kernel(const unsigned char* __restrict__ recs, const int2* __restrict__ rhdr, const uint32_t* __restrict__ maskT,
           int64_t Bp, int64_t W, double* __restrict__ partial, int row_lo, int row_hi, int dbg) {
  using Geo = K3mGeom<N, NSPLIT>;
  constexpr int KP = Geo::KP, RPS = Geo::RPS, NA = k3m_depth<N, NTM>(), NACC = NA;
  constexpr int NW = N;                                                // accumulator columns of one (row, tile)
  constexpr uint32_t TCOLS = (uint32_t)k3m_tmem_cols<N, NTM>();
  constexpr uint32_t ACOL0 = NACC * NTM * NW;                            // the A ring follows the accumulator ring
  extern __shared__ __align__(128) unsigned char k3m_smem[];
  uint64_t* bfull = reinterpret_cast<uint64_t*>(k3m_smem);             // [NSTAGE] record stage landed (tx)
  uint64_t* bempty = bfull + K3M_NSTAGE;                               // [NSTAGE] 128 epilogue threads are done with it
  uint64_t* afull = bempty + K3M_NSTAGE;                               // [NA] 128 producer threads stored the row's A tiles (TMEM)
  uint64_t* aempty = afull + 8;                                        // [NA] the MMAs that read them have completed
  uint64_t* accfull = aempty + 8;                                      // [NACC] the row's accumulators are complete
  uint64_t* accempty = accfull + 8;                                    // [NACC] 128 epilogue threads have read them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 8);
  unsigned char* ring = k3m_smem + 512;
  uint32_t* mwin = reinterpret_cast<uint32_t*>(ring + K3M_NSTAGE * Geo::STAGE);   // [2][MW + 1][128 * NTM]

  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t nrows_all = row_hi - row_lo;
  const int r0 = row_lo + (int)(nrows_all * blockIdx.x / gridDim.x);
  const int r1 = row_lo + (int)(nrows_all * (blockIdx.x + 1) / gridDim.x);
  const int nrows = r1 - r0;
  const int nstages = (nrows + RPS - 1) / RPS;
  const int64_t bbase = (int64_t)blockIdx.y * (128 * NTM);

  if (tid == 0) {
    for (int i = 0; i < K3M_NSTAGE; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 4); }
    for (int i = 0; i < NA; ++i) { mbar_init(&afull[i], 4); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&accfull[i], 1); mbar_init(&accempty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  k3m_fence_before();
  __syncthreads();
  k3m_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // =========================================================== record loads + MMA issue (one thread)
    if ((tid & 31) == 0) {
      auto load_stage = [&](int m) {
        const int slot = m % K3M_NSTAGE;
        const int cnt = (nrows - m * RPS < RPS) ? nrows - m * RPS : RPS;
        mbar_expect_tx(&bfull[slot], (uint32_t)(cnt * Geo::RB));
        bulk_g2s(ring + (size_t)slot * Geo::STAGE, recs + (int64_t)(r0 + m * RPS) * Geo::RB, (uint32_t)(cnt * Geo::RB), &bfull[slot]);
      };
      for (int m = 0; m < K3M_NSTAGE && m < nstages; ++m) load_stage(m);
      constexpr uint32_t IDESC = k3m_idesc(N);
      for (int it = 0; it < nrows; ++it) {
        const int m = it / RPS;
        if (it % RPS == 0) {
          const int m2 = m + 2;                                        // its slot held stage m - 2
          if (m >= 2 && m2 < nstages) {
            mbar_wait(&bempty[m2 % K3M_NSTAGE], (uint32_t)(((m2 / K3M_NSTAGE) - 1) & 1));
            load_stage(m2);
          }
          mbar_wait(&bfull[m % K3M_NSTAGE], (uint32_t)((m / K3M_NSTAGE) & 1));
        }
        mbar_wait(&afull[it % NA], (uint32_t)((it / NA) & 1));
        mbar_wait(&accempty[it % NACC], (uint32_t)(((it / NACC) & 1) ^ 1));
        k3m_fence_after();
        // descriptor of the row's first B tile; the other tiles differ in the 14-bit start-address field only
        const uint32_t blo = ((smem_u32(ring) + (uint32_t)((m % K3M_NSTAGE) * Geo::STAGE + (it % RPS) * Geo::RB)) >> 4) | ((128u >> 4) << 16);
        constexpr uint32_t BHI = ((uint32_t)Geo::SBO >> 4) | (1u << 14);      // stride byte offset, descriptor version 1
        const uint32_t abase = tmem + ACOL0 + (uint32_t)((it % NA) * NTM * Geo::ACOLS);
        const uint32_t dbase = tmem + (uint32_t)((it % NACC) * NTM * NW);
        // the NSPLIT terms accumulate into the same N columns; consecutive MMAs go to DIFFERENT tiles, so an MMA never
        // waits for the one issued just before it (back-to-back accumulation into one tile cost ~45 cycles per MMA)
        if (!(dbg & 1))
#pragma unroll
        for (int s = 0; s < NSPLIT; ++s)
#pragma unroll
          for (int ks = 0; ks < KP / 16; ++ks)
#pragma unroll
            for (int j = 0; j < NTM; ++j) {
              const uint32_t bl = blo + (uint32_t)((s * Geo::TILE_B + ks * 256) >> 4);
              if (s == 0 && ks == 0) k3m_mma<0>(dbase + j * NW, abase + j * Geo::ACOLS + ks * 8, bl, BHI, IDESC);
              else k3m_mma<1>(dbase + j * NW, abase + j * Geo::ACOLS + ks * 8, bl, BHI, IDESC);
            }
        k3m_commit(&accfull[it % NACC]);
      }
    }
  } else if (warp >= 4) {
    // =========================================================== producers: kept bits -> bf16 A tiles
    const int t = tid - 128;
    const uint32_t* mp = maskT + bbase + t;
    constexpr int MWS = (K3M_MW + 1) * 128 * NTM;                      // words of one window buffer
    auto fetch = [&](int buf, int64_t wb) {                            // words [wb, wb + MW] of this thread's trajectories
#pragma unroll
      for (int w = 0; w <= K3M_MW; ++w)
#pragma unroll
        for (int j = 0; j < NTM; ++j) {
          uint32_t* dst = mwin + buf * MWS + w * (128 * NTM) + j * 128 + t;
          if (wb + w < W) k3m_cp4(dst, mp + (wb + w) * Bp + j * 128);
          else *dst = 0u;
        }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int64_t wb = nrows > 0 ? (int64_t)(__ldg(&rhdr[r0].x) >> 5) : 0;
    int cur = 0;
    fetch(0, wb);
    fetch(1, wb + K3M_MW);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    // row headers: lane l of every producer warp holds the header of row (block * 32 + l); the next block of 32 is in
    // flight while this one is consumed (one coalesced 512-byte load per 32 rows instead of a dependent load per row)
    const int lane = tid & 31;
    auto hload = [&](int blk) {
      const int i = blk * 32 + lane;
      return i < nrows ? __ldg(rhdr + r0 + i) : make_int2(0, 0);
    };
    int2 hcur = hload(0), hnext = hload(1);
    for (int it = 0; it < nrows; ++it) {
      if (it && (it & 31) == 0) { hcur = hnext; hnext = hload((it >> 5) + 1); }
      const int sp = __shfl_sync(0xffffffffu, hcur.x, it & 31), k = __shfl_sync(0xffffffffu, hcur.y, it & 31);
      const int64_t wd = sp >> 5;
      if (wd - wb >= 2 * K3M_MW) {                                      // far jump (sparse row class): re-base both buffers
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        wb = wd;
        cur = 0;
        fetch(0, wb);
        fetch(1, wb + K3M_MW);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else if (wd - wb >= K3M_MW) {                                   // the row's two words must lie in [wb, wb + MW]
        wb += K3M_MW;
        cur ^= 1;
        fetch(cur ^ 1, wb + K3M_MW);                                    // the buffer just left is free (only this thread reads it)
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      }
      const int sh = sp & 31;
      const uint32_t kmask = k >= 32 ? 0xffffffffu : ((1u << k) - 1u);
      const uint32_t* wp = mwin + cur * MWS + (int)(wd - wb) * (128 * NTM) + t;
      uint32_t bits[NTM];
#pragma unroll
      for (int j = 0; j < NTM; ++j) bits[j] = __funnelshift_r(wp[j * 128], wp[128 * NTM + j * 128], sh) & kmask;
      // the A slot of row it - NA is free once that row's MMAs have completed: the same tcgen05.commit that publishes its
      // accumulators (one commit per row; NA == NACC)
      if (it >= NA) mbar_wait(&accfull[it % NA], (uint32_t)(((it / NA) - 1) & 1));
      k3m_fence_after();
      const uint32_t a0 = tmem + ((uint32_t)((warp - 4) * 32) << 16) + ACOL0 + (uint32_t)((it % NA) * NTM * Geo::ACOLS);
      if (!(dbg & 2))
#pragma unroll
      for (int j = 0; j < NTM; ++j)
#pragma unroll
        for (int h = 0; h < KP / 16; ++h) {
          uint32_t w[8];
          k3m_expand16(bits[j] >> (16 * h), w);
          k3m_st8(a0 + (uint32_t)(j * Geo::ACOLS + h * 8), w);
        }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      k3m_fence_before();
      __syncwarp();
      if (lane == 0) k3m_arrive(&afull[it % NA]);                    // one arrival per producer warp
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    // =========================================================== epilogue: r = c + sum_j (Z_j - v_j)^2
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's TMEM lane quarter
    float rs[NTM];
    double tot[NTM];
#pragma unroll
    for (int j = 0; j < NTM; ++j) { rs[j] = 0.f; tot[j] = 0.0; }
    for (int it = 0; it < nrows; ++it) {
      const int m = it / RPS, slot = m % K3M_NSTAGE;
      if (it % RPS == 0) mbar_wait(&bfull[slot], (uint32_t)((m / K3M_NSTAGE) & 1));
      mbar_wait(&accfull[it % NACC], (uint32_t)((it / NACC) & 1));
      k3m_fence_after();
      float z[NTM][N];
      if (!(dbg & 4)) {
#pragma unroll
        for (int j = 0; j < NTM; ++j)
#pragma unroll
          for (int q0 = 0; q0 < N; q0 += 16) k3m_ld16(lane_addr + (uint32_t)((it % NACC) * NTM * NW + j * NW + q0), &z[j][q0]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < NTM; ++j)
#pragma unroll
          for (int q = 0; q < N; ++q) z[j][q] = (float)(it + q);
      }
      k3m_fence_before();
      __syncwarp();
      if ((tid & 31) == 0) k3m_arrive(&accempty[it % NACC]);           // the accumulators are in registers (one arrival per warp)
      const unsigned char* tl = ring + (size_t)slot * Geo::STAGE + (size_t)(it % RPS) * Geo::RB + NSPLIT * Geo::TILE_B;
      const float c = reinterpret_cast<const K3mTail*>(tl)->c;
      const float4* vv = reinterpret_cast<const float4*>(tl + 16);
      float r[NTM];
#pragma unroll
      for (int j = 0; j < NTM; ++j) r[j] = c;
#pragma unroll
      for (int q = 0; q < N; q += 4) {
        const float4 v4 = vv[q >> 2];                                  // warp-uniform broadcast
#pragma unroll
        for (int j = 0; j < NTM; ++j) {
          const float d0 = z[j][q] - v4.x, d1 = z[j][q + 1] - v4.y, d2 = z[j][q + 2] - v4.z, d3 = z[j][q + 3] - v4.w;
          r[j] = fmaf(d0, d0, r[j]);
          r[j] = fmaf(d1, d1, r[j]);
          r[j] = fmaf(d2, d2, r[j]);
          r[j] = fmaf(d3, d3, r[j]);
        }
      }
      if (it % RPS == RPS - 1 || it == nrows - 1) {
        __syncwarp();
        if ((tid & 31) == 0) k3m_arrive(&bempty[slot]);
      }
#pragma unroll
      for (int j = 0; j < NTM; ++j) rs[j] += r[j];
      if ((it & 15) == 15) {
#pragma unroll
        for (int j = 0; j < NTM; ++j) { tot[j] += (double)rs[j]; rs[j] = 0.f; }
      }
    }
#pragma unroll
    for (int j = 0; j < NTM; ++j) {
      const int64_t bj = bbase + (int64_t)j * 128 + tid;
      if (bj < Bp) partial[(int64_t)blockIdx.x * Bp + bj] = tot[j] + (double)rs[j];
    }
  }
  k3m_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS));
}

}  // namespace spai
