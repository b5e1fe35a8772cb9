// gemm.cuh -- the generic tile GEMM kernel (C tile 128 x 64 per CTA) on the DMMA mainloop of
// common.cuh, with structured k-ranges (triangular operands) and pluggable epilogues.
#pragma once
#include "common.cuh"

namespace agp {

// k-range of a row tile (tile_m = blockIdx.x), in elements
constexpr int KR_FULL = 0;   // [0, K)
constexpr int KR_LOWER = 1;  // [0, (tile_m+1)*BM)        A lower triangular (block rows)
constexpr int KR_UPPER = 2;  // [tile_m*BM, K)            A upper triangular
constexpr int KR_DIAG = 3;   // [tile_m*BM, (tile_m+1)*BM)
// which output tiles are computed (others are left untouched)
constexpr int TS_ALL = 0;
constexpr int TS_NBLK_LT = 1;  // column block (of BM) <  tile_m
constexpr int TS_NBLK_GT = 2;  // column block (of BM) >  tile_m
constexpr int TS_NBLK_LE = 3;  // column block (of BM) <= tile_m   (lower triangle incl. diagonal block)
constexpr int TS_NBLK_LE1 = 4; // column block (of BM) <= tile_m + 1   (the same when the row tiles start one block below the column tiles)

struct GemmArgs {
  const double* A;
  int64_t lda;
  const double* B;
  int64_t ldb;
  int K;
  int kmode;
  int tmode;
  int swizzle = 0;  // > 0: 1-D grid of tiles_m * tiles_n blocks, see gemm_kernel
  int tiles_m = 0, tiles_n = 0;
};

// tri / diag_step: when the A operand is triangular (TRI_LOWER: k <= m, TRI_UPPER: k >= m) the k-steps
// [diag_step, diag_step + BM/BK) cover its diagonal 128 x 128 block and a warp skips the steps in which its
// 32-row slice of A is identically zero (12.5 % of a triangular sweep at M = 1024).  whole_active == false
// skips every MMA of this warp (SYRK tiles above the diagonal).
constexpr int TRI_NONE = 0, TRI_LOWER = 1, TRI_UPPER = 2;

template <int LA, int LB, int S = StageCfg<LA, LB>::stages>
__device__ __forceinline__ void gemm_mainloop(Acc& acc, double* smem, const double* __restrict__ gA, int64_t lda,
                                              const double* __restrict__ gB, int64_t ldb, int nsteps,
                                              const ThreadMap& tm, int tri = TRI_NONE, int diag_step = 0,
                                              bool whole_active = true) {
  using Cfg = StageCfg<LA, LB>;
  const int tid = threadIdx.x;
  const int64_t a_step = (LA == A_KM) ? (int64_t)BK * lda : (int64_t)BK;
  const int64_t b_step = (LB == B_KN) ? (int64_t)BK * ldb : (int64_t)BK;
  ALoad<LA> la;
  BLoad<LB> lb;
  la.init(lda, tid);
  lb.init(ldb, tid);
  const double* pa = gA;  // tile bases of the next stage to issue
  const double* pb = gB;
#pragma unroll
  for (int s = 0; s < S - 1; s++) {
    if (s < nsteps) {
      double* st = smem + s * Cfg::elems;
      la.load(st, pa);
      lb.load(st + Cfg::a_elems, pb);
      pa += a_step;
      pb += b_step;
    }
    cp_async_commit();
  }
  int slot_issue = S - 1, slot_cons = 0;
  for (int step = 0; step < nsteps; step++) {
    cp_async_wait<S - 2>();
    __syncthreads();
    if (step + S - 1 < nsteps) {
      double* st = smem + slot_issue * Cfg::elems;
      la.load(st, pa);
      lb.load(st + Cfg::a_elems, pb);
      pa += a_step;
      pb += b_step;
    }
    cp_async_commit();
    slot_issue = (slot_issue + 1 == S) ? 0 : slot_issue + 1;
    const double* st = smem + slot_cons * Cfg::elems;
    slot_cons = (slot_cons + 1 == S) ? 0 : slot_cons + 1;
    bool active = whole_active;
    if (tri != TRI_NONE) {
      const int d = step - diag_step;
      if (d >= 0 && d < BM / BK) active = (tri == TRI_LOWER) ? tm.tri_active_lower(d * BK) : tm.tri_active_upper(d * BK);
    }
    if (active) mma_stage<LA, LB>(acc, st, st + Cfg::a_elems, tm);
  }
  cp_async_wait<0>();
  __syncthreads();  // smem is free for the epilogue
}

// S: pipeline stages (shared memory = S * 25.6 KB).  The 2-stage instantiation exists for the look-ahead trailing update of
// the blocked Cholesky: one such CTA (51 KB) leaves room for the 166 KB diagonal-block kernel on the same SM.
template <int LA, int LB, class Epi, int S = StageCfg<LA, LB>::stages>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_kernel(GemmArgs g, Epi epi) {
  extern __shared__ __align__(128) double smem[];
  ThreadMap tm;
  int tile_m = blockIdx.x, tile_n = blockIdx.y;
  if (g.swizzle > 0) {
    // Triangular sweeps: linear block id -> (super-group of `swizzle` column tiles, row tile by decreasing work, column
    // tile).  Heavy row tiles are scheduled first (no long CTA in the tail of the launch) while the columns of a
    // super-group (~75 MB of the B operand at M = 1024) stay L2-resident across its row-tile passes.
    const int tiles_m = g.tiles_m, tiles_n = g.tiles_n, G = g.swizzle;
    const int id = blockIdx.x;
    const int sg = id / (G * tiles_m);
    const int gw = min(G, tiles_n - sg * G);  // width of this (possibly last, narrower) super-group
    const int r = id - sg * G * tiles_m;
    const int rank = r / gw;
    tile_n = sg * G + r % gw;
    tile_m = (g.kmode == KR_LOWER) ? tiles_m - 1 - rank : rank;
  }
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  const int nblk = n0 / BM;
  if (g.tmode == TS_NBLK_LT && !(nblk < tile_m)) return;
  if (g.tmode == TS_NBLK_GT && !(nblk > tile_m)) return;
  if (g.tmode == TS_NBLK_LE && !(nblk <= tile_m)) return;
  if (g.tmode == TS_NBLK_LE1 && !(nblk <= tile_m + 1)) return;
  int kb = 0, ke = g.K;
  if (g.kmode == KR_LOWER) ke = min(g.K, (tile_m + 1) * BM);
  if (g.kmode == KR_UPPER) kb = tile_m * BM;
  if (g.kmode == KR_DIAG) {
    kb = tile_m * BM;
    ke = kb + BM;
  }
  const double* gA = (LA == A_KM) ? g.A + (int64_t)kb * g.lda + m0 : g.A + (int64_t)m0 * g.lda + kb;
  const double* gB = (LB == B_KN) ? g.B + (int64_t)kb * g.ldb + n0 : g.B + (int64_t)n0 * g.ldb + kb;
  Acc acc;
  acc_zero(acc);
  const int nsteps = (ke - kb) / BK;
  int tri = TRI_NONE, diag_step = 0;
  if (g.kmode == KR_LOWER && ke == (tile_m + 1) * BM) {
    tri = TRI_LOWER;
    diag_step = nsteps - BM / BK;
  } else if (g.kmode == KR_UPPER) {
    tri = TRI_UPPER;
  }
  gemm_mainloop<LA, LB, S>(acc, smem, gA, g.lda, gB, g.ldb, nsteps, tm, tri, diag_step);
  epi(acc, tm, m0, n0, smem);
}

// ---- generic store epilogue for the once-per-step M x M algebra -----------------------------------
constexpr int MASK_NONE = 0;
constexpr int MASK_LOWER = 1;  // zero where col > row
constexpr int MASK_PHI = 2;    // lower triangle, diagonal halved (Cholesky pullback)

struct EpiStore {
  double* C;
  int64_t ldc;
  int rowmajor;  // 1: C[row*ldc + col], 0: C[col*ldc + row]
  double alpha;
  double beta;  // adds beta * C_old (same location)
  int mask;
  double diag_add;
  const double* r1u;  // optional rank-1 term r1 * u[row] * v[col]
  const double* r1v;
  double r1;
  __device__ __forceinline__ void operator()(Acc& acc, const ThreadMap& tm, int m0, int n0, double*) const {
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
      const int row = m0 + tm.row(mi);
      const double ur = r1u ? r1 * r1u[row] : 0.0;
#pragma unroll
      for (int ni = 0; ni < 4; ni++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int col = n0 + tm.col(ni, e);
          double* p = rowmajor ? C + (int64_t)row * ldc + col : C + (int64_t)col * ldc + row;
          double v = alpha * acc[mi][ni][e];
          if (beta != 0.0) v += beta * (*p);
          if (r1u) v += ur * r1v[col];
          if (row == col) v += diag_add;
          if (mask != MASK_NONE) {
            if (col > row) v = 0.0;
            if (mask == MASK_PHI && col == row) v *= 0.5;
          }
          *p = v;
        }
      }
    }
  }
};

}  // namespace agp
