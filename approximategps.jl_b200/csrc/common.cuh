// common.cuh -- shared device primitives for the sm_100a FP64 hot path.
//
// FP64 tensor-core math on sm_100a is the warp-level DMMA.8x8x4 instruction (every
// mma.sync.*.f64 shape lowers to it; tcgen05.mma has no f64 kind).  Measured on B200
// (profiles/r01_fp64_peak.jsonl): 37.1 TFLOP/s issue peak with 64x32 / 32x32 warp tiles fed
// by LDS.64, cuBLAS DGEMM 35.5 TFLOP/s.  Everything GEMM-shaped in this library runs through
// the one CTA-level mainloop defined here:
//
//   CTA tile   BM x BN = 128 (rows, the inducing-point index) x 64 (columns, data points)
//   threads    256 = 8 warps arranged 4 (m) x 2 (n); warp tile 32 x 32 = 4 x 4 DMMA tiles
//   k-step     BK = 16 per pipeline stage, cp.async (LDGSTS.128, L2-only) multistage ring
//   smem       padded rows (ld = 4 mod 16 doubles) -> conflict-free LDS.64 fragment reads
//   occupancy  2 CTAs / SM (<= 128 registers / thread, <= ~110 KB smem / CTA)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace agp {

constexpr int BM = 128;
constexpr int BN = 64;
constexpr int BK = 16;
constexpr int NTHREADS = 256;
constexpr int PAD = 4;

// Operand layouts in global memory.
//   A operand, element (m, k):  A_KM: A[k*lda + m]  (m contiguous, a column-major matrix)
//                               A_MK: A[m*lda + k]  (k contiguous, a row-major matrix)
//   B operand, element (k, n):  B_KN: B[k*ldb + n]  (n contiguous, a row-major matrix)
//                               B_NK: B[n*ldb + k]  (k contiguous, a column-major matrix)
constexpr int A_KM = 0, A_MK = 1;
constexpr int B_KN = 0, B_NK = 1;

template <int LA>
struct ATile {
  static constexpr int ld = (LA == A_KM) ? (BM + PAD) : (BK + PAD);
  static constexpr int elems = (LA == A_KM) ? BK * (BM + PAD) : BM * (BK + PAD);
};
template <int LB>
struct BTile {
  static constexpr int ld = (LB == B_KN) ? (BN + PAD) : (BK + PAD);
  static constexpr int elems = (LB == B_KN) ? BK * (BN + PAD) : BN * (BK + PAD);
};
template <int LA, int LB>
struct StageCfg {
  static constexpr int a_elems = ATile<LA>::elems;
  static constexpr int b_elems = BTile<LB>::elems;
  static constexpr int elems = a_elems + b_elems;
  static constexpr int bytes = elems * 8;
  // 4 stages when both operands are in the compact layout (25.6 KB / stage), else 3.
  static constexpr int stages = (LA == A_KM && LB == B_KN) ? 4 : 3;
  static constexpr int smem_bytes = stages * bytes;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// ---- TMA 1-D bulk copy + mbarrier (used to stage the X tile of a column tile) ----------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  unsigned s = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s), "r"(bytes));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned s = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(s), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::); }
// global -> shared bulk copy (TMA engine, SASS UBLKCP); bytes % 16 == 0, both addresses 16B aligned.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
               "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}

// ---- tile loaders (all 256 threads) -----------------------------------------------------------
// A thread copies NCH 16-byte chunks per tile; chunk i lives at (global) thread_ptr + i * gstep and (shared)
// tile + soff + i * SSTEP, so that per stage only the tile base changes (no index arithmetic in the main loop).
template <int LA>
struct ALoad {
  static constexpr int NCH = (BM * BK / 2) / NTHREADS;  // 4
  static constexpr int SSTEP = (LA == A_KM) ? (NTHREADS / (BM / 2)) * ATile<LA>::ld : (NTHREADS / (BK / 2)) * ATile<LA>::ld;
  int64_t goff, gstep;  // element offsets relative to the tile's (m0, k0) element
  int soff;
  __device__ __forceinline__ void init(int64_t lda, int tid) {
    if (LA == A_KM) {
      const int k = tid / (BM / 2), mc = tid % (BM / 2);
      goff = (int64_t)k * lda + mc * 2;
      gstep = (int64_t)(NTHREADS / (BM / 2)) * lda;
      soff = k * ATile<LA>::ld + mc * 2;
    } else {
      const int m = tid / (BK / 2), kc = tid % (BK / 2);
      goff = (int64_t)m * lda + kc * 2;
      gstep = (int64_t)(NTHREADS / (BK / 2)) * lda;
      soff = m * ATile<LA>::ld + kc * 2;
    }
  }
  __device__ __forceinline__ void load(double* sA, const double* __restrict__ gtile) const {
    const double* g = gtile + goff;
    double* s = sA + soff;
#pragma unroll
    for (int i = 0; i < NCH; i++) cp_async16(s + i * SSTEP, g + i * gstep);
  }
};
template <int LB>
struct BLoad {
  static constexpr int NCH = (BK * BN / 2) / NTHREADS;  // 2
  static constexpr int SSTEP = (LB == B_KN) ? (NTHREADS / (BN / 2)) * BTile<LB>::ld : (NTHREADS / (BK / 2)) * BTile<LB>::ld;
  int64_t goff, gstep;
  int soff;
  __device__ __forceinline__ void init(int64_t ldb, int tid) {
    if (LB == B_KN) {
      const int k = tid / (BN / 2), nc = tid % (BN / 2);
      goff = (int64_t)k * ldb + nc * 2;
      gstep = (int64_t)(NTHREADS / (BN / 2)) * ldb;
      soff = k * BTile<LB>::ld + nc * 2;
    } else {
      const int n = tid / (BK / 2), kc = tid % (BK / 2);
      goff = (int64_t)n * ldb + kc * 2;
      gstep = (int64_t)(NTHREADS / (BK / 2)) * ldb;
      soff = n * BTile<LB>::ld + kc * 2;
    }
  }
  __device__ __forceinline__ void load(double* sB, const double* __restrict__ gtile) const {
    const double* g = gtile + goff;
    double* s = sB + soff;
#pragma unroll
    for (int i = 0; i < NCH; i++) cp_async16(s + i * SSTEP, g + i * gstep);
  }
};

// Per-thread coordinates inside the CTA tile.
struct ThreadMap {
  int warp, lane, g, t, wm, wn;
  __device__ __forceinline__ ThreadMap() {
    int tid = threadIdx.x;
    warp = tid >> 5;
    lane = tid & 31;
    g = lane >> 2;
    t = lane & 3;
    // Row quarter of the warp: warps w and w+4 live on the same SM sub-partition (w mod 4); giving them the quarters
    // q and 3-q balances the sub-partitions when triangular operands let a warp skip the zero part of a diagonal block.
    wm = ((warp < 4) ? warp : 3 - (warp & 3)) * 32;
    wn = (warp >> 2) * 32;
  }
  // Triangular A operand, diagonal 128 x 128 block, k-step starting at column krel of that block: does this warp's
  // 32-row slice hold any non-zero?  lower: k <= m, upper: k >= m.
  __device__ __forceinline__ bool tri_active_lower(int krel) const { return krel < wm + 32; }
  __device__ __forceinline__ bool tri_active_upper(int krel) const { return krel + BK > wm; }
  // accumulator element acc[mi][ni][e] is C(row(mi), col(ni, e)) of the CTA tile
  __device__ __forceinline__ int row(int mi) const { return wm + mi * 8 + g; }
  __device__ __forceinline__ int col(int ni, int e) const { return wn + ni * 8 + 2 * t + e; }
};

typedef double Acc[4][4][2];

__device__ __forceinline__ void acc_zero(Acc& acc) {
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// One BK = 16 stage of DMMA work for this warp.
template <int LA, int LB>
__device__ __forceinline__ void mma_stage(Acc& acc, const double* __restrict__ sA, const double* __restrict__ sB,
                                          const ThreadMap& tm) {
  constexpr int lda = ATile<LA>::ld, ldb = BTile<LB>::ld;
#pragma unroll
  for (int kk = 0; kk < BK; kk += 4) {
    double a[4], b[4];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
      a[mi] = (LA == A_KM) ? sA[(kk + tm.t) * lda + tm.wm + mi * 8 + tm.g] : sA[(tm.wm + mi * 8 + tm.g) * lda + kk + tm.t];
#pragma unroll
    for (int ni = 0; ni < 4; ni++)
      b[ni] = (LB == B_KN) ? sB[(kk + tm.t) * ldb + tm.wn + ni * 8 + tm.g] : sB[(tm.wn + ni * 8 + tm.g) * ldb + kk + tm.t];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
      for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni], a[mi], b[ni]);
  }
}

// Deterministic block-wide sum of one double per thread (result valid in thread 0).
__device__ __forceinline__ double block_sum(double v, double* sred /* >= 8 doubles */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sred[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; w++) r += sred[w];
  }
  return r;
}

}  // namespace agp
