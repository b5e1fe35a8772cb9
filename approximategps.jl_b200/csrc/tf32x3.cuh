// tf32x3.cuh -- the Float32 fast mode's GEMM engine: 3xTF32 split products on the 5th-generation tensor cores.
//
//   D[128 x 128] (FP32, in TMEM)  =  sum_k  Ah Bh^T + Ah Bl^T + Al Bh^T        x = xh + xl,  xh = tf32(x),  xl = tf32(x - xh)
//
// i.e. an FP32-accurate product (22 mantissa bits per operand, FP32 accumulation) at a third of the TF32 rate -- still ~10x the
// FP64 DMMA rate that bounds the reference-precision path.  One CTA computes one 128 x 128 output tile:
//   warp 0      TMA producer: cp.async.bulk.tensor (128-byte swizzle) of the four operand tiles of a k-block into a 3-stage ring
//   warp 1      allocates TMEM (two 128-column buffers), issues tcgen05.mma.kind::tf32 (one elected thread) in groups of 64 k,
//               alternating the buffers, commits stages back to the producer and groups to the epilogue
//   warps 2..9  epilogue: tcgen05.ld of each finished group (one TMEM lane = one output row, 64 columns per thread), summed over
//               the groups in FP64 registers (the tensor core's FP32 accumulator rounds toward zero: without this carry a k = 1024
//               product has a systematic relative bias of 1e-5); then the functor and the global stores
// Operands are read in place from the hi / lo FP32 planes the producing kernels write; both K-major (k contiguous) and MN-major
// (m / n contiguous) operands are supported, so that one point-major copy of every intermediate serves all four sweep stages
// (DESIGN.md section 8).  SASS: UTCHMMA-class tensor instructions, LDTM, UTMALDG.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace agp {
namespace t5 {

constexpr int TM = 128;      // output rows per CTA = TMEM lanes
constexpr int TN = 128;      // output columns per CTA = TMEM columns
constexpr int TK = 32;       // k per pipeline stage: 32 floats = one 128-byte swizzle row
constexpr int UK = 8;        // k of one tcgen05.mma.kind::tf32
constexpr int STAGES = 3;
constexpr int TILE_BYTES = TM * TK * 4;        // 16 KB: one operand plane of one stage
constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // Ah | Al | Bh | Bl
constexpr int EPI_WARPS = 8;      // two per TMEM lane quarter, 64 accumulator columns each
constexpr int T5_THREADS = 64 + 32 * EPI_WARPS;
constexpr int GROUP_KB = 2;       // k-blocks per accumulation group (64 k = 24 accumulator updates), then carried over in FP64
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

// k-range of an output tile
constexpr int KM_FULL = 0;     // [0, K)
constexpr int KM_FROM_N = 1;   // [tile_n * TN, K)          the N-side operand is lower triangular stored as [n][k], k >= n
constexpr int KM_UPTO_N = 2;   // [0, (tile_n + 1) * TN)    ...                                                       k <= n
constexpr int KM_SPLIT = 3;    // [z * kchunk, min(K, (z + 1) * kchunk)), z = blockIdx.z


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst), "l"(map),
               "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void mbar_init_u32(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, one 128 x 128 x 8 TF32 MMA.  The issuing warp walks its loops convergently and `leader` (one elected
// lane) predicates the instruction itself: with a divergent `if (lane == 0)` around the loop the compiler wraps every uniform-datapath
// instruction (descriptor arithmetic, the MMA, the commit) in an ELECT / BRA.U.ANY retry loop -- 13 instructions and ~95 cycles per MMA,
// measured on the INT8 engine that shares this structure (profiles/r2q_i8emu_v0.jsonl -> v1).
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.ne.b32 q, %5, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pred(uint32_t bar, uint32_t leader) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %1, 0;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar),
      "r"(leader)
      : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t r;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(r));
  return r;
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
      "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
      "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), version 1
//   K-major  tile [rows][32 floats], 128-byte swizzle (layout type 2): rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused (1)
//   MN-major tile: 4 boxes of [32 k][32 floats of m / n].  For 32-bit operands the only MN-major layout the tensor core accepts is
//            the 128-byte swizzle with a 32-byte atom (layout type 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; CUTLASS
//            Layout_MN_SW128_32B_Atom): swizzle atoms of 4 k-rows x 128 B, so SBO = 512 B between the 4-k groups and
//            LBO = 4096 B between the 32-wide m / n groups (one TMA box each)
struct MnDesc {
  uint32_t lbo16 = 4096 >> 4, sbo16 = 512 >> 4, layout = 1;
};
struct Args {
  int K;
  int kmode;
  int kchunk;
  int lower_only;  // 1: skip tiles with tile_n > tile_m (symmetric output)
  MnDesc mn;       // descriptor fields of MN-major operands (fixed; a parameter only so that tools/tf32x3_test.cu can probe them)
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major, const MnDesc& mn) {
  const uint64_t lbo = mn_major ? mn.lbo16 : 1, sbo = mn_major ? mn.sbo16 : (1024 >> 4), lt = mn_major ? mn.layout : 2;
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (lt << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, M = 128, N = 128, dense, no negate
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// Epi: struct State (per-thread registers);  begin(State&, tile_m, tile_n, z, row, half);  end(State&, ..., half);
//      operator()(State&, tile_m, tile_n, z, row, c0, const double (&v)[32]) -- a thread owns output row `row` (0..127 of the tile)
//      and the column half `half` (0: columns 0..63, 1: 64..127); it is called for c0 = 64 half and 64 half + 32, bracketed by
//      begin / end (per-row reductions: two partial results per row, one per half).
template <bool AMN, bool BMN, class Epi>
__global__ void __launch_bounds__(T5_THREADS, 1)
tf32x3_gemm_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl, const __grid_constant__ CUtensorMap mBh,
                   const __grid_constant__ CUtensorMap mBl, Args g, Epi epi) {
  extern __shared__ uint8_t t5_smem_raw[];
  // launch grid = (m tiles, n tiles, z); the tiles are walked with n as the FAST index: the n-tile CTAs that share an A row tile (the
  // streamed operand: two planes of 128 x K floats) run together and read it from HBM once, through L2.  With m fast every n tile streamed
  // all of A again -- 8 x 1.24 GB per launch at C4, i.e. the stage was DRAM-bound (1.5 of its 1.8 ms).
  const int lin = blockIdx.x + gridDim.x * blockIdx.y;
  const int tile_n = lin % gridDim.y, tile_m = lin / gridDim.y, z = blockIdx.z;
  if (g.lower_only && tile_n > tile_m) return;
  int kb = 0, ke = g.K;
  if (g.kmode == KM_FROM_N) kb = tile_n * TN;
  if (g.kmode == KM_UPTO_N) ke = min(g.K, (tile_n + 1) * TN);
  if (g.kmode == KM_SPLIT) {
    kb = z * g.kchunk;
    ke = min(g.K, kb + g.kchunk);
  }
  const int nk = (ke - kb + TK - 1) / TK;          // k-blocks of 32
  const int ngroups = (nk + GROUP_KB - 1) / GROUP_KB;  // accumulation groups: one TMEM buffer each, carried over in FP64

  const uint32_t base = (smem_u32(t5_smem_raw) + 1023u) & ~1023u;  // the 128-byte swizzle pattern repeats every 1024 B
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (2 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(t5_smem_raw + (tmem_slot - smem_u32(t5_smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init_u32(full_bar(s), 1);
      mbar_init_u32(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init_u32(tfull_bar(b), 1);
      mbar_init_u32(tempty_bar(b), EPI_WARPS);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tmem_slot), "r"((uint32_t)(2 * TN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0 && nk > 0) {
      // ---- TMA producer ---------------------------------------------------------------------------------------------
      for (int i = 0; i < nk; i++) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait_u32(empty_bar(s), ((i / STAGES) - 1) & 1);
        const uint32_t st = base + s * STAGE_BYTES;
        mbar_expect_tx_u32(full_bar(s), STAGE_BYTES);
        const int k0 = kb + i * TK;
        if (AMN) {
#pragma unroll
          for (int b = 0; b < 4; b++) {
            tma_load_2d(st + b * 4096, &mAh, full_bar(s), tile_m * TM + 32 * b, k0);
            tma_load_2d(st + TILE_BYTES + b * 4096, &mAl, full_bar(s), tile_m * TM + 32 * b, k0);
          }
        } else {
          tma_load_2d(st, &mAh, full_bar(s), k0, tile_m * TM);
          tma_load_2d(st + TILE_BYTES, &mAl, full_bar(s), k0, tile_m * TM);
        }
        if (BMN) {
#pragma unroll
          for (int b = 0; b < 4; b++) {
            tma_load_2d(st + 2 * TILE_BYTES + b * 4096, &mBh, full_bar(s), tile_n * TN + 32 * b, k0);
            tma_load_2d(st + 3 * TILE_BYTES + b * 4096, &mBl, full_bar(s), tile_n * TN + 32 * b, k0);
          }
        } else {
          tma_load_2d(st + 2 * TILE_BYTES, &mBh, full_bar(s), k0, tile_n * TN);
          tma_load_2d(st + 3 * TILE_BYTES, &mBl, full_bar(s), k0, tile_n * TN);
        }
      }
    }
  } else if (warp == 1) {
    if (nk > 0) {
      // ---- MMA issuer (whole warp, uniform control flow; one elected lane issues): group gi accumulates GROUP_KB k-blocks into TMEM
      //      buffer gi & 1, starting from zero ----------------------------------------------------------------------------------------
      const uint32_t leader = elect_one();
      constexpr uint32_t idesc = make_idesc(AMN, BMN);
      constexpr uint32_t a_step = AMN ? 1024 : UK * 4, b_step = BMN ? 1024 : UK * 4;  // bytes per k-step of 8
      int i = 0;
      for (int gi = 0; gi < ngroups; gi++) {
        const int buf = gi & 1;
        if (gi >= 2) {  // the epilogue warps have drained this buffer's previous group
          mbar_wait_u32(tempty_bar(buf), ((gi >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t dcol = tmem_d + (uint32_t)(buf * TN);
        const int iend = min(nk, i + GROUP_KB);
        for (bool first = true; i < iend; i++) {
          const int s = i % STAGES;
          mbar_wait_u32(full_bar(s), (i / STAGES) & 1);
          tc_fence_after();
          const uint32_t st = base + s * STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < TK / UK; k++) {
            const uint64_t ah = make_desc(st + k * a_step, AMN, g.mn), al = make_desc(st + TILE_BYTES + k * a_step, AMN, g.mn);
            const uint64_t bh = make_desc(st + 2 * TILE_BYTES + k * b_step, BMN, g.mn), bl = make_desc(st + 3 * TILE_BYTES + k * b_step, BMN, g.mn);
            tc_mma_tf32(dcol, al, bh, idesc, first ? 0u : 1u, leader);  // the small terms first
            tc_mma_tf32(dcol, ah, bl, idesc, 1, leader);
            tc_mma_tf32(dcol, ah, bh, idesc, 1, leader);
            first = false;
          }
          tc_commit_pred(empty_bar(s), leader);  // the stage may be refilled once these MMAs have read it
        }
        tc_commit_pred(tfull_bar(buf), leader);  // this group's partial sums are complete
      }
    }
  } else {
    // ---- epilogue: warp w may only touch TMEM lanes 32 (w % 4) .. 32 (w % 4) + 31; two warps share a lane quarter, 64 columns each.
    // The FP32 accumulator of the tensor core rounds toward zero on every update (a relative bias of ~6e-8 per update, 1e-5 over
    // the 384 updates of a k = 1024 product: measured as a systematic -1e-5 on |c|^2, i.e. on the marginal variances).  So a TMEM
    // buffer only ever holds GROUP_KB * 12 updates; the groups are summed here in FP64 registers.
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    double acc[64];
#pragma unroll
    for (int j = 0; j < 64; j++) acc[j] = 0.0;
    for (int gi = 0; gi < ngroups; gi++) {
      const int buf = gi & 1;
      mbar_wait_u32(tfull_bar(buf), (gi >> 1) & 1);
      tc_fence_after();
      float v0[32], v1[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TN + half * 64);
      tc_ld32_nowait(taddr, v0);
      tc_ld32_nowait(taddr + 32, v1);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tempty_bar(buf)) : "memory");
#pragma unroll
      for (int j = 0; j < 32; j++) {
        acc[j] += (double)v0[j];
        acc[32 + j] += (double)v1[j];
      }
    }
    typename Epi::State st;
    epi.begin(st, tile_m, tile_n, z, row, half);
    epi(st, tile_m, tile_n, z, row, half * 64, *reinterpret_cast<const double(*)[32]>(acc));
    epi(st, tile_m, tile_n, z, row, half * 64 + 32, *reinterpret_cast<const double(*)[32]>(acc + 32));
    epi.end(st, tile_m, tile_n, z, row, half);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"((uint32_t)(2 * TN)) : "memory");
  }
}

// ---- host side: tensor maps -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}
// FP32 matrix with `inner` contiguous elements per row and `outer` rows (row pitch `ld` elements); a box is 32 inner elements
// (128 bytes, the swizzle width) x box_outer rows.  K-major operand: inner = k, box_outer = 128.  MN-major: inner = m / n, 32.
inline bool make_map(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer, bool mn_major = false) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {32, box_outer};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// x -> (tf32(x), tf32(x - tf32(x))): both parts exactly representable in TF32 (round to nearest), so the tensor core's truncation
// of the low 13 mantissa bits loses nothing
__device__ __forceinline__ void split_tf32(double x, float& hi, float& lo) {
  const float f = (float)x;
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(h) : "f"(f));
  hi = __uint_as_float(h);
  const float r = (float)(x - (double)hi);
  uint32_t l;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(l) : "f"(r));
  lo = __uint_as_float(l);
}

// the same for an FP32 value (the accumulator of a previous product)
__device__ __forceinline__ void split_tf32f(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(l) : "f"(x - hi));
  lo = __uint_as_float(l);
}

}  // namespace t5
}  // namespace agp
