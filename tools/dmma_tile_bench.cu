// DMMA GEMM tile-shape microbenchmark (development aid): C[M x N] = A[M x K] * B[K x N], A column-major (m contiguous),
// B row-major (n contiguous) -- the operand layouts of the sweep GEMMs -- for several CTA / warp tile configurations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/dmma_tile_bench tools/dmma_tile_bench.cu
// Prints TFLOP/s per configuration (M = 1024, N = 37888, K = 1024 by default: one chunk of the C4 sweep).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// BM x BN CTA tile, WARPS_M x WARPS_N warps, BK = 16, S stages, MINB = CTAs per SM
template <int BM, int BN, int WARPS_M, int WARPS_N, int S, int MINB>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, MINB) gemm_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                                                                             int M, int N, int K) {
  constexpr int BK = 16, PAD = 4, NT = WARPS_M * WARPS_N * 32;
  constexpr int LDA = BM + PAD, LDB = BN + PAD;
  constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N, MI = WM / 8, NI = WN / 8;
  constexpr int STAGE = BK * LDA + BK * LDB;
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = (warp % WARPS_M) * WM, wn = (warp / WARPS_M) * WN;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const double* gA = A + m0;                 // element (m, k) at A[k*M + m]
  const double* gB = B + n0;                 // element (k, n) at B[k*N + n]
  constexpr int ACH = BK * BM / 2 / NT, BCH = BK * BN / 2 / NT;  // 16-byte chunks per thread
  auto load = [&](int slot, int kstep) {
    double* sA = smem + slot * STAGE;
    double* sB = sA + BK * LDA;
#pragma unroll
    for (int i = 0; i < ACH; i++) {
      const int c = tid + i * NT, k = c / (BM / 2), mc = c % (BM / 2);
      cp_async16(sA + k * LDA + mc * 2, gA + (int64_t)(kstep * BK + k) * M + mc * 2);
    }
#pragma unroll
    for (int i = 0; i < BCH; i++) {
      const int c = tid + i * NT, k = c / (BN / 2), nc = c % (BN / 2);
      cp_async16(sB + k * LDB + nc * 2, gB + (int64_t)(kstep * BK + k) * N + nc * 2);
    }
  };
  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int nsteps = K / BK;
#pragma unroll
  for (int s = 0; s < S - 1; s++) {
    if (s < nsteps) load(s, s);
    cp_async_commit();
  }
  for (int step = 0; step < nsteps; step++) {
    cp_async_wait<S - 2>();
    __syncthreads();
    if (step + S - 1 < nsteps) load((step + S - 1) % S, step + S - 1);
    cp_async_commit();
    const double* sA = smem + (step % S) * STAGE;
    const double* sB = sA + BK * LDA;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[MI], b[NI];
#pragma unroll
      for (int i = 0; i < MI; i++) a[i] = sA[(kk + t) * LDA + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < NI; j++) b[j] = sB[(kk + t) * LDB + wn + j * 8 + g];
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) dmma884(acc[i][j], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++)
      *reinterpret_cast<double2*>(C + (int64_t)(m0 + wm + i * 8 + g) * N + n0 + wn + j * 8 + 2 * t) = make_double2(acc[i][j][0], acc[i][j][1]);
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int S, int MINB>
void run(const char* name, const double* A, const double* B, double* C, int M, int N, int K) {
  constexpr int smem = S * (16 * (BM + 4) + 16 * (BN + 4)) * 8;
  auto kern = gemm_kernel<BM, BN, WARPS_M, WARPS_N, S, MINB>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(M / BM, N / BN);
  for (int i = 0; i < 2; i++) kern<<<grid, WARPS_M * WARPS_N * 32, smem>>>(A, B, C, M, N, K);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  const int reps = 5;
  for (int i = 0; i < reps; i++) kern<<<grid, WARPS_M * WARPS_N * 32, smem>>>(A, B, C, M, N, K);
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS_M * WARPS_N * 32, smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  printf("{\"config\": \"%s\", \"ms\": %.3f, \"tflops\": %.2f, \"ctas_per_sm\": %d, \"regs\": %d, \"smem_kb\": %.1f}\n", name, ms, 2.0 * M * N * K / ms * 1e-9, occ,
         fa.numRegs, smem / 1024.0);
}

int main(int argc, char** argv) {
  const int M = 1024, N = 37888, K = 1024;
  double *A, *B, *C;
  CK(cudaMalloc(&A, sizeof(double) * M * K));
  CK(cudaMalloc(&B, sizeof(double) * (size_t)K * N));
  CK(cudaMalloc(&C, sizeof(double) * (size_t)M * N));
  CK(cudaMemset(A, 0, sizeof(double) * M * K));
  CK(cudaMemset(B, 0, sizeof(double) * (size_t)K * N));
  //            BM   BN  WM WN  S  CTAs/SM
  run<128, 64, 4, 2, 4, 2>("128x64 cta, 8 warps of 32x32, 4 stages, 2/SM (current)", A, B, C, M, N, K);
  run<128, 64, 2, 2, 4, 2>("128x64 cta, 4 warps of 64x32, 4 stages, 2/SM", A, B, C, M, N, K);
  run<128, 64, 2, 2, 4, 3>("128x64 cta, 4 warps of 64x32, 4 stages, 3/SM", A, B, C, M, N, K);
  run<128, 128, 4, 2, 4, 1>("128x128 cta, 8 warps of 32x64, 4 stages, 1/SM", A, B, C, M, N, K);
  run<128, 128, 2, 4, 4, 1>("128x128 cta, 8 warps of 64x32, 4 stages, 1/SM", A, B, C, M, N, K);
  run<128, 128, 4, 2, 3, 1>("128x128 cta, 8 warps of 32x64, 3 stages, 1/SM", A, B, C, M, N, K);
  run<128, 128, 4, 4, 4, 1>("128x128 cta, 16 warps of 32x32, 4 stages, 1/SM", A, B, C, M, N, K);
  run<256, 64, 4, 2, 3, 1>("256x64 cta, 8 warps of 64x32, 3 stages, 1/SM", A, B, C, M, N, K);
  run<64, 64, 2, 2, 4, 4>("64x64 cta, 4 warps of 32x32, 4 stages, 4/SM", A, B, C, M, N, K);
  return 0;
}
