// Stand-alone HBM evidence for the per-point stage (S3, sweep.cuh perpoint_kernel): one launch over N points with the C4 operand
// shapes (nb = 8 partial |c|^2 rows, Poisson analytic / Bernoulli GH-20), timed with CUDA events.  Inside the sweep the same kernel
// only sees 151 552 points per launch and is launch-latency-bound; this shows what it does when given enough work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/perpoint_bench tools/perpoint_bench.cu
#include <cstdio>
#include <vector>
#include "../approximategps.jl_b200/csrc/sweep.cuh"
using namespace agp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void fill_k(double* p, size_t n, double a, double b) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = a + b * (double)(i % 97) / 97.0;
}

int main() {
  const int N = 10'000'000 / 64 * 64, nb = 8, D = 8;
  double *saa, *sam, *scc, *y, *dmu, *dv, *part;
  int* flag;
  CK(cudaMalloc(&saa, 8ull * N)); CK(cudaMalloc(&sam, 8ull * N)); CK(cudaMalloc(&scc, 8ull * N * nb)); CK(cudaMalloc(&y, 8ull * N));
  CK(cudaMalloc(&dmu, 8ull * N)); CK(cudaMalloc(&dv, 8ull * N)); CK(cudaMalloc(&part, 8ull * (N / 256 + 2) * NSC)); CK(cudaMalloc(&flag, 16));
  fill_k<<<1024, 256>>>(saa, N, 0.2, 0.1); fill_k<<<1024, 256>>>(sam, N, -0.3, 0.6); fill_k<<<1024, 256>>>(scc, (size_t)N * nb, 0.01, 0.02);
  fill_k<<<1024, 256>>>(y, N, 0.0, 3.0);
  CK(cudaMemset(flag, 0, 16));
  double h_x[AGP_MAX_GH_POINTS] = {0}, h_w[AGP_MAX_GH_POINTS] = {0};
  for (int i = 0; i < 20; i++) { h_x[i] = -5.0 + 0.5 * i; h_w[i] = 0.05; }  // timing only
  CK(cudaMemcpyToSymbol(c_gh_x, h_x, sizeof h_x)); CK(cudaMemcpyToSymbol(c_gh_w, h_w, sizeof h_w));
  PerPointArgs p{};
  p.saa = saa; p.sam = sam; p.scc_part = scc; p.ldp = N; p.nb = nb; p.pts = nullptr; p.y = y; p.npts = N; p.ncols = N; p.scale = 1.0; p.mean_const = 0.0;
  p.kp.kind = AGP_KERNEL_SE; p.kp.D = D; p.kp.M = 1024; p.kp.variance = 1.0;
  p.dmu = dmu; p.dv = dv; p.sc_part = part; p.flag = flag; p.predict_only = 0;
  const double bytes = 8.0 * (5 + nb) * N;
  for (int mode = 0; mode < 2; mode++) {
    p.lp.kind = mode == 0 ? AGP_LIK_POISSON_EXP : AGP_LIK_BERNOULLI_LOGIT;
    p.lp.method = mode == 0 ? AGP_EXPECT_ANALYTIC : AGP_EXPECT_GAUSS_HERMITE;
    p.lp.ngh = 20;
    for (int i = 0; i < 3; i++) perpoint_kernel<<<(N + PP_POINTS_PER_BLOCK - 1) / PP_POINTS_PER_BLOCK, 256>>>(p);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) perpoint_kernel<<<(N + PP_POINTS_PER_BLOCK - 1) / PP_POINTS_PER_BLOCK, 256>>>(p);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("{\"kernel\": \"perpoint_kernel\", \"likelihood\": \"%s\", \"points\": %d, \"ms\": %.4f, \"alg_bytes_per_point\": %d, \"GB_per_s\": %.1f, \"frac_of_6550\": %.3f}\n",
           mode == 0 ? "Poisson analytic" : "Bernoulli Gauss-Hermite 20", N, ms, 8 * (5 + nb), bytes / ms * 1e-6, bytes / ms * 1e-6 / 6550.1);
  }
  return 0;
}
