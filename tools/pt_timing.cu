// Phase timing of the 128 x 128 diagonal-block kernel (clock64 ticks), development aid.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DAGP_PT_TIMING -o build/pt_timing tools/pt_timing.cu
#include <cstdio>
#include <vector>
#include "../approximategps.jl_b200/csrc/dense.cuh"
using namespace agp;
int main() {
  const int n = 128, ld = 8192;
  std::vector<double> A((size_t)ld * n, 0.0);
  for (int c = 0; c < n; c++)
    for (int r = 0; r < n; r++) A[(size_t)c * ld + r] = (r == c ? 2.0 : 0.0) + exp(-0.01 * (r - c) * (r - c));
  double *dA, *dL, *dLt, *dUt; int* info;
  cudaMalloc(&dA, sizeof(double) * ld * n); cudaMalloc(&dL, sizeof(double) * ld * n); cudaMalloc(&dLt, sizeof(double) * ld * n); cudaMalloc(&dUt, sizeof(double) * ld * n);
  cudaMalloc(&info, 16); cudaMemset(info, 0, 16);
  cudaMemcpy(dA, A.data(), sizeof(double) * ld * n, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(potrf_trinv128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM_BYTES);
  for (int it = 0; it < 3; it++) potrf_trinv128_kernel<<<1, 256, PT_SMEM_BYTES>>>(dA, dL, dLt, dUt, ld, 0, info, 128);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int it = 0; it < 10; it++) potrf_trinv128_kernel<<<1, 256, PT_SMEM_BYTES>>>(dA, dL, dLt, dUt, ld, 0, info, 128);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long t[16]; cudaMemcpyFromSymbol(t, pt_ticks, sizeof t);
  printf("kernel %.1f us (err %s)\n", 1e3 * ms / 10, cudaGetErrorString(cudaGetLastError()));
  const char* names[] = {"load", "diag32(o=0)", "panel(o=0)", "trailing(o=0)", "rest of chol", "store L", "inv diag", "offdiag 32 x2", "offdiag 64", "store inv"};
  for (int i = 1; i <= 9; i++) printf("  %-16s %8.2f us\n", names[i], (t[i] - t[i - 1]) / 1965.0);
  // residual check
  std::vector<double> L((size_t)ld * n); cudaMemcpy(L.data(), dL, sizeof(double) * ld * n, cudaMemcpyDeviceToHost);
  double err = 0; for (int r = 0; r < n; r++) for (int c = 0; c <= r; c++) { double s = 0; for (int k = 0; k <= c; k++) s += L[(size_t)k * ld + r] * L[(size_t)k * ld + c]; err = fmax(err, fabs(s - A[(size_t)c * ld + r])); }
  printf("max |L L^T - A| = %.2e\n", err);
  return 0;
}
