// Dependent-chain latencies of the FP64 / shuffle / shared-memory operations on the critical path of the diagonal-block Cholesky (one warp).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/lat_bench tools/lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ double g_sink[64];
__device__ long long g_t[16];
template <int OP>
__global__ void chain(double x0, double y0, int n) {
  __shared__ double sm[64];
  double x = x0 + threadIdx.x * 1e-9, y = y0;
  sm[threadIdx.x & 63] = x;
  __syncthreads();
  int idx = threadIdx.x & 31;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (OP == 0) x = fma(x, y, y);                                   // DFMA
      if (OP == 1) x = x * y;                                          // DMUL
      if (OP == 2) x = x + y;                                          // DADD
      if (OP == 3) x = __shfl_sync(0xffffffffu, x, (idx + 1) & 31);    // 64-bit shuffle (2 SHFL)
      if (OP == 4) x = rsqrt(x) + 1.5;                                 // rsqrt + DADD
      if (OP == 5) x = __drcp_rn(x) + 1.5;                             // reciprocal + DADD
      if (OP == 6) x = 1.0 / x + 1.5;                                  // division + DADD
      if (OP == 7) { idx = (int)sm[idx] & 31; }                        // dependent LDS.64 (+ F2I)
      if (OP == 8) x = sqrt(x) + 1.5;                                  // sqrt + DADD
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) g_t[OP] = t1 - t0;
  g_sink[threadIdx.x & 63] = x + idx;
}
int main() {
  const char* names[] = {"DFMA", "DMUL", "DADD", "shfl64", "rsqrt+DADD", "drcp_rn+DADD", "div+DADD", "LDS+F2I", "sqrt+DADD"};
  const int n = 64;
  chain<0><<<1, 32>>>(1.0, 0.999, n); chain<1><<<1, 32>>>(1.0, 0.999, n); chain<2><<<1, 32>>>(1.0, 0.999, n); chain<3><<<1, 32>>>(1.0, 0.999, n);
  chain<4><<<1, 32>>>(1.3, 0.999, n); chain<5><<<1, 32>>>(1.3, 0.999, n); chain<6><<<1, 32>>>(1.3, 0.999, n); chain<7><<<1, 32>>>(1.3, 0.999, n);
  chain<8><<<1, 32>>>(1.3, 0.999, n);
  cudaDeviceSynchronize();
  long long t[16];
  cudaMemcpyFromSymbol(t, g_t, sizeof t);
  for (int i = 0; i < 9; i++) printf("{\"op\": \"%s\", \"cycles_per_op\": %.1f}\n", names[i], (double)t[i] / (16.0 * n));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
