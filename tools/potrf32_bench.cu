// One-warp 32 x 32 Cholesky variants in isolation (warm, 50 repetitions each): what bounds the per-column time of the diagonal-block kernel?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/potrf32_bench tools/potrf32_bench.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "../approximategps.jl_b200/csrc/dense.cuh"
using namespace agp;
__device__ long long g_t[16];
__device__ double g_sink[64];

// V2: sliding window, broadcast of the scaled column through shared memory instead of shuffles
__device__ __forceinline__ void potrf32_lds(double* blk, int lane, double* rdiag, double* buf) {
  double x[32];
#pragma unroll
  for (int c = 0; c < 32; c++) x[c] = blk[lane * PT_LD + c];
#pragma unroll 1
  for (int j = 0; j < 32; j++) {
    const double v = x[0];
    const double d = __shfl_sync(0xffffffffu, v, j);
    const double y = rsqrt(d);
    const double s0 = d * y;
    const double sq = fma(fma(-s0, s0, d), 0.5 * y, s0);
    const double inv = fma(fma(-sq, y, 1.0), y, y);
    const double l0 = v * inv;
    const double lij = (lane == j) ? sq : fma(fma(-l0, sq, v), inv, l0);
    buf[lane] = lij;
    buf[lane + 32] = lij;
    __syncwarp();
    const double* b = buf + j;
#pragma unroll
    for (int k = 1; k < 32; k++) x[k - 1] = fma(-lij, b[k], x[k]);
    if (lane >= j) blk[lane * PT_LD + j] = lij;
    if (lane == j) rdiag[j] = inv;
    __syncwarp();
  }
}
// V3: chain only (no exchange): lower bound of the pivot arithmetic
__device__ __forceinline__ void potrf32_chain(double* blk, int lane, double* rdiag) {
  double x[32];
#pragma unroll
  for (int c = 0; c < 32; c++) x[c] = blk[lane * PT_LD + c];
#pragma unroll 1
  for (int j = 0; j < 32; j++) {
    const double v = x[0];
    const double d = __shfl_sync(0xffffffffu, v, j);
    const double y = rsqrt(fabs(d) + 1.0);
    const double s0 = d * y;
    const double sq = fma(fma(-s0, s0, d), 0.5 * y, s0);
    const double inv = fma(fma(-sq, y, 1.0), y, y);
    const double l0 = v * inv;
    const double lij = (lane == j) ? sq : fma(fma(-l0, sq, v), inv, l0);
#pragma unroll
    for (int k = 1; k < 32; k++) x[k - 1] = fma(-lij, lij, x[k]);
    if (lane >= j) blk[lane * PT_LD + j] = lij;
    if (lane == j) rdiag[j] = inv;
  }
}
// V4: sliding window with 32-bit-pair shuffles replaced by one shuffle per step of a PACKED exchange: every lane first publishes lij, then
// reads the 31 partners with LDS.128 from a 16-byte aligned doubled buffer (two aligned copies: even and odd start)
__device__ __forceinline__ void potrf32_lds128(double* blk, int lane, double* rdiag, double* buf) {
  double x[32];
#pragma unroll
  for (int c = 0; c < 32; c++) x[c] = blk[lane * PT_LD + c];
  double* bufe = buf;        // bufe[i] = l_i            (i = 0..63, wraps)
  double* bufo = buf + 64;   // bufo[i] = l_(i+1)        so that an odd start is 16-byte aligned here
#pragma unroll 1
  for (int j = 0; j < 32; j++) {
    const double v = x[0];
    const double d = __shfl_sync(0xffffffffu, v, j);
    const double y = rsqrt(d);
    const double s0 = d * y;
    const double sq = fma(fma(-s0, s0, d), 0.5 * y, s0);
    const double inv = fma(fma(-sq, y, 1.0), y, y);
    const double l0 = v * inv;
    const double lij = (lane == j) ? sq : fma(fma(-l0, sq, v), inv, l0);
    bufe[lane] = lij;
    bufe[lane + 32] = lij;
    bufo[(lane + 31) & 31] = lij;
    bufo[((lane + 31) & 31) + 32] = lij;
    __syncwarp();
    // partners l_(j+1) .. l_(j+31): start index j+1; if j+1 is even read bufe + j + 1, else bufo + j (bufo[j] = l_(j+1)), both 16-byte aligned
    const double2* b = reinterpret_cast<const double2*>(((j + 1) & 1) ? (bufo + j) : (bufe + j + 1));
#pragma unroll
    for (int k = 0; k < 15; k++) {
      const double2 p = b[k];
      x[2 * k] = fma(-lij, p.x, x[2 * k + 1]);
      x[2 * k + 1] = fma(-lij, p.y, x[2 * k + 2]);
    }
    {
      const double2 p = b[15];
      x[30] = fma(-lij, p.x, x[31]);
    }
    if (lane >= j) blk[lane * PT_LD + j] = lij;
    if (lane == j) rdiag[j] = inv;
    __syncwarp();
  }
}

template <int V>
__global__ void run(const double* src, double* out, int reps) {
  __shared__ double blk[32 * PT_LD];
  __shared__ double rdiag[32];
  __shared__ __align__(16) double buf[128];
  __shared__ int info;
  const int lane = threadIdx.x & 31;
  long long tot = 0;
  for (int r = 0; r < reps; r++) {
    __syncthreads();  // with blockDim.x = 256 the other seven warps wait here, as they do in the diagonal-block kernel
    if (threadIdx.x >= 32) continue;
    for (int c = 0; c < 32; c++) blk[lane * PT_LD + c] = (c <= lane) ? src[c * 32 + lane] : 0.0;
    __syncwarp();
    const long long t0 = clock64();
    if (V == 0) pt_potrf32(blk, lane, rdiag, 0, &info);
    if (V == 1) pt_potrf32r(blk, lane, rdiag, 0, &info);
    if (V == 2) potrf32_lds(blk, lane, rdiag, buf);
    if (V == 3) potrf32_chain(blk, lane, rdiag);
    if (V == 4) potrf32_lds128(blk, lane, rdiag, buf);
    __syncwarp();
    const long long t1 = clock64();
    if (r >= 2) tot += t1 - t0;
  }
  if (threadIdx.x >= 32) return;
  if (lane == 0) g_t[V + (blockDim.x > 32 ? 8 : 0)] = tot / (reps - 2);
  for (int c = 0; c < 32; c++) out[c * 32 + lane] = blk[lane * PT_LD + c];
}

int main() {
  const int n = 32;
  std::vector<double> A(n * n);
  for (int c = 0; c < n; c++)
    for (int r = 0; r < n; r++) A[c * n + r] = (r == c ? 2.0 : 0.0) + exp(-0.01 * (r - c) * (r - c));
  double *dA, *dL;
  cudaMalloc(&dA, sizeof(double) * n * n); cudaMalloc(&dL, sizeof(double) * n * n * 8);
  cudaMemcpy(dA, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
  run<0><<<1, 32>>>(dA, dL + 0 * n * n, 50);
  run<1><<<1, 32>>>(dA, dL + 1 * n * n, 50);
  run<2><<<1, 32>>>(dA, dL + 2 * n * n, 50);
  run<3><<<1, 32>>>(dA, dL + 3 * n * n, 50);
  run<4><<<1, 32>>>(dA, dL + 4 * n * n, 50);
  run<0><<<1, 256>>>(dA, dL + 5 * n * n, 50);
  run<1><<<1, 256>>>(dA, dL + 5 * n * n, 50);
  run<4><<<1, 256>>>(dA, dL + 5 * n * n, 50);
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  long long t[16]; cudaMemcpyFromSymbol(t, g_t, sizeof t);
  const char* names[] = {"smem left-looking (pt_potrf32)", "sliding window + shuffles (pt_potrf32r)", "sliding window + LDS.64 broadcast", "chain only (no exchange)", "sliding window + LDS.128 broadcast"};
  std::vector<double> L(n * n * 5); cudaMemcpy(L.data(), dL, sizeof(double) * n * n * 5, cudaMemcpyDeviceToHost);
  for (int v = 0; v < 5; v++) {
    double err = 0;
    const double* l = L.data() + v * n * n;
    for (int r = 0; r < n; r++) for (int c = 0; c <= r; c++) { double s = 0; for (int k = 0; k <= c; k++) s += l[k * n + r] * l[k * n + c]; err = fmax(err, fabs(s - A[c * n + r])); }
    printf("{\"variant\": \"%s\", \"cycles\": %lld, \"cycles_per_column\": %.1f, \"us_at_1965MHz\": %.2f, \"residual\": %.2e}\n", names[v], t[v], t[v] / 32.0, t[v] / 1965.0, err);
  }
  printf("{\"with seven warps waiting at the block barrier\": {\"smem left-looking\": %lld, \"shuffles\": %lld, \"LDS.128\": %lld}}\n", t[8], t[9], t[12]);
  return 0;
}
