// FP64 peak microbenchmarks for B200 (sm_100a): DFMA vs DMMA.8x8x4 issue rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peak tools/fp64_peak.cu
// Prints one JSON line per measurement. Used only to choose the MMA path and to
// obtain the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double s) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9 + i;
  double a = s, b = 1.0 - s;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_kernel(double* out, int iters, double s) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  double a = s + threadIdx.x * 1e-12, b = 1e-3 * s;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) r += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// DMMA with distinct A/B registers per MMA in a 64x32 warp-tile pattern (8 A frags x 4 B frags),
// operands re-read from shared memory every k-step: approximates a real GEMM mainloop.
__global__ void __launch_bounds__(256) dmma_smem_kernel(double* out, int iters) {
  __shared__ double sA[2][4][64 + 4 * 17];   // not a real layout; just LDS.64 traffic
  __shared__ double sB[2][4][32 + 4 * 17];
  int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int i = threadIdx.x; i < 2 * 4 * (64 + 68); i += blockDim.x) (&sA[0][0][0])[i] = 1e-3 * i;
  for (int i = threadIdx.x; i < 2 * 4 * (32 + 68); i += blockDim.x) (&sB[0][0][0])[i] = 1e-3 * i;
  __syncthreads();
  double c[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
  for (int it = 0; it < iters; it++) {
    int buf = it & 1;
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = sA[buf][t][i * 8 + g];
#pragma unroll
    for (int j = 0; j < 4; j++) b[j] = sB[buf][t][j * 8 + g];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) r += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  int iters = 20000;
  for (int threads : {256, 512, 1024}) {
    for (int bps : {1, 2}) {
      if (threads * bps > 2048) continue;
      int blocks = sms * bps;
      {
        float ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 0.5); }, 5);
        double fl = 2.0 * 8 * iters * (double)threads * blocks;
        printf("{\"bench\": \"dfma_ilp8\", \"threads\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", threads, bps, ms, fl / ms * 1e-9);
      }
      {
        float ms = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 0.5); }, 5);
        double fl = 2.0 * 256 * 8 * iters * (double)(threads / 32) * blocks;
        printf("{\"bench\": \"dmma884_acc8\", \"threads\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", threads, bps, ms, fl / ms * 1e-9);
      }
      {
        float ms = time_ms([&] { dmma_kernel<16><<<blocks, threads>>>(out, iters, 0.5); }, 5);
        double fl = 2.0 * 256 * 16 * iters * (double)(threads / 32) * blocks;
        printf("{\"bench\": \"dmma884_acc16\", \"threads\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", threads, bps, ms, fl / ms * 1e-9);
      }
    }
  }
  for (int warps_per_smsp : {1, 2}) {
    int threads = 128 * warps_per_smsp; if (threads > 256) threads = 256;
    int blocks = sms * (warps_per_smsp == 2 ? 1 : 1);
    float ms = time_ms([&] { dmma_kernel<16><<<blocks, threads>>>(out, iters, 0.5); }, 5);
    double fl = 2.0 * 256 * 16 * iters * (double)(threads / 32) * blocks;
    printf("{\"bench\": \"dmma884_acc16_lowocc\", \"threads\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", threads, ms, fl / ms * 1e-9);
  }
  for (int bps : {1, 2}) {
    int blocks = sms * bps;
    float ms = time_ms([&] { dmma_smem_kernel<<<blocks, 256>>>(out, iters / 4); }, 5);
    double fl = 2.0 * 256 * 32 * (iters / 4) * 8.0 * blocks;
    printf("{\"bench\": \"dmma884_smem_64x32\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9);
  }
  // sustained: ~3 s of back-to-back DMMA
  {
    int blocks = sms * 2, threads = 512;
    float ms1 = time_ms([&] { dmma_kernel<16><<<blocks, threads>>>(out, iters, 0.5); }, 1);
    int reps = (int)(3000.0f / ms1) + 1;
    float ms = time_ms([&] { dmma_kernel<16><<<blocks, threads>>>(out, iters, 0.5); }, reps);
    double fl = 2.0 * 256 * 16 * iters * (double)(threads / 32) * blocks;
    printf("{\"bench\": \"dmma884_sustained_3s\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, fl / ms * 1e-9);
    float msf1 = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 0.5); }, 1);
    reps = (int)(3000.0f / msf1) + 1;
    float msf = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 0.5); }, reps);
    double flf = 2.0 * 8 * iters * (double)threads * blocks;
    printf("{\"bench\": \"dfma_sustained_3s\", \"ms\": %.3f, \"tflops\": %.2f}\n", msf, flf / msf * 1e-9);
  }
  return 0;
}
