// Standalone check of the tcgen05 3xTF32 GEMM engine (csrc/tf32x3.cuh) against an FP64 host reference, for the operand layouts
// and k-ranges the Float32 sweep uses.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tf32x3_test tools/tf32x3_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../approximategps.jl_b200/csrc/tf32x3.cuh"
using namespace agp;
using namespace agp::t5;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void split_kernel(const double* in, float* hi, float* lo, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) split_tf32(in[i], hi[i], lo[i]);
}
struct EpiStoreF32 {
  double* D;
  int ld;
  struct State {};
  __device__ void begin(State&, int, int, int, int, int) const {}
  __device__ void end(State&, int, int, int, int, int) const {}
  __device__ void operator()(State&, int tm, int tn, int z, int row, int c0, const double (&v)[32]) const {
    double* d = D + (size_t)(tm * TM + row) * ld + tn * TN + c0;
#pragma unroll
    for (int j = 0; j < 32; j += 2) *reinterpret_cast<double2*>(d + j) = make_double2(v[j], v[j + 1]);
  }
};

struct Planes { float *h, *l; };
static Planes upload_split(const std::vector<double>& x) {
  double* d; Planes p;
  CK(cudaMalloc(&d, x.size() * 8)); CK(cudaMalloc(&p.h, x.size() * 4)); CK(cudaMalloc(&p.l, x.size() * 4));
  CK(cudaMemcpy(d, x.data(), x.size() * 8, cudaMemcpyHostToDevice));
  split_kernel<<<256, 256>>>(d, p.h, p.l, (int64_t)x.size());
  CK(cudaDeviceSynchronize()); CK(cudaFree(d));
  return p;
}

template <bool AMN, bool BMN>
static double run_case(const char* name, int M, int N, int K, int kmode, bool timing, MnDesc mn = MnDesc(), bool atom32 = true) {
  // logical A[m][k], B[n][k]; storage: K-major -> [rows][K]; MN-major -> [K][rows]
  std::vector<double> A((size_t)M * K), B((size_t)N * K);
  uint64_t st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return ((st >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 2.0; };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  std::vector<double> As(A.size()), Bs(B.size());
  for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) As[AMN ? (size_t)k * M + m : (size_t)m * K + k] = A[(size_t)m * K + k];
  for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) Bs[BMN ? (size_t)k * N + n : (size_t)n * K + k] = B[(size_t)n * K + k];
  Planes pa = upload_split(As), pb = upload_split(Bs);
  CUtensorMap mah, mal, mbh, mbl;
  bool ok = true;
  if (AMN) ok &= make_map(&mah, pa.h, M, K, M, 32, atom32) && make_map(&mal, pa.l, M, K, M, 32, atom32);
  else ok &= make_map(&mah, pa.h, K, M, K, 128) && make_map(&mal, pa.l, K, M, K, 128);
  if (BMN) ok &= make_map(&mbh, pb.h, N, K, N, 32, atom32) && make_map(&mbl, pb.l, N, K, N, 32, atom32);
  else ok &= make_map(&mbh, pb.h, K, N, K, 128) && make_map(&mbl, pb.l, K, N, K, 128);
  if (!ok) { printf("%s: tensor map creation failed\n", name); return -1; }
  double* D; CK(cudaMalloc(&D, (size_t)M * N * 8)); CK(cudaMemset(D, 0, (size_t)M * N * 8));
  Args g{K, kmode, 0, 0, mn};
  EpiStoreF32 epi{D, N};
  auto kern = tf32x3_gemm_kernel<AMN, BMN, EpiStoreF32>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  dim3 grid(M / TM, N / TN, 1);
  kern<<<grid, T5_THREADS, SMEM_BYTES>>>(mah, mal, mbh, mbl, g, epi);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<double> Dh((size_t)M * N);
  CK(cudaMemcpy(Dh.data(), D, Dh.size() * 8, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0, bias = 0, cnt = 0;
  for (int m = 0; m < M; m += (timing ? 1237 : 1))  // (the large timing case is checked on a sample of rows)
    for (int n = 0; n < N; n++) {
      int kb = 0, ke = K;
      if (kmode == KM_FROM_N) kb = (n / TN) * TN;
      if (kmode == KM_UPTO_N) ke = std::min(K, (n / TN + 1) * TN);
      double s = 0;
      for (int k = kb; k < ke; k++) s += A[(size_t)m * K + k] * B[(size_t)n * K + k];
      maxerr = std::max(maxerr, std::fabs(s - (double)Dh[(size_t)m * N + n]));
      bias += (std::fabs((double)Dh[(size_t)m * N + n]) - std::fabs(s)); cnt += std::fabs(s);
      maxref = std::max(maxref, std::fabs(s));
    }
  printf("{\"case\": \"%s\", \"M\": %d, \"N\": %d, \"K\": %d, \"kmode\": %d, \"max_abs_err\": %.3e, \"max_ref\": %.3e, \"rel\": %.3e, \"rel_bias_of_magnitude\": %.3e", name, M, N, K, kmode, maxerr, maxref, maxerr / maxref, bias / cnt);
  if (timing) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) kern<<<grid, T5_THREADS, SMEM_BYTES>>>(mah, mal, mbh, mbl, g, epi);
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int i = 0; i < reps; i++) kern<<<grid, T5_THREADS, SMEM_BYTES>>>(mah, mal, mbh, mbl, g, epi);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    printf(", \"ms\": %.4f, \"fp32_equiv_tflops\": %.1f, \"tf32_issue_tflops\": %.1f", ms, 2.0 * M * N * (double)K / ms * 1e-9, 6.0 * M * N * (double)K / ms * 1e-9);
  }
  printf("}\n");
  cudaFree(D); cudaFree(pa.h); cudaFree(pa.l); cudaFree(pb.h); cudaFree(pb.l);
  return maxerr / maxref;
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  double worst = 0;
  worst = std::max(worst, run_case<false, false>("K-major x K-major", 256, 256, 512, KM_FULL, false));
  worst = std::max(worst, run_case<false, false>("K-major x K-major, k >= n-tile", 256, 512, 512, KM_FROM_N, false));
  worst = std::max(worst, run_case<false, false>("K-major x K-major, k <= n-tile", 256, 512, 512, KM_UPTO_N, false));
  double mn = run_case<false, true>("K-major x MN-major (32B atom, SBO 512)", 256, 256, 512, KM_FULL, false);
  if (!(mn < 1e-5)) {  // probe the alternatives so that one GPU run tells which descriptor the hardware wants
    MnDesc d;
    d.sbo16 = 1024 >> 4; run_case<false, true>("probe: 32B atom, SBO 1024", 256, 256, 512, KM_FULL, false, d);
    d.sbo16 = 256 >> 4;  run_case<false, true>("probe: 32B atom, SBO 256", 256, 256, 512, KM_FULL, false, d);
    d = MnDesc(); d.lbo16 = 512 >> 4; d.sbo16 = 4096 >> 4; run_case<false, true>("probe: 32B atom, LBO/SBO swapped", 256, 256, 512, KM_FULL, false, d);
    d = MnDesc(); d.layout = 2; d.sbo16 = 1024 >> 4; run_case<false, true>("probe: 16B atom layout, TMA 32B atom", 256, 256, 512, KM_FULL, false, d);
    d = MnDesc(); d.layout = 2; d.sbo16 = 1024 >> 4; run_case<false, true>("probe: 16B atom layout, TMA 128B", 256, 256, 512, KM_FULL, false, d, false);
    d = MnDesc(); d.layout = 2; d.lbo16 = 1024 >> 4; d.sbo16 = 4096 >> 4; run_case<false, true>("probe: 16B atom, swapped, TMA 128B", 256, 256, 512, KM_FULL, false, d, false);
  }
  worst = std::max(worst, mn);
  worst = std::max(worst, run_case<true, true>("MN-major x MN-major", 256, 256, 512, KM_FULL, false));
  worst = std::max(worst, run_case<true, false>("MN-major x K-major", 256, 256, 512, KM_FULL, false));
  if (argc > 1) {
    // the sweep shape: 151 552 points x 1024 x 1024
    run_case<false, false>("sweep shape K x K", 151552, 1024, 1024, KM_FULL, true);
  }
  printf("{\"worst_rel\": %.3e, \"ok\": %s}\n", worst, (worst >= 0 && worst < 1e-5) ? "true" : "false");
  return (worst >= 0 && worst < 1e-5) ? 0 : 1;
}
