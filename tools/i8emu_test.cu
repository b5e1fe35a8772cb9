// Standalone check of the INT8-slice FP64 emulation engine (csrc/i8emu.cuh) against a host reference in long double, and its throughput at
// the sweep's shape.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/i8emu_test tools/i8emu_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../approximategps.jl_b200/csrc/i8emu.cuh"
using namespace agp;
using namespace agp::i8e;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct EpiStore {
  double* D;
  int ld;
  const double* sA;
  const double* sB;
  __device__ void operator()(int tm, int tn, int z, int row, int c0, const double (&v)[32]) const {
    const int m = tm * EM + row, n0 = tn * EN + c0;
    const double sa = sA[m] * (1.0 / 16384.0);
    double* d = D + (size_t)m * ld + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 2) *reinterpret_cast<double2*>(d + j) = make_double2(v[j] * sa * sB[n0 + j], v[j + 1] * sa * sB[n0 + j + 1]);
  }
};

struct Sliced { signed char* planes; double* scale; int64_t plane_bytes; };
static Sliced upload_slice(const std::vector<double>& x, int rows, int K) {
  double* d; Sliced s;
  s.plane_bytes = (int64_t)rows * K;
  CK(cudaMalloc(&d, x.size() * 8)); CK(cudaMalloc(&s.planes, (size_t)S * s.plane_bytes)); CK(cudaMalloc(&s.scale, rows * 8));
  CK(cudaMemcpy(d, x.data(), x.size() * 8, cudaMemcpyHostToDevice));
  slice_rows_kernel<S><<<(rows + 7) / 8, 256>>>(d, K, rows, K, s.planes, K, s.plane_bytes, s.scale);
  CK(cudaDeviceSynchronize()); CK(cudaFree(d));
  return s;
}

static double run_case(const char* name, int M, int N, int K, int kmode, bool timing, double spread) {
  std::vector<double> A((size_t)M * K), B((size_t)N * K);
  uint64_t st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return ((st >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 2.0; };
  // entries with a wide dynamic range inside a row (spread decades) so that the slicing is exercised on small entries next to large ones
  for (auto& v : A) { const double r = rnd(); v = r * std::pow(10.0, -spread * std::fabs(rnd())); }
  for (auto& v : B) { const double r = rnd(); v = r * std::pow(10.0, -spread * std::fabs(rnd())); }
  if (kmode == KM_FROM_N) for (int n = 0; n < N; n++) for (int k = 0; k < n && k < K; k++) B[(size_t)n * K + k] = 0.0;   // lower triangular [n][k], k >= n
  if (kmode == KM_UPTO_N) for (int n = 0; n < N; n++) for (int k = n + 1; k < K; k++) B[(size_t)n * K + k] = 0.0;        // k <= n
  Sliced a = upload_slice(A, M, K), b = upload_slice(B, N, K);
  CUtensorMap ma, mb;
  if (!make_map3(&ma, a.planes, K, M, K, a.plane_bytes, EM) || !make_map3(&mb, b.planes, K, N, K, b.plane_bytes, EN)) { printf("%s: tensor map creation failed\n", name); return -1; }
  double* D; CK(cudaMalloc(&D, (size_t)M * N * 8)); CK(cudaMemset(D, 0xff, (size_t)M * N * 8));
  Args g{K, kmode, 0, 0};
  EpiStore epi{D, N, a.scale, b.scale};
  auto kern = i8emu_gemm_kernel<EpiStore>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  dim3 grid(N / EN, M / EM, 1);
  kern<<<grid, E_THREADS, SMEM_BYTES>>>(ma, mb, g, epi);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<double> Dh((size_t)M * N);
  CK(cudaMemcpy(Dh.data(), D, Dh.size() * 8, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0, maxrel_bound = 0, dgemm_err = 0;
  for (int m = 0; m < M; m += (timing ? 1237 : 1))
    for (int n = 0; n < N; n++) {
      long double s = 0, sabs = 0;
      double sd = 0;
      for (int k = 0; k < K; k++) {
        const double p = A[(size_t)m * K + k], q = B[(size_t)n * K + k];
        s += (long double)p * q; sabs += fabsl((long double)p * q); sd = std::fma(p, q, sd);
      }
      const double err = (double)fabsl(s - (long double)Dh[(size_t)m * N + n]);
      maxerr = std::max(maxerr, err);
      maxref = std::max(maxref, (double)fabsl(s));
      if (sabs > 0) maxrel_bound = std::max(maxrel_bound, err / (double)sabs);
      dgemm_err = std::max(dgemm_err, (double)fabsl(s - (long double)sd));
    }
  printf("{\"case\": \"%s\", \"M\": %d, \"N\": %d, \"K\": %d, \"kmode\": %d, \"decades_within_row\": %.0f, \"max_abs_err\": %.3e, \"max_ref\": %.3e, \"rel_to_max\": %.3e, "
         "\"max_err_over_sum_abs_terms\": %.3e, \"fp64_fma_chain_max_abs_err\": %.3e", name, M, N, K, kmode, spread, maxerr, maxref, maxerr / maxref, maxrel_bound, dgemm_err);
  if (timing) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) kern<<<grid, E_THREADS, SMEM_BYTES>>>(ma, mb, g, epi);
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int i = 0; i < reps; i++) kern<<<grid, E_THREADS, SMEM_BYTES>>>(ma, mb, g, epi);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    const double frac = (kmode == KM_FULL) ? 1.0 : 0.5 * (1.0 + (double)EK / K);  // executed share of the k-range (triangular operand)
    printf(", \"ms\": %.4f, \"fp64_equiv_tflops_full_count\": %.1f, \"int8_issue_tops\": %.1f", ms, 2.0 * M * N * (double)K / ms * 1e-9, 28 * frac * 2.0 * M * N * (double)K / ms * 1e-9);
  }
  printf("}\n");
  fflush(stdout);
  cudaFree(D); cudaFree(a.planes); cudaFree(a.scale); cudaFree(b.planes); cudaFree(b.scale);
  return maxerr / maxref;
}

// <4 slices, 128 columns>: the Float32 mode's reverse-pass solve.  The A operand is given inducing-major [K][M] (as the sweep stores it) and sliced
// by transpose_slice_kernel; B lower triangular [n][k], k >= n.
struct EpiStore4 {
  double* D;
  int ld;
  const double* sA;
  const double* sB;
  __device__ void operator()(int tm, int tn, int z, int row, int c0, const double (&v)[32]) const {
    const int m = tm * EM + row, n0 = tn * 128 + c0;
    const double sa = sA[m] * (1.0 / 16384.0);
    double* d = D + (size_t)m * ld + n0;
#pragma unroll
    for (int j = 0; j < 32; j++) d[j] = v[j] * sa * sB[n0 + j];
  }
};
static double run_case4(const char* name, int M, int N, int K, bool timing) {
  constexpr int NS = 4, N4 = 128;
  std::vector<double> At((size_t)K * M), B((size_t)N * K);  // At[k][m]
  uint64_t st = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return ((st >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 2.0; };
  for (auto& v : At) v = rnd();
  for (auto& v : B) v = rnd();
  for (int n = 0; n < N; n++) for (int k = 0; k < n && k < K; k++) B[(size_t)n * K + k] = 0.0;
  double *dAt, *dB, *sA, *sB, *D; signed char *qa, *qb;
  CK(cudaMalloc(&dAt, At.size() * 8)); CK(cudaMalloc(&dB, B.size() * 8)); CK(cudaMalloc(&sA, M * 8)); CK(cudaMalloc(&sB, N * 8));
  CK(cudaMalloc(&qa, (size_t)NS * M * K)); CK(cudaMalloc(&qb, (size_t)NS * N * K)); CK(cudaMalloc(&D, (size_t)M * N * 8));
  CK(cudaMemcpy(dAt, At.data(), At.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
  transpose_slice_kernel<NS><<<(M + 31) / 32, 256>>>(dAt, M, K, M, qa, K, (int64_t)M * K, sA);
  slice_rows_kernel<NS><<<(N + 7) / 8, 256>>>(dB, K, N, K, qb, K, (int64_t)N * K, sB);
  CK(cudaDeviceSynchronize());
  CUtensorMap ma, mb;
  if (!make_map3(&ma, qa, K, M, K, (uint64_t)M * K, EM, NS) || !make_map3(&mb, qb, K, N, K, (uint64_t)N * K, N4, NS)) { printf("%s: tensor map creation failed\n", name); return -1; }
  Args g{K, KM_FROM_N, 0, 0};
  EpiStore4 epi{D, N, sA, sB};
  auto kern = i8emu_gemm_kernel<EpiStore4, NS, N4>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<NS, N4>::smem_bytes));
  dim3 grid(N / N4, M / EM, 1);
  kern<<<grid, E_THREADS, Cfg<NS, N4>::smem_bytes>>>(ma, mb, g, epi);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  std::vector<double> Dh((size_t)M * N);
  CK(cudaMemcpy(Dh.data(), D, Dh.size() * 8, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; m += (timing ? 1237 : 1))
    for (int n = 0; n < N; n++) {
      long double s = 0;
      for (int k = 0; k < K; k++) s += (long double)At[(size_t)k * M + m] * B[(size_t)n * K + k];
      maxerr = std::max(maxerr, (double)fabsl(s - (long double)Dh[(size_t)m * N + n])); maxref = std::max(maxref, (double)fabsl(s));
    }
  printf("{\"case\": \"%s\", \"slices\": 4, \"M\": %d, \"N\": %d, \"K\": %d, \"max_abs_err\": %.3e, \"max_ref\": %.3e, \"rel_to_max_4slices\": %.3e", name, M, N, K, maxerr, maxref, maxerr / maxref);
  if (timing) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) kern<<<grid, E_THREADS, Cfg<NS, N4>::smem_bytes>>>(ma, mb, g, epi);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 10; i++) kern<<<grid, E_THREADS, Cfg<NS, N4>::smem_bytes>>>(ma, mb, g, epi);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 10;
    cudaEvent_t f0, f1; CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
    CK(cudaEventRecord(f0));
    for (int i = 0; i < 10; i++) transpose_slice_kernel<NS><<<(M + 31) / 32, 256>>>(dAt, M, K, M, qa, K, (int64_t)M * K, sA);
    CK(cudaEventRecord(f1)); CK(cudaEventSynchronize(f1));
    float ms2; CK(cudaEventElapsedTime(&ms2, f0, f1)); ms2 /= 10;
    printf(", \"ms\": %.4f, \"transpose_slice_ms\": %.4f", ms, ms2);
  }
  printf("}\n"); fflush(stdout);
  cudaFree(dAt); cudaFree(dB); cudaFree(sA); cudaFree(sB); cudaFree(qa); cudaFree(qb); cudaFree(D);
  return maxerr / maxref;
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  double worst = 0;
  worst = std::max(worst, run_case("one tile, one k-block", 128, 64, 128, KM_FULL, false, 0));
  worst = std::max(worst, run_case("K-major x K-major", 256, 256, 512, KM_FULL, false, 0));
  worst = std::max(worst, run_case("wide dynamic range", 256, 256, 512, KM_FULL, false, 6));
  worst = std::max(worst, run_case("k >= n (lower triangular B)", 256, 512, 512, KM_FROM_N, false, 0));
  worst = std::max(worst, run_case("k <= n", 256, 512, 512, KM_UPTO_N, false, 0));
  const double w4 = run_case4("4 slices x 128 columns, transposed A, k >= n", 256, 512, 512, false);
  if (!(w4 >= 0 && w4 < 1e-6)) worst = 1.0;
  if (argc > 1) {
    run_case4("sweep shape, 4 slices (Float32-mode S5)", 151552, 1024, 1024, true);
    run_case("sweep shape", 151552, 1024, 1024, KM_FULL, true, 0);
    run_case("sweep shape, triangular B (S2)", 151552, 1024, 1024, KM_FROM_N, true, 0);
  }
  printf("{\"worst_rel_to_max\": %.3e, \"ok\": %s}\n", worst, (worst >= 0 && worst < 2e-13) ? "true" : "false");
  return 0;
}
