// dense.cuh -- once-per-step M x M building blocks: Kuu assembly, blocked Cholesky (diagonal-block
// factorisation + inversion here, panel / trailing updates through gemm.cuh), small element-wise kernels.
// Reference counterparts: `cholesky(Symmetric(cov(fz)))` (utils.jl:17 via SVA.jl:131,181) and
// `cholesky(Symmetric(B))` (Laplace.jl:216).
#pragma once
#include "gemm.cuh"
#include "kfun.cuh"

namespace agp {

// zs = s (.) z, zn = |zs|^2 ; rows >= M are zero.  z is point-major [M][D].  zsp (optional) is the padded copy the
// Kuf generator streams: row = [zs_0 .. zs_{D-1}, 0.., zn, 0] with Dp + 2 doubles, Dp = D rounded up to 2.
__global__ void prep_z_kernel(const double* z, double* zs, double* zn, double* zsp, int Mp, KernelParams kp) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= Mp) return;
  const int Dp = (kp.D + 1) & ~1, Sx = Dp + 2;
  double s = 0.0;
  for (int d = 0; d < kp.D; d++) {
    const double v = (row < kp.M) ? z[(int64_t)row * kp.D + d] * kp.s[d] : 0.0;
    zs[(int64_t)row * kp.D + d] = v;
    if (zsp) zsp[(int64_t)row * Sx + d] = v;
    s = fma(v, v, s);
  }
  zn[row] = s;
  if (zsp) {
    for (int d = kp.D; d < Sx; d++) zsp[(int64_t)row * Sx + d] = 0.0;
    zsp[(int64_t)row * Sx + Dp] = s;
  }
}

// Kuu = k(Z, Z) + jitter I on the leading M x M block, identity on the padding.  Column-major, ld = Mp.
__global__ void build_kuu_kernel(double* K, int Mp, const double* zs, const double* zn, double jitter, KernelParams kp) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;  // row
  const int b = blockIdx.y;                             // column
  if (a >= Mp) return;
  double v;
  if (a >= kp.M || b >= kp.M) {
    v = (a == b) ? 1.0 : 0.0;
  } else {
    double u;
    if (a == b && kp.kind != AGP_KERNEL_LINEAR) {
      u = 0.0;  // pairwise(d, x) has an exactly-zero diagonal
    } else if (kp.D == 1 && kp.kind != AGP_KERNEL_LINEAR) {
      const double df = zs[a] - zs[b];
      u = df * df;
    } else {
      double dot = 0.0;
      const int lo = min(a, b), hi = max(a, b);  // same operation order for (a,b) and (b,a): exactly symmetric
      for (int d = 0; d < kp.D; d++) dot = fma(zs[(int64_t)lo * kp.D + d], zs[(int64_t)hi * kp.D + d], dot);
      u = u_from_dot(kp.kind, zn[lo], zn[hi], dot);
    }
    v = kp.variance * kappa(kp.kind, u, kp.c);
    if (a == b) v += jitter;
  }
  K[(int64_t)b * Mp + a] = v;
}

// In-place Cholesky of one 128 x 128 diagonal block (column-major, leading dimension ld): reads the lower
// triangle of src, writes L (strict upper zeroed) to dst.  info[0] = first failing global column + 1.
__global__ void __launch_bounds__(256) potrf128_kernel(const double* src, double* dst, int64_t ld, int col0, int* info) {
  extern __shared__ double s[];  // [128][129]
  constexpr int N = 128, LD = 129;
  const int tid = threadIdx.x;
  for (int i = tid; i < N * N; i += 256) {
    const int r = i % N, c = i / N;
    s[r * LD + c] = (r >= c) ? src[(int64_t)c * ld + r] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < N; j++) {
    if (tid == 0) {
      const double d = s[j * LD + j];
      if (!(d > 0.0)) atomicCAS(info, 0, col0 + j + 1);
      s[j * LD + j] = sqrt(d);
    }
    __syncthreads();
    const double dj = s[j * LD + j];
    if (tid > j && tid < N) s[tid * LD + j] /= dj;
    __syncthreads();
    // trailing update of the lower triangle: rows i > j, columns j < k <= i
    const int i = j + 1 + (tid & 127);
    if (i < N) {
      const double lij = s[i * LD + j];
      const int half = tid >> 7;  // two threads per row split the columns
      for (int k = j + 1 + half; k <= i; k += 2) s[i * LD + k] = fma(-lij, s[k * LD + j], s[i * LD + k]);
    }
    __syncthreads();
  }
  for (int i = tid; i < N * N; i += 256) {
    const int r = i % N, c = i / N;
    dst[(int64_t)c * ld + r] = s[r * LD + c];
  }
}

// Inverse of the 128 x 128 lower-triangular block L (column-major, ld): inv -> dstL (column-major),
// inv^T -> dstU (column-major).  One thread per column of the inverse, in place in shared memory: row i of
// L is last read at step i, so X(i, :) can overwrite it (one barrier between the reads and the write).
__global__ void __launch_bounds__(128) trinv128_kernel(const double* L, int64_t ld, double* dstL, double* dstU, int64_t ldd) {
  extern __shared__ double s[];  // [128][129]
  constexpr int N = 128, LD = 129;
  const int j = threadIdx.x;
  for (int i = j; i < N * N; i += 128) {
    const int r = i % N, c = i / N;
    s[r * LD + c] = (r >= c) ? L[(int64_t)c * ld + r] : 0.0;
  }
  __syncthreads();
  for (int i = 0; i < N; i++) {
    double acc = (i == j) ? 1.0 : 0.0;
    for (int k = 0; k < i; k++) acc = fma(-s[i * LD + k], s[k * LD + j], acc);
    const double v = (i >= j) ? acc / s[i * LD + i] : 0.0;
    __syncthreads();
    s[i * LD + j] = v;
  }
  __syncthreads();
  for (int i = j; i < N * N; i += 128) {
    const int r = i % N, c = i / N;
    dstL[(int64_t)c * ldd + r] = s[r * LD + c];
    dstU[(int64_t)c * ldd + r] = s[c * LD + r];
  }
}

// out = in^T for square n x n matrices with leading dimension ld (out != in)
__global__ void transpose_kernel(const double* in, double* out, int n, int64_t ld) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) tile[j][threadIdx.x] = in[(int64_t)(by + j) * ld + bx + threadIdx.x];
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) out[(int64_t)(bx + j) * ld + by + threadIdx.x] = tile[threadIdx.x][j];
}

__global__ void fill_kernel(double* p, int64_t n, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// out[i] = sum_s in[s*stride + i]   (fixed order)
__global__ void sum_slices_kernel(const double* in, int nslices, int64_t stride, int64_t n, double* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < nslices; k++) s += in[(int64_t)k * stride + i];
    out[i] = s;
  }
}

// G (column-major, lower triangle valid) -> full symmetric
__global__ void symmetrize_from_lower_kernel(double* G, int n, int64_t ld) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (r < n && r > c) G[(int64_t)r * ld + c] = G[(int64_t)c * ld + r];
}

// S <- 0.5 (S + S^T)   (two-array form: out = 0.5 (A + At) where At is the transposed copy)
__global__ void sym_average_kernel(const double* A, const double* At, double* out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = 0.5 * (A[i] + At[i]);
}

}  // namespace agp
