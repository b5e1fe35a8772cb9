// dense.cuh -- once-per-step M x M building blocks: Kuu assembly, blocked Cholesky (diagonal-block
// factorisation + inversion here, panel / trailing updates through gemm.cuh), small element-wise kernels.
// Reference counterparts: `cholesky(Symmetric(cov(fz)))` (utils.jl:17 via SVA.jl:131,181) and
// `cholesky(Symmetric(B))` (Laplace.jl:216).
#pragma once
#include "gemm.cuh"
#include "kfun.cuh"

namespace agp {

// zs = s (.) z, zn = |zs|^2 ; rows >= M are zero.  z is point-major [M][D].  zsp (optional) is the padded copy the
// Kuf generator streams: row = [zs_0 .. zs_{D-1}, 0.., zn, 0] with kuf_dp(D) + 2 doubles.
__global__ void prep_z_kernel(const double* z, double* zs, double* zn, double* zsp, int Mp, KernelParams kp) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= Mp) return;
  const int Dp = kuf_dp(kp.D), Sx = Dp + 2;
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
    v = kp.variance * kappa_kp(kp, u);
    if (a == b) v += jitter;
  }
  K[(int64_t)b * Mp + a] = v;
}

// K[a + b*ld] = k(x_a, y_b) for two point sets given as scaled points / squared norms (cov(f.prior, x, y));
// entries with a >= nx or b >= ny are zero.
__global__ void cross_k_kernel(double* K, int64_t ld, const double* xs, const double* xn, int nx, const double* ys, const double* yn, int ny,
                               int rows, KernelParams kp) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (a >= rows) return;
  double v = 0.0;
  if (a < nx && b < ny) {
    double u;
    if (kp.D == 1 && kp.kind != AGP_KERNEL_LINEAR) {
      const double df = xs[a] - ys[b];
      u = df * df;
    } else {
      double dot = 0.0;
      for (int d = 0; d < kp.D; d++) dot = fma(xs[(int64_t)a * kp.D + d], ys[(int64_t)b * kp.D + d], dot);
      u = u_from_dot(kp.kind, xn[a], yn[b], dot);
    }
    v = kp.variance * kappa_kp(kp, u);
  }
  K[(int64_t)b * ld + a] = v;
}

// ---- diagonal-block kernel of the blocked Cholesky -----------------------------------------------------------
// One CTA factorises a 128 x 128 diagonal block (lower triangle of src) and inverts the factor:
//   dstL  <- L (strict upper zeroed),  dstLt <- inv(L),  dstUt <- inv(L)^T        (all column-major, leading dim ld)
// info[0] = first failing global column + 1 (PosDefException).  Everything runs out of shared memory with 32-wide
// blocking so that only a dozen block-wide barriers are needed: per 32-block a warp-level factorisation of the
// diagonal, a register-resident row solve of the panel and a rank-32 trailing update; then the inverse by the
// recursive rule  inv([L11 0; L21 L22]) = [X11 0; -X22 L21 X11  X22].
constexpr int PT_LD = 129;
constexpr int PT_SMEM_BYTES = (128 * PT_LD + 64 * 65 + 128) * 8;

// Cholesky of a 32 x 32 diagonal block by one warp, in place in shared memory (lane i owns row i; left-looking: column J is
// a_iJ - sum_{k<J} l_ik l_Jk, a dot product of the lane's own row with row J, which every lane reads as a broadcast).
// This warp's per-column dependency chain (shuffle of the diagonal -> rsqrt -> corrections -> store) is the critical path of
// the whole diagonal-block kernel: tools/pt_timing.cu measures ~14 us per 32-block (~850 cycles per column) for this loop,
// for a fully unrolled right-looking register version (16 / 10 us cold / warm, ~25 KB of straight-line code) and for a variant
// with the chain cut to rsqrt + one multiply and the k < J part of the next dot product issued under the rsqrt (14.6 us), so
// the compact loop is kept (profiles/r03_pt_timing.txt).
__device__ __forceinline__ void pt_potrf32(double* blk, int lane, double* rdiag, int col0, int* info) {
  double* row = blk + lane * PT_LD;
  for (int J = 0; J < 32; J++) {
    const double* rJ = blk + J * PT_LD;
    double acc0 = row[J], acc1 = 0.0;
    int k = 0;
    for (; k + 1 < J; k += 2) {
      acc0 = fma(-row[k], rJ[k], acc0);
      acc1 = fma(-row[k + 1], rJ[k + 1], acc1);
    }
    if (k < J) acc0 = fma(-row[k], rJ[k], acc0);
    const double v = acc0 + acc1;
    const double d = __shfl_sync(0xffffffffu, v, J);
    if (lane == 0 && !(d > 0.0)) atomicCAS(info, 0, col0 + J + 1);
    // sqrt(d), 1/sqrt(d) and v/sqrt(d) from one rsqrt plus FMA corrections (the results agree with sqrt() and the
    // true quotients to the last bit or one ulp; a dependent sqrt + two divisions cost 4x the latency per column)
    const double y = rsqrt(d);
    const double s0 = d * y;
    const double sq = fma(fma(-s0, s0, d), 0.5 * y, s0);
    const double inv = fma(fma(-sq, y, 1.0), y, y);
    const double l0 = v * inv;
    const double lij = (lane == J) ? sq : fma(fma(-l0, sq, v), inv, l0);
    if (lane >= J) row[J] = lij;  // rows above the diagonal keep the zero they were loaded with
    if (lane == J) rdiag[J] = inv;
    __syncwarp();  // column J is read (as row J's entries) by every later column
  }
}
template <int J>
struct PtSolveRow {  // x L_d^T = a for one row held in x[] (right-looking: x[J] is final, then x[k > J] -= x[J] L[k][J])
  static __device__ __forceinline__ void run(double (&x)[32], const double* Ld, const double* rdiag) {
    x[J] *= rdiag[J];
#pragma unroll
    for (int k = J + 1; k < 32; k++) x[k] = fma(-x[J], Ld[k * PT_LD + J], x[k]);
    PtSolveRow<J + 1>::run(x, Ld, rdiag);
  }
};
template <>
struct PtSolveRow<32> {
  static __device__ __forceinline__ void run(double (&)[32], const double*, const double*) {}
};
template <int I>
struct PtInvCol {  // column `lane` of inv(L_d) in x[] (right-looking forward substitution of the unit vector e_lane)
  static __device__ __forceinline__ void run(double (&x)[32], const double* Ld, const double* rdiag, int lane) {
    x[I] = (I >= lane) ? x[I] * rdiag[I] : 0.0;
#pragma unroll
    for (int k = I + 1; k < 32; k++) x[k] = fma(-x[I], Ld[k * PT_LD + I], x[k]);
    PtInvCol<I + 1>::run(x, Ld, rdiag, lane);
  }
};
template <>
struct PtInvCol<32> {
  static __device__ __forceinline__ void run(double (&)[32], const double*, const double*, int) {}
};

// X[r0.., c0..] (sz x sz) <- -X22 * L21 * X11 with X22 = s[r0.., r0..], L21 = s[r0.., c0..], X11 = s[c0.., c0..]
// (both already inverted, upper parts zero); tb is a [sz][sz+1] scratch.  All 256 threads must call this.
__device__ __forceinline__ void pt_offdiag(double* s, double* tb, int r0, int c0, int sz) {
  constexpr int LD = PT_LD;
  const int q = sz / 4, ldt = sz + 1;
  const int t = threadIdx.x;
  const bool on = t < q * q;
  const int ti = t / q, tj = t % q;  // the thread owns elements (ti + a q, tj + b q): neighbouring lanes, neighbouring columns
  double acc[4][4];
  if (on) {
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 4
    for (int k = 0; k < sz; k++) {  // T = L21 * X11
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; a++) av[a] = s[(r0 + ti + a * q) * LD + c0 + k];
#pragma unroll
      for (int b = 0; b < 4; b++) bv[b] = s[(c0 + k) * LD + c0 + tj + b * q];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) tb[(ti + a * q) * ldt + tj + b * q] = acc[a][b];
  }
  __syncthreads();
  if (on) {
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 4
    for (int k = 0; k < sz; k++) {  // X22 * T  (X22[i][k] = 0 for k > i)
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; a++) av[a] = s[(r0 + ti + a * q) * LD + r0 + k];
#pragma unroll
      for (int b = 0; b < 4; b++) bv[b] = tb[k * ldt + tj + b * q];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) s[(r0 + ti + a * q) * LD + c0 + tj + b * q] = -acc[a][b];  // L21 was consumed before the barrier
  }
  __syncthreads();
}

#ifdef AGP_PT_TIMING
#define PT_TICK(k) if (threadIdx.x == 0) pt_ticks[k] = clock64();
__device__ long long pt_ticks[16];
#else
#define PT_TICK(k)
#endif
// nvalid: leading rows / columns of the block that hold data; the rest must be identity padding, for which the 32-blocks
// of the factorisation and of the inverse are skipped (a 20 x 20 problem costs one 32-block instead of four).
__global__ void __launch_bounds__(256) potrf_trinv128_kernel(const double* src, double* dstL, double* dstLt, double* dstUt, int64_t ld, int col0,
                                                             int* info, int nvalid) {
  extern __shared__ double s[];  // [128][129] + scratch [64][65]
  constexpr int N = 128, LD = PT_LD;
  double* tb = s + N * LD;
  double* rdiag = tb + 64 * 65;  // 1 / L_jj
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < N * N; i += 256) {
    const int r = i % N, c = i / N;
    s[r * LD + c] = (r >= c) ? src[(int64_t)c * ld + r] : 0.0;
  }
  __syncthreads();
  PT_TICK(0)
  // ---- Cholesky, 32-wide blocks -------------------------------------------------------------------------------
  const int nact = min(N, (nvalid + 31) & ~31);  // active leading part (multiple of 32)
  if (tid < N) rdiag[tid] = 1.0;
  __syncthreads();
  for (int o = 0; o < nact; o += 32) {
    if (warp == 0) pt_potrf32(s + o * LD + o, lane, rdiag + o, col0 + o, info);  // diagonal 32 x 32 block
    __syncthreads();
    if (o == 0) { PT_TICK(1) }
    const int T = nact - o - 32;  // rows below the diagonal block (rows >= nact are identity padding: zero below the diagonal)
    if (tid < T) {             // panel: X L_d^T = A, one row per thread, held in registers
      double* row = s + (o + 32 + tid) * LD + o;
      double x[32];
#pragma unroll
      for (int j = 0; j < 32; j++) x[j] = row[j];
      PtSolveRow<0>::run(x, s + o * LD + o, rdiag + o);
#pragma unroll
      for (int j = 0; j < 32; j++) row[j] = x[j];
    }
    __syncthreads();
    if (o == 0) { PT_TICK(2) }
    // trailing update of the lower triangle: A[i][c] -= sum_k P[i][k] P[c][k]; a thread owns the 4 x 4 elements
    // (ti + a q, tc + b q) so that neighbouring lanes read neighbouring rows (conflict-free with ld = 129)
    const int q = T / 4;
    for (int mt = tid; mt < q * q; mt += 256) {
      const int ti = mt / q, tc = mt % q;
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
      const double* pi = s + (o + 32 + ti) * LD + o;
      const double* pc = s + (o + 32 + tc) * LD + o;
#pragma unroll 4
      for (int k = 0; k < 32; k++) {
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; a++) av[a] = pi[a * q * LD + k];
#pragma unroll
        for (int b = 0; b < 4; b++) bv[b] = pc[b * q * LD + k];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
      }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
          if (tc + b * q <= ti + a * q) s[(o + 32 + ti + a * q) * LD + o + 32 + tc + b * q] -= acc[a][b];
    }
    __syncthreads();
    if (o == 0) { PT_TICK(3) }
  }
  PT_TICK(4)
  for (int i = tid; i < N * N; i += 256) {
    const int r = i % N, c = i / N;
    dstL[(int64_t)c * ld + r] = s[r * LD + c];
  }
  __syncthreads();
  PT_TICK(5)
  // ---- inverse: the four 32 x 32 diagonal blocks (one warp each; lane j builds column j in registers) ------------
  if (warp < 4 && warp * 32 < nact) {
    const int o = warp * 32;
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; i++) x[i] = (i == lane) ? 1.0 : 0.0;
    PtInvCol<0>::run(x, s + o * LD + o, rdiag + o, lane);
    __syncwarp();  // every lane has finished reading the L rows of this block
#pragma unroll
    for (int i = 0; i < 32; i++) s[(o + i) * LD + o + lane] = x[i];
  }
  __syncthreads();
  PT_TICK(6)
  if (nact > 32) pt_offdiag(s, tb, 32, 0, 32);
  if (nact > 96) pt_offdiag(s, tb, 96, 64, 32);
  PT_TICK(7)
  if (nact > 64) pt_offdiag(s, tb, 64, 0, 64);
  PT_TICK(8)
  for (int i = tid; i < N * N; i += 256) {
    const int r = i % N, c = i / N;
    dstLt[(int64_t)c * ld + r] = s[r * LD + c];
    dstUt[(int64_t)c * ld + r] = s[c * LD + r];
  }
  PT_TICK(9)
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
