// small.cuh -- the latency path: one `elbo` + all gradients of a SMALL problem in ONE kernel launch of ONE CTA.
//
// BASELINE.json configs[0] (examples/a-regression/script.jl:170-194) evaluates the ELBO on minibatches of 100 points with
// M = 20 inducing points, 30 000 times.  The throughput path of sweep.cuh needs ~50 launches and ~15 small copies for one such
// step (0.5 ms: slower than OpenBLAS on the host); here the whole step -- Kuu, its Cholesky and inverse, the Kuf tile, the
// marginals, the expected log-likelihood, the reverse pass and the O(M^3) epilogue -- runs as one CTA of 1024 threads over an
// L1-resident workspace, reads its parameters straight from the caller's (mapped, pinned) flat vector and writes ELBO and the
// gradient back into mapped pinned memory: one launch + one stream synchronisation per optimiser step.
//
// Same mathematics, in the same whitened variables, as the throughput path (agp_svgp_finish), i.e. the
// reference's posterior(sva) / mean_and_var / expected_loglikelihood / _prior_kl (SVA.jl:115-187, :246-253, :340-373) and their
// Zygote pullbacks.  The one structural difference: with M <= 128 the triangular solves go through the explicit inverse of the
// Cholesky factor (a column per thread, no barriers), so that everything else is a small dense product parallel over its outputs.
#pragma once
#include "common.cuh"
#include "kfun.cuh"

namespace agp {

constexpr int SM_THREADS = 512;  // 128 registers per thread: the ~35 array pointers of the workspace stay in registers
constexpr int SM_MAXM = 128;
constexpr int SM_TILE = 256;  // points per pass

struct SmallArgs {
  const double* flat;  // [variance | inv_lengthscale (n_scale) | linear_c | mean_const | lik parameter | Z (M*D) | m (M) | Lq (M*M col-major)]
  double* out;         // [elbo | gradient in the same flat layout | status, info]  (mapped pinned host memory)
  const double* X;     // device: points [count][D] (already offset)
  const double* y;     // device: [count]
  int count;
  double scale;        // num_data / count
  int M, D, n_scale, kind, centered;
  double jitter;
  LikParams lp;        // kind, method, ngh, seed, gh table; sigma2 is read from flat
  long long point_base;
  double* ws;          // device workspace (small_ws_doubles)
  int want_grad;
  int ldb;             // min(count, SM_TILE) rounded up to 32
  int64_t smem_doubles; // svgp_small_kernel<false>: doubles of dynamic shared memory handed to the first arrays of the carve-up
};

__host__ __device__ inline int64_t small_ws_doubles(int M, int D) {
  const int64_t MM = (int64_t)M * M, MB = (int64_t)M * SM_TILE;
  return (int64_t)M * D * 5 + 9 * (int64_t)M + 11 * MM + (int64_t)SM_TILE * (D + 8) + 6 * MB + 64 + 2 * MAXD + 128;  // (+ rounding of each array to 2)
}

// exact footprint for a given tile row length (the kernel's `take` sequence)
__host__ __device__ inline int64_t small_doubles_exact(int M, int D, int ldb) {
  auto r2 = [](int64_t n) { return (n + 1) & ~(int64_t)1; };
  const int64_t MM = (int64_t)M * M, MB = (int64_t)M * ldb, MD = (int64_t)M * D;
  return 5 * r2(MD) + 9 * r2(M) + 11 * r2(MM) + 6 * r2(MB) + r2((int64_t)ldb * D) + 3 * r2(ldb) + r2(64 + 2 * MAXD);
}

// deterministic block reduction of one value per thread (all threads must call); result broadcast to every thread
__device__ __forceinline__ double sm_block_sum(double v, double* sred /* >= 33 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sred[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double r = (lane < (int)(blockDim.x >> 5)) ? sred[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0) sred[32] = r;
  }
  __syncthreads();
  return sred[32];
}

// Small dense product on the FP64 tensor path (mma.sync.m8n8k4), all warps of the CTA.  The one-output-per-thread dot products this
// replaces are bound by shared-memory instruction issue (two LDS per FMA: 10-14 us per 50 x 50 x 100 product on one SM); a DMMA needs
// two LDS per 256 FMAs, which leaves the SM's FP64 rate as the bound (2 us for that product).
//   C(i, j) = sum_k A(i, k) B(k, j),   A(i, k) = A[i * sai + k * sak],   B(k, j) = B[k * sbk + j * sbj],   out(i, j, value) for i < R, j < C
// Output blocks of 16 x 16 (2 x 2 DMMA tiles sharing their fragments) are dealt to the warps round-robin; k runs in ascending quads.
// Rows / columns of a padded block are clamped to the last valid one (their results are dropped); the last k-quad is masked.
// Structural zeros of triangular operands shorten the k-range of a block (the operands also hold those zeros explicitly, so the sums
// are the same as the full ones):  ta / tb = SM_TRI_LE: A(i, k) [B(k, j)] = 0 for k > i [k > j];  SM_TRI_GE: = 0 for k < i [k < j].
// lower_out: blocks strictly above the diagonal are skipped.
constexpr int SM_TRI_NONE = 0, SM_TRI_LE = 1, SM_TRI_GE = 2;
template <class FO>
__device__ __forceinline__ void sm_dmma_gemm(int R, int C, int K, const double* A, int sai, int sak, int ta, const double* B, int sbk, int sbj, int tb,
                                             bool lower_out, FO out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5, g = lane >> 2, t = lane & 3;
  const int tr = (R + 15) >> 4, tc = (C + 15) >> 4;
  for (int p = warp; p < tr * tc; p += nw) {
    const int i0 = (p % tr) * 16, j0 = (p / tr) * 16;
    if (lower_out && j0 > i0 + 15) continue;
    int kb = 0, ke = K;
    if (ta == SM_TRI_LE) ke = min(ke, i0 + 16);
    if (tb == SM_TRI_LE) ke = min(ke, j0 + 16);
    if (ta == SM_TRI_GE) kb = max(kb, i0);
    if (tb == SM_TRI_GE) kb = max(kb, j0);
    kb &= ~3;
    const int ia = min(i0 + g, R - 1), ib = min(i0 + 8 + g, R - 1), ja = min(j0 + g, C - 1), jb = min(j0 + 8 + g, C - 1);
    double c00[2] = {0.0, 0.0}, c01[2] = {0.0, 0.0}, c10[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
    const double* pa0 = A + ia * sai + (kb + t) * sak;
    const double* pa1 = A + ib * sai + (kb + t) * sak;
    const double* pb0 = B + (kb + t) * sbk + ja * sbj;
    const double* pb1 = B + (kb + t) * sbk + jb * sbj;
    const int kfull = kb + ((ke - kb) & ~3);
    for (int k0 = kb; k0 < kfull; k0 += 4) {
      const double a0 = *pa0, a1 = *pa1, b0 = *pb0, b1 = *pb1;
      dmma884(c00, a0, b0);
      dmma884(c01, a0, b1);
      dmma884(c10, a1, b0);
      dmma884(c11, a1, b1);
      pa0 += 4 * sak;
      pa1 += 4 * sak;
      pb0 += 4 * sbk;
      pb1 += 4 * sbk;
    }
    if (kfull < ke) {  // last, partial quad: clamp the address, mask the value
      const bool kv = kfull + t < ke;
      const int back = kv ? 0 : (kfull + t - (ke - 1));
      const double a0 = kv ? *(pa0 - back * sak) : 0.0, a1 = kv ? *(pa1 - back * sak) : 0.0;
      const double b0 = kv ? *(pb0 - back * sbk) : 0.0, b1 = kv ? *(pb1 - back * sbk) : 0.0;
      dmma884(c00, a0, b0);
      dmma884(c01, a0, b1);
      dmma884(c10, a1, b0);
      dmma884(c11, a1, b1);
    }
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int ca = j0 + 2 * t + e, cb = ca + 8, ra = i0 + g, rb = ra + 8;
      if (ra < R && ca < C) out(ra, ca, c00[e]);
      if (ra < R && cb < C) out(ra, cb, c01[e]);
      if (rb < R && ca < C) out(rb, ca, c10[e]);
      if (rb < R && cb < C) out(rb, cb, c11[e]);
    }
  }
}

// development aid (-DAGP_SMALL_TIMING, tools/c1_phases.py): %globaltimer at the phase boundaries, returned behind the status words
#ifdef AGP_SMALL_TIMING
#define SM_TICK(k)                                                                    \
  if (threadIdx.x == 0) {                                                             \
    unsigned long long t_;                                                            \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                            \
    a.out[3 + (4 + a.n_scale + (int64_t)a.M * a.D + a.M + (int64_t)a.M * a.M) + (k)] = (double)t_; \
  }
#else
#define SM_TICK(k)
#endif

template <bool ALLSMEM>
__global__ void __launch_bounds__(SM_THREADS, 1) svgp_small_kernel(SmallArgs a) {
  __shared__ double sred[40];
  __shared__ double s_scale[MAXD];
  __shared__ double s_par[8];  // variance, c, mean_const, likpar
  __shared__ int s_status[2];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int M = a.M, D = a.D, ns = a.n_scale, kind = a.kind;
  const bool centered = a.centered != 0, linear = kind == AGP_KERNEL_LINEAR, direct = D == 1 && !linear;
  const int MM = M * M;
  // ---- workspace carve-up ------------------------------------------------------------------------------------------------------
  // ALLSMEM: every array lives in shared memory (32-bit shared addresses: the ~35 array pointers cost one register each and a
  // phase -- "read what the previous phase wrote, one FMA chain, write" -- pays a shared-memory round trip instead of an L2 one:
  // global stores write through and invalidate the L1 line, so data handed from thread to thread through the global workspace
  // comes back from L2).  Otherwise (M too large for 227 KB; M = 50 with a 100-point tile needs 525 KB) the arrays are taken from
  // shared memory in the order below -- the vectors, the operands of the column-by-column factorisation, the factors every product
  // reads, the first per-point matrices -- until a.smem_doubles are used up, and from the global workspace after that.
  extern __shared__ __align__(16) double sm_pool[];
  double* w = ALLSMEM ? sm_pool : a.ws;
  double* w_s = sm_pool;
  int64_t s_left = ALLSMEM ? 0 : a.smem_doubles;
  auto take = [&](int n) -> double* {
    const int n2 = (n + 1) & ~1;
    if (!ALLSMEM && n2 <= s_left) {
      double* p = w_s;
      w_s += n2;
      s_left -= n2;
      return p;
    }
    double* p = w;
    w += n2;
    return p;
  };
  const int LDB = a.ldb;  // row length of the per-point matrices: the tile size rounded up to 32
  const int MB = M * LDB;
  // vectors first
  double* zs = take(M * D);  // scaled Z
  double* zn = take(M);
  double* mt = take(M);   // whitened mean
  double* ipiv = take(M);  // 1 / Lk_jj
  double* xs = take(LDB * D);  // scaled points of the tile [n][D]
  double* xn = take(LDB);
  double* pdmu = take(LDB);
  double* pdv = take(LDB);
  double* zr = take(M * D);  // raw Z
  double* dZ = take(M * D);
  double* wx = take(M * D);  // kernel-gradient partial sums (kgrad_kernel's wx / wxx)
  double* wxx = take(M * D);
  double* mv = take(M);   // m
  double* g = take(M);
  double* rs = take(M);
  double* dvr = take(M);
  double* dcc = take(M);
  double* mbar = take(M);
  double* acc = take(64 + 2 * MAXD);  // scalar accumulators: [0] E [1] dmu [2] dkxx [3] ds2 [4] dc [8..8+D) ds_lin [8+MAXD .. ) theta
  // the factorisation's operands, then the factors
  double* Lk = take(MM);   // all M x M matrices column-major: X[i + j*M]
  double* W1 = take(MM);   // (the residual rows Y of the inverse during the factorisation)
  double* Li = take(MM);   // Lk^-1 (lower)
  double* Bt = take(MM);
  double* A = take(MB);    // per-point matrices [i][n], n contiguous (ld = LDB)
  double* Cm = take(MB);
  double* G = take(MM);
  double* LiT = take(MM);  // transposed copies: coalesced access when the thread index runs over the column
  double* W2 = take(MM);
  double* W3 = take(MM);
  double* W4 = take(MM);
  double* LkT = take(MM);
  double* Ab = take(MB);
  double* At = take(MB);   // A transposed, [n][i] (ld = M): the reduction over the points of G / g runs with i contiguous
  double* Kuf = take(MB);
  double* DK = take(MB);
  double* Lq = take(MM);
  const double* f = a.flat;
  const double* fZ = f + 4 + ns;
  const double* fm = fZ + M * D;
  const double* fLq = fm + M;

  SM_TICK(0)
  // ---- P0: parameters ------------------------------------------------------------------------------------------------
  if (tid == 0) {
    s_par[0] = f[0];
    s_par[1] = f[1 + ns];
    s_par[2] = f[2 + ns];
    s_par[3] = f[3 + ns];
    s_status[0] = 0;
    s_status[1] = 0;
  }
  if (tid < D) s_scale[tid] = f[1 + (ns == 1 ? 0 : tid)];
  for (int i = tid; i < 64 + 2 * MAXD; i += nt) acc[i] = 0.0;
  for (int i = tid; i < M * D; i += nt) {
    zr[i] = fZ[i];
    dZ[i] = wx[i] = wxx[i] = 0.0;
  }
  for (int i = tid; i < M; i += nt) {
    mv[i] = fm[i];
    g[i] = rs[i] = dvr[i] = dcc[i] = 0.0;
  }
  for (int i = tid; i < MM; i += nt) {
    const int r = (int)(i % M), c = (int)(i / M);
    Lq[i] = (r >= c) ? fLq[i] : 0.0;  // LowerTriangular(A) view, utils.jl:18
    G[i] = 0.0;
  }
  __syncthreads();
  const double variance = s_par[0], lin_c = s_par[1], mean_const = s_par[2];
  LikParams lp = a.lp;
  lp.sigma2 = s_par[3];
  for (int i = tid; i < M; i += nt) {
    double nrm = 0.0;
    for (int d = 0; d < D; d++) {
      const double v = zr[i * D + d] * s_scale[d];
      zs[i * D + d] = v;
      nrm = fma(v, v, nrm);
    }
    zn[i] = nrm;
    if (!(Lq[i + i * M] > 0.0)) atomicExch(&s_status[0], AGP_ERR_DOMAIN);  // logdet(q.Sigma) would throw
  }
  __syncthreads();
  SM_TICK(1)
  // ---- P1: Kuu (lower triangle; build_kuu_kernel's operation order) ------------------------------------------------------
  for (int i = tid; i < MM; i += nt) {
    const int r = (int)(i % M), c = (int)(i / M);
    double v = 0.0;
    if (r >= c) {
      double u;
      if (r == c && !linear) u = 0.0;
      else if (direct) {
        const double df = zs[r] - zs[c];
        u = df * df;
      } else {
        double dot = 0.0;
        for (int d = 0; d < D; d++) dot = fma(zs[c * D + d], zs[r * D + d], dot);  // lo = c, hi = r
        u = u_from_dot(kind, zn[c], zn[r], dot);
      }
      v = variance * kappa(kind, u, lin_c);
      if (r == c) v += a.jitter;
    }
    Lk[i] = v;
  }
  __syncthreads();
  SM_TICK(2)
  // ---- P2 + P3: Cholesky and the inverse of the factor in one right-looking sweep, ONE barrier per column ------------------
  // Unscaled (LDL^T-style) columns: with d_j the pivot, L[i][j] = K(j)[i][j] / sqrt(d_j) is only formed at the end, so that a
  // step reads column j and writes columns > j (no intra-step hazard):   K[i][k] -= K[i][j] K[k][j] / d_j   (i >= k > j).
  // The forward substitution L X = I rides along: residual rows Y (= I at the start), Y[i][c] -= K[i][j] Y[j][c] / d_j for
  // i > j, c <= j; X[j][c] = Y[j][c] / sqrt(d_j).
  double* Y = W1;
  for (int i = tid; i < MM; i += nt) Y[i] = ((i % M) == (i / M)) ? 1.0 : 0.0;
  __syncthreads();
  for (int j = 0; j < M; j++) {
    const double d = Lk[j + j * M];
    if (!(d > 0.0)) {
      if (tid == 0 && s_status[0] == 0) {
        s_status[0] = AGP_ERR_NOT_PD;
        s_status[1] = j + 1;
      }
      break;  // uniform: every thread reads the same d
    }
    const double invd = __drcp_rn(d);
    if (tid == 0) ipiv[j] = rsqrt(d);
    // one warp per column (columns j+1 .. M-1 of the trailing matrix, then columns 0 .. j of Y), lanes over the rows
    {
      const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
      const double* colj = Lk + j * M;
      for (int c = warp; c < M; c += nw) {
        if (c < M - 1 - j) {
          const int k = j + 1 + c;
          const double f = colj[k] * invd;
          double* col = Lk + k * M;
          for (int i = k + lane; i < M; i += 32) col[i] = fma(-colj[i], f, col[i]);
        } else {
          const int cy = c - (M - 1 - j);  // 0 .. j
          double* col = Y + cy * M;
          const double f = col[j] * invd;
          for (int i = j + 1 + lane; i < M; i += 32) col[i] = fma(-colj[i], f, col[i]);
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_status[0] != 0) {
    if (tid == 0) {
      const int nflat = 4 + ns + M * D + M + MM;
      a.out[1 + nflat] = (double)s_status[0];
      a.out[2 + nflat] = (double)s_status[1];
    }
    return;
  }
  for (int e = tid; e < MM; e += nt) {
    const int i = (int)(e % M), j = (int)(e / M);
    double l = 0.0, x = 0.0;
    if (i >= j) {
      l = (i == j) ? 1.0 / ipiv[j] : Lk[e] * ipiv[j];
      x = Y[e] * ipiv[i];
    }
    Li[e] = x;
    LiT[j + i * M] = x;
    LkT[j + i * M] = l;
    W2[e] = l;  // (Lk itself is rewritten after the barrier: other threads still read the unscaled column entries above)
  }
  __syncthreads();
  for (int e = tid; e < MM; e += nt) Lk[e] = W2[e];
  __syncthreads();
  SM_TICK(3)
  // ---- P4: whitened variables ---------------------------------------------------------------------------------------------
  if (!centered) {
    for (int i = tid; i < M; i += nt) mt[i] = mv[i];
    for (int i = tid; i < MM; i += nt) Bt[i] = Lq[i];
  } else {
    for (int i = tid; i < M; i += nt) {
      double s = 0.0;
      for (int k = 0; k <= i; k++) s = fma(Li[i + k * M], mv[k] - mean_const, s);
      mt[i] = s;
    }
    for (int e = tid; e < MM; e += nt) {
      const int i = (int)(e % M), j = (int)(e / M);
      double s = 0.0;
      for (int k = j; k <= i; k++) s = fma(Li[i + k * M], Lq[k + j * M], s);
      Bt[e] = s;
    }
  }
  __syncthreads();

  SM_TICK(4)
  // ---- P5: the points, SM_TILE at a time -------------------------------------------------------------------------------------
  for (int t0 = 0; t0 < a.count; t0 += SM_TILE) {
    const int nb = min(SM_TILE, a.count - t0);
    for (int n = tid; n < nb; n += nt) {
      double nrm = 0.0;
      for (int d = 0; d < D; d++) {
        const double v = a.X[(t0 + n) * D + d] * s_scale[d];
        xs[n * D + d] = v;
        nrm = fma(v, v, nrm);
      }
      xn[n] = nrm;
    }
    __syncthreads();
    for (int e = tid; e < M * nb; e += nt) {
      const int l = e / nb, n = e % nb;
      double u;
      if (direct) {
        const double df = xs[n] - zs[l];
        u = df * df;
      } else {
        double dot = 0.0;
        for (int d = 0; d < D; d++) dot = fma(zs[l * D + d], xs[n * D + d], dot);
        u = u_from_dot(kind, xn[n], zn[l], dot);
      }
      double k, dk;
      kappa_and_du(kind, u, lin_c, k, dk);
      Kuf[l * LDB + n] = variance * k;
      DK[l * LDB + n] = variance * dk;
    }
    __syncthreads();
    // A = Li Kuf
    sm_dmma_gemm(M, nb, M, Li, 1, M, SM_TRI_LE, Kuf, LDB, 1, SM_TRI_NONE, false, [&](int i, int n, double v) {
      A[i * LDB + n] = v;
      At[n * M + i] = v;
    });
    __syncthreads();
    // C = Bt^T A
    sm_dmma_gemm(M, nb, M, Bt, M, 1, SM_TRI_GE, A, LDB, 1, SM_TRI_NONE, false, [&](int j, int n, double v) { Cm[j * LDB + n] = v; });
    __syncthreads();
    SM_TICK(5)
    // marginals + expected log-likelihood, one point per thread (all threads take part in the reductions)
    {
      double E = 0.0, dmu = 0.0, dvar = 0.0, ds2 = 0.0, kf = 0.0;
      const int n = tid;
      if (n < nb) {
        double saa = 0.0, sam = 0.0, scc = 0.0;
        for (int i = 0; i < M; i++) {
          const double av = A[i * LDB + n], cv = Cm[i * LDB + n];
          saa = fma(av, av, saa);
          sam = fma(av, mt[i], sam);
          scc = fma(cv, cv, scc);
        }
        kf = linear ? xn[n] + lin_c : 1.0;
        const double mu = mean_const + sam;
        const double var = variance * kf - saa + scc + 1e-18;  // AbstractGPs default jitter of f_post(x), SVA.jl:354
        if (!(var > 0.0)) atomicExch(&s_status[0], AGP_ERR_DOMAIN);
        expected_loglik(lp, mu, var, a.y[t0 + n], E, dmu, dvar, ds2, a.point_base + t0 + n);
        dmu *= a.scale;
        dvar *= a.scale;
        ds2 *= a.scale;
        pdmu[n] = dmu;
        pdv[n] = dvar;
      }
      // (SM_TILE <= blockDim.x: every point of the tile has its own thread)
      static_assert(SM_TILE <= SM_THREADS, "one thread per point of a tile");
      const double r0 = sm_block_sum(E, sred), r1 = sm_block_sum(dmu, sred), r2 = sm_block_sum(dvar * kf, sred), r3 = sm_block_sum(ds2, sred);
      const double r4 = linear ? sm_block_sum(dvar, sred) : 0.0;
      if (tid == 0) {
        acc[0] += r0;
        acc[1] += r1;
        acc[2] += r2;
        acc[3] += r3;
        acc[4] += linear ? r4 * variance : 0.0;
      }
      if (linear && a.want_grad) {
        for (int d = 0; d < D; d++) {
          double t = 0.0;
          if (n < nb) {
            const double x = a.X[(t0 + n) * D + d];
            t = dvar * 2.0 * variance * s_scale[d] * x * x;
          }
          const double r = sm_block_sum(t, sred);
          if (tid == 0) acc[8 + d] += r;
        }
      }
    }
    __syncthreads();
    SM_TICK(6)
    if (!a.want_grad) continue;
    // Ab = dmu (x) mt + 2 dv (Bt C - A)
    sm_dmma_gemm(M, nb, M, Bt, 1, M, SM_TRI_LE, Cm, LDB, 1, SM_TRI_NONE, false,
                 [&](int i, int n, double v) { Ab[i * LDB + n] = fma(pdmu[n], mt[i], 2.0 * pdv[n] * (v - A[i * LDB + n])); });
    __syncthreads();
    // (dv A)^T for the product behind G, in the layout of At; A itself is no longer needed
    for (int e = tid; e < M * nb; e += nt) A[e] = pdv[e / M] * At[e];
    // Kb = Li^T Ab  (into Cm)
    sm_dmma_gemm(M, nb, M, Li, M, 1, SM_TRI_GE, Ab, LDB, 1, SM_TRI_NONE, false, [&](int j, int n, double v) { Cm[j * LDB + n] = v; });
    __syncthreads();  // the scaled copy of A^T is complete
    SM_TICK(7)
    // G += (dv A) A^T (lower blocks), g += A dmu
    sm_dmma_gemm(M, M, nb, A, 1, M, SM_TRI_NONE, At, M, 1, SM_TRI_NONE, true, [&](int i, int j, double v) {
      if (i >= j) G[i + j * M] += v;
    });
    for (int i = tid; i < M; i += nt) {
      double s0 = 0.0;
      for (int n = 0; n < nb; n++) s0 = fma(pdmu[n], At[n * M + i], s0);
      g[i] += s0;
    }
    __syncthreads();
    SM_TICK(8)
    // kernel-gradient partial sums of this tile (kgrad_kernel's quantities): W = Kb * variance * kappa'(u); one warp per (row, d)
    {
      const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
      for (int e = warp; e < M * (D + 1); e += nw) {
        const int l = e / (D + 1), q = e % (D + 1);
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
        if (q == D) {
          for (int n = lane; n < nb; n += 32) {
            const double kb = Cm[l * LDB + n];
            v0 = fma(kb, DK[l * LDB + n], v0);
            v1 = fma(kb, Kuf[l * LDB + n], v1);
            v2 += kb;
          }
        } else {
          for (int n = lane; n < nb; n += 32) {
            const double W = Cm[l * LDB + n] * DK[l * LDB + n];
            const double x = xs[n * D + q];
            const double wxd = W * x;
            v0 += wxd;
            v1 = fma(wxd, x, v1);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          v0 += __shfl_xor_sync(0xffffffffu, v0, o);
          v1 += __shfl_xor_sync(0xffffffffu, v1, o);
          v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        }
        if (lane == 0) {
          if (q == D) {
            rs[l] += v0;
            dvr[l] += v1 / variance;  // sum Kb kappa(u)
            dcc[l] += v2;
          } else {
            wx[l * D + q] += v0;
            wxx[l * D + q] += v1;
          }
        }
      }
    }
    __syncthreads();
  }
  const int nflat = 4 + ns + M * D + M + MM;
  double* out = a.out;
  SM_TICK(9)
  // ---- KL ---------------------------------------------------------------------------------------------------------------
  double kl;
  {
    double tr = 0.0, mm = 0.0, ldq = 0.0, ldk = 0.0;
    for (int i = tid; i < MM; i += nt) tr = fma(Bt[i], Bt[i], tr);
    for (int j = tid; j < M; j += nt) {
      mm = fma(mt[j], mt[j], mm);
      ldq += log(Lq[j + j * M]);
      if (centered) ldk += log(Lk[j + j * M]);
    }
    const double s0 = sm_block_sum(tr, sred), s1 = sm_block_sum(mm, sred), s2 = sm_block_sum(ldq, sred), s3 = sm_block_sum(ldk, sred);
    kl = 0.5 * (s0 + s1 - (double)M) + s3 - s2;
  }
  if (!a.want_grad) {
    if (tid == 0) {
      out[0] = acc[0] * a.scale - kl;
      out[1 + nflat] = (double)s_status[0];
      out[2 + nflat] = 0.0;
    }
    return;
  }
  SM_TICK(10)
  // ---- P6: replicated epilogue (agp_svgp_finish) -----------------------------------------------------------------------------
  // data part of dZ / theta from the accumulated partial sums (kgrad_finish_kernel, zfac = 1)
  double* theta = acc + 8 + MAXD;  // [0] dvariance [1] dc [2 + d] ds_d
  {
    double tv = 0.0, tc = 0.0;
    for (int l = tid; l < M; l += nt) {
      tv += dvr[l];
      tc += dcc[l];
    }
    const double r0 = sm_block_sum(tv, sred), r1 = sm_block_sum(tc, sred);
    if (tid == 0) {
      theta[0] = r0;
      theta[1] = linear ? variance * r1 : 0.0;
    }
    for (int d = 0; d < D; d++) {
      double v = 0.0;
      for (int l = tid; l < M; l += nt) {
        const double z = zs[l * D + d], sx = wx[l * D + d], sxx = wxx[l * D + d];
        v += linear ? z * sx : fma(z, fma(z, rs[l], -2.0 * sx), sxx);
      }
      const double r = sm_block_sum(v, sred);
      if (tid == 0) theta[2 + d] = (s_scale[d] != 0.0) ? 2.0 / s_scale[d] * r : 0.0;
    }
    for (int i = tid; i < M * D; i += nt) {
      const int l = (int)(i / D), d = (int)(i % D);
      const double sd = s_scale[d];
      dZ[i] = linear ? sd * wx[i] : 2.0 * sd * (zs[i] * rs[l] - wx[i]);
    }
  }
  __syncthreads();
  SM_TICK(11)
  // G: lower -> symmetric
  for (int e = tid; e < MM; e += nt) {
    const int i = (int)(e % M), j = (int)(e / M);
    if (i < j) G[e] = G[j + i * M];
  }
  __syncthreads();
  // W1 = P1 = Bt Bt^T - I
  sm_dmma_gemm(M, M, M, Bt, 1, M, SM_TRI_LE, Bt, M, 1, SM_TRI_LE, false, [&](int i, int j, double v) { W1[i + j * M] = (i == j) ? v - 1.0 : v; });
  __syncthreads();
  // W2 = Asum = mt g^T + 2 P1 G ;  W3 = Bt-bar = tril(2 G Bt) [- Bt]
  sm_dmma_gemm(M, M, M, W1, 1, M, SM_TRI_NONE, G, 1, M, SM_TRI_NONE, false, [&](int i, int j, double v) { W2[i + j * M] = fma(mt[i], g[j], 2.0 * v); });
  sm_dmma_gemm(M, M, M, G, 1, M, SM_TRI_NONE, Bt, 1, M, SM_TRI_GE, false,
               [&](int i, int j, double v) { W3[i + j * M] = (i >= j) ? 2.0 * v - (centered ? Bt[i + j * M] : 0.0) : 0.0; });
  __syncthreads();
  // W4 = V = Li^T Asum
  sm_dmma_gemm(M, M, M, LiT, 1, M, SM_TRI_GE, W2, 1, M, SM_TRI_NONE, false, [&](int i, int j, double v) { W4[i + j * M] = v; });
  double summbar = 0.0;
  if (centered) {
    // mbar = Li^T (g - mt)
    for (int i = tid; i < M; i += nt) {
      double s = 0.0;
      for (int k = i; k < M; k++) s = fma(LiT[i + k * M], g[k] - mt[k], s);
      mbar[i] = s;
    }
    __syncthreads();
    double v = 0.0;
    for (int i = tid; i < M; i += nt) v += mbar[i];
    summbar = sm_block_sum(v, sred);
    // W1 = Y = Li^T Bt-bar  (P1 is no longer needed: Asum has been formed)
    sm_dmma_gemm(M, M, M, LiT, 1, M, SM_TRI_GE, W3, 1, M, SM_TRI_GE, false, [&](int i, int j, double v) { W1[i + j * M] = v; });
    __syncthreads();
    // W2 = X2 = Y Bt^T + mbar mt^T   (Asum is no longer needed: V has been formed)
    sm_dmma_gemm(M, M, M, W1, 1, M, SM_TRI_NONE, Bt, M, 1, SM_TRI_LE, false, [&](int i, int j, double v) { W2[i + j * M] = fma(mbar[i], mt[j], v); });
  }
  __syncthreads();
  SM_TICK(12)
  // dLq, dm, scalars: everything that does not depend on the Cholesky pullback
  {
    const double* src = centered ? W1 : W3;
    double* oLq = out + 1 + 4 + ns + M * D + M;
    for (int e = tid; e < MM; e += nt) {
      const int r = (int)(e % M), c = (int)(e / M);
      double v = 0.0;
      if (r >= c) {
        v = src[e] - (centered ? 0.0 : Lq[e]);
        if (r == c) v += 1.0 / Lq[e];
      }
      oLq[e] = v;
    }
    double* om = out + 1 + 4 + ns + M * D;
    for (int i = tid; i < M; i += nt) om[i] = centered ? mbar[i] : g[i] - mv[i];
  }
  // G <- Lbar = -tril(V) [- tril(X2) - diag(1 / Lk_jj)]     (G is no longer needed)
  for (int e = tid; e < MM; e += nt) {
    const int r = (int)(e % M), c = (int)(e / M);
    double v = 0.0;
    if (r >= c) {
      v = -W4[e];
      if (centered) {
        v -= W2[e];
        if (r == c) v -= 1.0 / Lk[e];
      }
    }
    G[e] = v;
  }
  __syncthreads();
  // W3 = Phi(Lk^T Lbar): lower, diagonal halved
  sm_dmma_gemm(M, M, M, LkT, 1, M, SM_TRI_GE, G, 1, M, SM_TRI_GE, false,
               [&](int i, int j, double v) { W3[i + j * M] = (i > j) ? v : (i == j ? 0.5 * v : 0.0); });
  __syncthreads();
  // W4 = Y1 = Li^T Phi ;  W2 = Y2 = Li^T Y1^T ;  G = Kuu-bar = (Y2 + Y2^T) / 2
  sm_dmma_gemm(M, M, M, LiT, 1, M, SM_TRI_GE, W3, 1, M, SM_TRI_GE, false, [&](int i, int j, double v) { W4[i + j * M] = v; });
  __syncthreads();
  sm_dmma_gemm(M, M, M, LiT, 1, M, SM_TRI_GE, W4, M, 1, SM_TRI_NONE, false, [&](int i, int j, double v) { W2[i + j * M] = v; });
  __syncthreads();
  for (int e = tid; e < MM; e += nt) {
    const int i = (int)(e % M), j = (int)(e / M);
    G[e] = 0.5 * (W2[e] + W2[j + i * M]);
  }
  __syncthreads();
  SM_TICK(13)
  // Kuu part of dZ / theta: contraction of Kuu-bar with the derivatives of k(z_l, z_n); both arguments move -> zfac = 2.
  // W3 = Kuu-bar .* variance kappa'(u), W4 = kappa(u) (both symmetric), then the same per-row sums as for the data part.
  for (int e = tid; e < MM; e += nt) {
    const int l = (int)(e % M), n = (int)(e / M);
    double u;
    if (direct) {
      const double df = zs[n] - zs[l];
      u = df * df;
    } else {
      double dot = 0.0;
      const int lo = min(l, n), hi = max(l, n);
      for (int d = 0; d < D; d++) dot = fma(zs[lo * D + d], zs[hi * D + d], dot);
      u = u_from_dot(kind, zn[lo], zn[hi], dot);
    }
    double k, dk;
    kappa_and_du(kind, u, lin_c, k, dk);
    W3[e] = G[e] * variance * dk;
    W4[e] = k;
  }
  __syncthreads();
  {
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int e = warp; e < M * (D + 1); e += nw) {
      const int l = e / (D + 1), q = e % (D + 1);
      double v0 = 0.0, v1 = 0.0, v2 = 0.0;
      if (q == D) {
        for (int n = lane; n < M; n += 32) {
          const double kb = G[n + l * M];
          v0 += W3[n + l * M];
          v1 = fma(kb, W4[n + l * M], v1);
          v2 += kb;
        }
      } else {
        for (int n = lane; n < M; n += 32) {
          const double x = zs[n * D + q];
          const double wxd = W3[n + l * M] * x;
          v0 += wxd;
          v1 = fma(wxd, x, v1);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
      }
      if (lane == 0) {
        if (q == D) {
          rs[l] = v0;
          dvr[l] = v1;
          dcc[l] = v2;
        } else {
          wx[l * D + q] = v0;
          wxx[l * D + q] = v1;
        }
      }
    }
  }
  __syncthreads();
  {
    double tv = 0.0, tc = 0.0;
    for (int l = tid; l < M; l += nt) {
      tv += dvr[l];
      tc += dcc[l];
    }
    const double r0 = sm_block_sum(tv, sred), r1 = sm_block_sum(tc, sred);
    if (tid == 0) {
      theta[0] += r0;
      theta[1] += linear ? variance * r1 : 0.0;
    }
    for (int d = 0; d < D; d++) {
      double v = 0.0;
      for (int l = tid; l < M; l += nt) {
        const double z = zs[l * D + d], sx = wx[l * D + d], sxx = wxx[l * D + d];
        v += linear ? z * sx : fma(z, fma(z, rs[l], -2.0 * sx), sxx);
      }
      const double r = sm_block_sum(v, sred);
      if (tid == 0) theta[2 + d] += (s_scale[d] != 0.0) ? 2.0 / s_scale[d] * r : 0.0;
    }
    for (int i = tid; i < M * D; i += nt) {
      const int l = (int)(i / D), d = (int)(i % D);
      const double sd = s_scale[d];
      dZ[i] += 2.0 * (linear ? sd * wx[i] : 2.0 * sd * (zs[i] * rs[l] - wx[i]));
    }
  }
  __syncthreads();
  SM_TICK(14)
  {
    double* oZ = out + 1 + 4 + ns;
    for (int i = tid; i < M * D; i += nt) oZ[i] = dZ[i];
    if (tid == 0) {
      out[0] = acc[0] * a.scale - kl;
      out[1 + 0] = theta[0] + acc[2];                 // dvariance
      out[1 + 1 + ns] = theta[1] + acc[4];            // dlinear_c
      out[1 + 2 + ns] = acc[1] - (centered ? summbar : 0.0);  // dmean_const
      out[1 + 3 + ns] = acc[3];                       // d likelihood parameter
      if (ns == 1) {
        double s = 0.0;
        for (int d = 0; d < D; d++) s += theta[2 + d] + (linear ? acc[8 + d] : 0.0);
        out[1 + 1] = s;
      } else {
        for (int d = 0; d < D; d++) out[1 + 1 + d] = theta[2 + d] + (linear ? acc[8 + d] : 0.0);
      }
      out[1 + nflat] = (double)s_status[0];
      out[2 + nflat] = 0.0;
    }
  }
  SM_TICK(15)
}

}  // namespace agp
