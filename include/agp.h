/*
 * agp.h -- C ABI of the B200-native SVGP-ELBO / Laplace hot path.
 *
 * The reference (ApproximateGPs.jl) has no FFI boundary: its "plugin API" for this path is Julia
 * multiple dispatch on the methods below.  Each entry point here names the reference method
 * body it replaces (paths relative to /root/reference); INTEGRATION.md shows the Julia `ccall`
 * shim that overloads those methods and the Python ctypes binding used in this Julia-less image.
 *
 * Conventions
 *   - every function returns an int32 status (AGP_OK == 0); agp_last_error_string() describes
 *     the last failure on the calling thread;
 *   - all matrices are column-major (Julia native).  Inputs Z / X are "point-major": one point
 *     is D contiguous doubles (ColVecs(D x N)); AGP_FEATURE_MAJOR accepts RowVecs(N x D);
 *   - host arrays are only read during the call; outputs are caller-allocated host buffers;
 *   - there is NO CPU fallback: a missing / failing device is an error, an unsupported
 *     kernel / likelihood is AGP_ERR_UNSUPPORTED.
 */
#ifndef AGP_H_
#define AGP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (error conventions of SURVEY.md section 8b) ------------------------------ */
#define AGP_OK 0
#define AGP_ERR_INVALID 1     /* bad argument (AssertionError / ArgumentError in the shim)        */
#define AGP_ERR_UNSUPPORTED 2 /* kernel / likelihood / mean not implemented on device            */
#define AGP_ERR_NOT_PD 3      /* PosDefException: cholesky(Kuu) or cholesky(B) failed            */
#define AGP_ERR_DOMAIN 4      /* DomainError: sqrt of negative W (Laplace.jl:214) / variance     */
#define AGP_ERR_CUDA 5
#define AGP_ERR_NCCL 6
#define AGP_ERR_ALLOC 7

/* ---- enums --------------------------------------------------------------------------------- */
#define AGP_KERNEL_SE 0       /* SqExponentialKernel: exp(-d^2/2)                                */
#define AGP_KERNEL_MATERN32 1 /* (1+sqrt3 d) exp(-sqrt3 d)                                       */
#define AGP_KERNEL_MATERN52 2 /* (1+sqrt5 d+5d^2/3) exp(-sqrt5 d)                                */
#define AGP_KERNEL_LINEAR 3   /* x.y + c                                                         */
#define AGP_KERNEL_SUM 4      /* KernelSum of stationary components (agp_kernel.components)      */
#define AGP_KERNEL_PRODUCT 5  /* KernelProduct of stationary components                          */
#define AGP_MAX_COMPONENTS 4

#define AGP_LIK_GAUSSIAN 0        /* GaussianLikelihood(sigma2)                                  */
#define AGP_LIK_BERNOULLI_LOGIT 1 /* BernoulliLikelihood() (logistic link)                       */
#define AGP_LIK_POISSON_EXP 2     /* PoissonLikelihood() (exp link)                              */
#define AGP_LIK_EXPONENTIAL_EXP 3 /* ExponentialLikelihood() (exp link): Exponential(scale = exp(f)) */
#define AGP_LIK_GAMMA_EXP 4       /* GammaLikelihood(alpha) (exp link): Gamma(alpha, scale = exp(f)); alpha in `sigma2` */
#define AGP_LIK_BERNOULLI_PROBIT 5 /* BernoulliLikelihood(ProbitLink()): Bernoulli(normcdf(f))    */

#define AGP_EXPECT_DEFAULT 0       /* GPLikelihoods.DefaultExpectationMethod()                   */
#define AGP_EXPECT_ANALYTIC 1      /* AnalyticExpectation()                                      */
#define AGP_EXPECT_GAUSS_HERMITE 2 /* GaussHermiteExpectation(n): caller passes nodes / weights  */
#define AGP_EXPECT_MONTE_CARLO 3   /* MonteCarloExpectation(n): n reparameterised samples / point */

#define AGP_NONCENTERED 0 /* SparseVariationalApproximation{NonCentered} (the default, SVA.jl:93) */
#define AGP_CENTERED 1    /* SparseVariationalApproximation{Centered}                             */

#define AGP_POINT_MAJOR 0   /* X[i*ldx + d]  (ColVecs(D x N), or Vector for D == 1)               */
#define AGP_FEATURE_MAJOR 1 /* X[d*ldx + i]  (RowVecs(N x D))                                     */

#define AGP_Y_F64 0
#define AGP_Y_F32 1
#define AGP_Y_I64 2
#define AGP_Y_U8 3 /* Bool */

#define AGP_HOST 0
#define AGP_DEVICE 1

#define AGP_COMPUTE_F64 0 /* reference precision: every stage in Float64 (DMMA)                                   */
#define AGP_COMPUTE_F32 1 /* "Float32 fast mode" (a Float32 GP in the reference, SVA.jl:59-62 is type-generic): S2 / S4 / S6 run as 3xTF32 split
                           * products on the tcgen05 tensor path (partial sums carried over in Float64 every 64 k), the two triangular solves as
                           * exact-accumulation INT8 products with the explicit inverse (tcgen05.mma.kind::i8; forward: 7 slices of 7 bits =
                           * Float64-accurate, reverse: 5); Kuf, per-point stage, reductions and the O(M^3) epilogue stay Float64.  Parity 1e-4. */
#define AGP_COMPUTE_F32_TC_SOLVE 2 /* as AGP_COMPUTE_F32, with the reverse-pass solve Kb = Lk^-T Ab as a 3xTF32 product with the
                           * explicit inverse too (1.2x faster again; its error on dZ / d theta grows with cond(Lk): 1.4e-4 at
                           * the M = 1024 SqExponential twin of BASELINE config 4, below 1e-4 on the others)                */

#define AGP_COMPUTE_F64_EMU 3 /* Float64 results inside the Float64 tolerance (1e-10), with the reverse pass's point-sum product G += (dv A) A^T
                           * and the forward product C = Bt^T A (for M >= 768) formed by FP64-ACCURATE EMULATION on the INT8 tensor path: both operands cut into seven signed 7-bit slices per element
                           * (round to nearest) under one power-of-two scale per inducing row, the 28 slice products accumulated exactly in INT32
                           * (tcgen05.mma.kind::i8, tensor memory), recombined in Float64.  Every other stage as AGP_COMPUTE_F64.  Opt-in: the default
                           * Float64 mode uses Float64 DMMA / FMA arithmetic only.                                                         */

#define AGP_MAX_GH_POINTS 128
#define AGP_MAX_D 32 /* compile-time bound of the device code (kfun.cuh MAXD): larger D is AGP_ERR_UNSUPPORTED */

typedef struct agp_ctx agp_ctx;
typedef struct agp_dataset agp_dataset;
typedef struct agp_laplace_cache agp_laplace_cache;

/* `variance * (base o ScaleTransform(s))` (n_scale == 1) or `... o ARDTransform(v)` (n_scale == D);
 * KernelFunctions semantics as reached through cov(f.prior, z, x) at SVA.jl:216.                 */
/* One term / factor of a KernelSum / KernelProduct (KernelFunctions `k1 + k2`, `k1 * k2`, reached through the same cov(f.prior, z, x)
 * at SVA.jl:216): `variance * (base o ScaleTransform(inv_lengthscale))` with a stationary base (SE, Matern32, Matern52).      */
typedef struct {
  int32_t kind;
  double variance;
  double inv_lengthscale;
} agp_kernel_component;

typedef struct {
  int32_t kind;
  int32_t n_scale;
  double variance;
  const double* inv_lengthscale; /* host, n_scale entries */
  double linear_c;
  /* kind == AGP_KERNEL_SUM / AGP_KERNEL_PRODUCT:
   *   k = variance * ((c_1 + ... + c_n) o T)   or   variance * ((c_1 * ... * c_n) o T),   T = ScaleTransform / ARDTransform(inv_lengthscale)
   * i.e. the components share the outer transform T (pass inv_lengthscale = {1} for none) and each has its own scalar lengthscale,
   * so that one scaled squared distance serves all of them.  Ignored (may be 0 / NULL) for the plain kinds.                      */
  int32_t n_components;                   /* 1..AGP_MAX_COMPONENTS */
  const agp_kernel_component* components; /* host */
} agp_kernel;

typedef struct {
  int32_t kind;
  double sigma2; /* the likelihood's scalar parameter: GaussianLikelihood sigma2, GammaLikelihood alpha; ignored otherwise */
} agp_likelihood;

typedef struct {
  int32_t method;
  int32_t n_points;      /* Gauss-Hermite: nodes (<= AGP_MAX_GH_POINTS); Monte Carlo: samples per point */
  const double* nodes;   /* host; FastGaussQuadrature.gausshermite(n)[1] */
  const double* weights; /* host; FastGaussQuadrature.gausshermite(n)[2] */
  uint64_t seed;         /* Monte Carlo only: f = mu + sigma * eps, eps = N(0,1) from Philox4x32-10 keyed by `seed` with
                          * counter (global point index, sample index).  GPLikelihoods draws from Julia's task-local RNG, so
                          * only the distribution (not the stream) can match the reference; the oracle restates this stream. */
} agp_expectation;

/* One `SparseVariationalApproximation(fz, q)` (SVA.jl:59-95) plus the likelihood / quadrature of
 * the `elbo` call: fz = GP(mean_const, kernel)(Z, jitter); q = MvNormal(m, PDMat(Cholesky(Lq))). */
typedef struct {
  agp_kernel kernel;
  double mean_const; /* ConstMean value; 0 for ZeroMean */
  int32_t M;
  int32_t D;
  const double* Z; /* host, point-major M x D */
  double jitter;   /* fz.Sigma_y[1] */
  const double* m; /* host, M */
  const double* Lq; /* host, column-major M x M, lower triangle read (the PDMat Cholesky factor) */
  int32_t ldLq;
  int32_t parametrization;
  agp_likelihood lik;
  agp_expectation expect;
  int32_t compute_dtype; /* AGP_COMPUTE_F64 (0, default) | AGP_COMPUTE_F32 | AGP_COMPUTE_F32_TC_SOLVE | AGP_COMPUTE_F64_EMU: arithmetic of elbo / elbo_grad; inputs and outputs stay Float64 */
} agp_svgp_params;

/* Gradient of the ELBO (unit cotangent); every pointer is a caller-allocated host buffer or NULL.
 * These are the fields the new ChainRulesCore.rrule(elbo, ...) fills.     */
typedef struct {
  double* dm;               /* M                                             */
  double* dLq;              /* column-major M x M (ld = M), strict upper = 0 */
  double* dZ;               /* point-major M x D                             */
  double* dvariance;        /* 1                                             */
  double* dinv_lengthscale; /* n_scale                                       */
  double* dlinear_c;        /* 1                                             */
  double* dmean_const;      /* 1                                             */
  double* dlik_sigma2;      /* 1: d / d(likelihood parameter) (sigma2 or alpha)  */
  double* dcomp_variance;        /* n_components (AGP_KERNEL_SUM / AGP_KERNEL_PRODUCT only) */
  double* dcomp_inv_lengthscale; /* n_components                                            */
} agp_svgp_grads;

/* ---- context ------------------------------------------------------------------------------- */
int32_t agp_ctx_create(int32_t device, agp_ctx** out);
int32_t agp_ctx_destroy(agp_ctx* ctx);
const char* agp_last_error_string(void);
/* `info` of the PosDefException behind the calling thread's last AGP_ERR_NOT_PD: the 1-based column at which cholesky failed. */
int32_t agp_last_error_info(void);
/* Version / build probe: returns the compute capability the kernels were built for (100). */
int32_t agp_build_arch(void);
/* The CUDA stream (cudaStream_t) every kernel of this context is launched on. */
int32_t agp_ctx_stream(agp_ctx* ctx, void** stream_out);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int32_t agp_ctx_launch_count(agp_ctx* ctx, int64_t* out);

/* Per-kernel-class device timing (CUDA events on the context's stream), used by bench.py for its
 * roofline figures.  agp_ctx_profile_read synchronises, returns the accumulated milliseconds and
 * launch-group counts per class since the last read and resets them; class names come from
 * agp_profile_class_name (NULL past the last class).                                            */
int32_t agp_ctx_profile(agp_ctx* ctx, int32_t enable);
int32_t agp_ctx_profile_read(agp_ctx* ctx, int32_t max_classes, double* ms_out, int64_t* count_out,
                             int32_t* n_classes);
const char* agp_profile_class_name(int32_t cls);

/* ---- data-parallel communicator (NCCL over NVLink; one process per GPU) --------------------- */
/* unique_id: the 128 bytes of an ncclUniqueId created by rank 0 (agp_comm_unique_id).          */
int32_t agp_comm_unique_id(void* unique_id_128);
int32_t agp_comm_init(agp_ctx* ctx, int32_t nranks, int32_t rank, const void* unique_id_128);
int32_t agp_comm_destroy(agp_ctx* ctx);

/* ---- datasets: the device-resident (x, y) of `lfx.fx.x` / `y` in elbo(sva, lfx, y) ----------- */
int32_t agp_dataset_create(agp_ctx* ctx, int64_t capacity, int32_t D, agp_dataset** out);
/* (Re)fill from host or device memory; N <= capacity.  location: AGP_HOST | AGP_DEVICE.        */
int32_t agp_dataset_upload(agp_dataset* ds, const void* X, int64_t N, int64_t ldx, int32_t layout,
                           const void* y, int32_t ytype, int32_t location);
/* The same for Float32 inputs (a Float32 caller's ColVecs / RowVecs / Vector): half the bytes over PCIe; the points are widened
 * to the Float64 the kernels compute in (the Float32 *arithmetic* mode of north_star is not built).        */
int32_t agp_dataset_upload_f32(agp_dataset* ds, const float* X, int64_t N, int64_t ldx, int32_t layout,
                               const void* y, int32_t ytype, int32_t location);
int32_t agp_dataset_size(agp_dataset* ds, int64_t* N, int32_t* D);
int32_t agp_dataset_destroy(agp_dataset* ds);

/* ---- SVGP ------------------------------------------------------------------------------------ */
/* Replaces AbstractGPs.elbo(sva, lfx::LatentFiniteGP, y; num_data, quadrature) -- SVA.jl:340-360 --
 * its FiniteGP wrapper (:307-317, the shim passes GaussianLikelihood(fx.Sigma_y[1])) and
 * API.approx_lml (:276-280), over points [offset, offset+count) of the dataset, and returns the
 * Zygote gradient the reference would produce.  num_data <= 0 means num_data = count.
 * When a communicator is attached, `count` is this rank's share, num_data the global count and
 * the partial sums are all-reduced once; every rank returns the same
 * values.                                                                                      */
int32_t agp_svgp_elbo_grad(agp_ctx* ctx, agp_dataset* ds, int64_t offset, int64_t count,
                           const agp_svgp_params* p, double num_data, int64_t global_batch,
                           double* elbo_out, agp_svgp_grads* grads_out);
/* Forward only (no gradient buffers are touched). */
int32_t agp_svgp_elbo(agp_ctx* ctx, agp_dataset* ds, int64_t offset, int64_t count,
                      const agp_svgp_params* p, double num_data, int64_t global_batch,
                      double* elbo_out);

/* Flat-vector form for optimisers that work on one parameter vector (ParameterHandling.flatten + Optim in
 * examples/b-classification/script.jl:102-142, Flux.params in examples/a-regression/script.jl:188): `tmpl` fixes everything that is
 * not optimised (kernel kind, n_scale, M, D, jitter, parametrization, likelihood kind, expectation method); `flat` holds
 *   [variance | inv_lengthscale (n_scale) | linear_c | mean_const | likelihood parameter | Z (M*D point-major) | m (M) | Lq (M*M column-major)]
 * (agp_svgp_flat_size doubles) and `flat_grad` (same layout, or NULL for the value only) receives the gradient: one pointer in, one out. */
int32_t agp_svgp_flat_size(const agp_svgp_params* tmpl, int64_t* n_doubles);
int32_t agp_svgp_elbo_grad_flat(agp_ctx* ctx, agp_dataset* ds, int64_t offset, int64_t count, const agp_svgp_params* tmpl,
                                const double* flat, double num_data, int64_t global_batch, double* elbo_out, double* flat_grad);

/* Optimiser-step handle (the training loops of examples/a-regression/script.jl:176-194 -- 30 000
 * ELBO+gradient evaluations on minibatches of 100 points with M = 20 -- and examples/b-classification/script.jl:124-142):
 * `tmpl` fixes what an optimiser does not change, as for agp_svgp_elbo_grad_flat; every eval takes the current parameters as one
 * flat vector (same layout) and returns the gradient in that layout.  What a step needs stays resident behind the handle
 * (pinned parameter / result buffers the kernel reads and writes directly, workspace, quadrature table); a small problem
 * (M <= 128 and M^2 * count <= 4e6, no communicator) is evaluated by ONE kernel launch of one CTA, anything else by the
 * throughput path.  flat_grad == NULL -> value only.                                                                    */
typedef struct agp_svgp_stepper agp_svgp_stepper;
int32_t agp_svgp_stepper_create(agp_ctx* ctx, const agp_svgp_params* tmpl, agp_svgp_stepper** out);
int32_t agp_svgp_stepper_flat_size(agp_svgp_stepper* s, int64_t* n_doubles);
int32_t agp_svgp_stepper_eval(agp_svgp_stepper* s, agp_dataset* ds, int64_t offset, int64_t count, const double* flat,
                              double num_data, int64_t global_batch, double* elbo_out, double* flat_grad);
int32_t agp_svgp_stepper_counts(agp_svgp_stepper* s, int64_t* n_small, int64_t* n_large);
/* development aid: 16 phase time stamps (ns) of the last one-launch evaluation; zeros unless the library was built with -DAGP_SMALL_TIMING */
int32_t agp_svgp_stepper_phase_ticks(agp_svgp_stepper* s, double* ticks16);
int32_t agp_svgp_stepper_destroy(agp_svgp_stepper* s);

/* Split-phase form of agp_svgp_elbo_grad for hosts that own the collective (torch.distributed,
 * MPI.jl): sweep -> caller all-reduces the packed float64 buffer in place (sum) -> finish.     */
int32_t agp_svgp_sweep(agp_ctx* ctx, agp_dataset* ds, int64_t offset, int64_t count,
                       const agp_svgp_params* p, double num_data, int64_t global_batch,
                       int32_t want_grad);
int32_t agp_svgp_reduce_buffer(agp_ctx* ctx, void** device_ptr, int64_t* n_doubles);
int32_t agp_svgp_finish(agp_ctx* ctx, double* elbo_out, agp_svgp_grads* grads_out);

/* Replaces _prior_kl(sva) -- SVA.jl:362 (Centered), :364-373 (NonCentered). */
int32_t agp_svgp_prior_kl(agp_ctx* ctx, const agp_svgp_params* p, double* kl_out);

/* Replaces posterior(sva).data -- SVA.jl:115-136 / :160-187: lower Cholesky factor of
 * cov(fz) (M x M col-major), B (M x M col-major, lower) and alpha (M).  Any output may be NULL. */
int32_t agp_svgp_posterior(agp_ctx* ctx, const agp_svgp_params* p, double* Lk_out, double* B_out,
                           double* alpha_out);

/* Replaces StatsBase.mean_and_var(posterior(sva), x) -- SVA.jl:246-253 (and mean / var :208-235):
 * Xnew host point-major n x D; mu_out / var_out host, n entries (var WITHOUT the 1e-18 jitter). */
int32_t agp_svgp_mean_and_var(agp_ctx* ctx, const agp_svgp_params* p, const double* Xnew, int64_t n,
                              double* mu_out, double* var_out);

/* Replaces StatsBase.mean_and_cov(posterior(sva), x) -- SVA.jl:237-244 -- Statistics.cov(f, x) :223-228 and the
 * cross-covariance cov(f, x, y) :255-264.  X1 / X2 host point-major; X2 == NULL means y = x (the one-argument
 * kernel matrix with an exactly zero-distance diagonal).  mu1_out (n1, optional) = mean at X1; cov_out host
 * column-major n1 x n2.  Dense output: at most 16384 points per argument (AGP_ERR_UNSUPPORTED beyond).        */
int32_t agp_svgp_mean_and_cov(agp_ctx* ctx, const agp_svgp_params* p, const double* X1, int64_t n1,
                              const double* X2, int64_t n2, double* mu1_out, double* cov_out);

/* Replaces cov(f.prior, x) / cov(f.prior, x, y), i.e. KernelFunctions.kernelmatrix(k, x[, y]) as the reference reaches it at
 * SVA.jl:216, :227, :263 and Laplace.jl:174, :427 (without any observation noise): X1 / X2 host point-major n x D, X2 == NULL ->
 * the one-argument form with an exactly zero-distance diagonal; K_out host column-major n1 x n2 (n <= 16384).                */
int32_t agp_kernel_matrix(agp_ctx* ctx, const agp_kernel* kernel, int32_t D, const double* X1, int64_t n1, const double* X2,
                          int64_t n2, double* K_out);

/* Measurement aid (bench.py): achieved FP64 TFLOP/s of a register-resident instruction chain on this device, now --
 * which = 0: DMMA.8x8x4 (the tensor-path roofline denominator), 1: DFMA.  No reference counterpart.                          */
int32_t agp_fp64_peak(agp_ctx* ctx, int32_t which, double* tflops_out);

/* ---- Laplace --------------------------------------------------------------------------------- */
/* Newton callback(fnew, cache) of _newton_inner_loop (Laplace.jl:263-265): called after every Newton
 * step with a view of the device-resident LaplaceCache (fields fetched lazily with
 * agp_laplace_cache_fetch; the view is only valid during the call).  Return 0 to continue.      */
typedef int32_t (*agp_newton_callback)(void* user, int32_t iteration, agp_laplace_cache* cache);

/* `laplace_f_and_lml(lfx, ys; f_init, maxiter, callback)` (Laplace.jl:140-145) after
 * _check_laplace_inputs (:167-179).  cov(fx) is either given (K, host column-major n x n) or built on
 * the device from `kernel` at the inputs X (host point-major n x D) plus `jitter` (fx.Sigma_y[1]).  */
typedef struct {
  int32_t n;
  const double* K;          /* cov(fx), Laplace.jl:174; NULL -> use (kernel, X, D, jitter)        */
  const agp_kernel* kernel;
  const double* X;
  int32_t D;
  double jitter;
  const double* y;          /* n observations as float64                                          */
  agp_likelihood lik;       /* lfx.lik                                                            */
  const double* f_init;     /* NULL -> zeros (mean(fx) of the zero-mean prior, Laplace.jl:175)    */
  int32_t maxiter;          /* >= 1 (AssertionError otherwise, Laplace.jl:257); reference default 100 */
  agp_newton_callback callback; /* may be NULL */
  void* user;
} agp_laplace_problem;

/* Outputs.  Pointers are caller-allocated host buffers or NULL (= not wanted).  The gradient is the
 * total derivative the reference obtains from Zygote: the pullback of _laplace_train_intermediates /
 * _laplace_lml at f_opt plus rrule(newton_inner_loop) (Laplace.jl:330-369).                     */
typedef struct {
  double* f_opt;            /* n                                                                   */
  double lml;               /* _laplace_lml(f_opt, cache), Laplace.jl:250-254                     */
  int32_t steps;            /* Newton steps taken                                                  */
  int32_t converged;        /* 1 if isapprox(f, fnew) stopped the loop, 0 if maxiter did           */
  double* dK;               /* d lml / d K, column-major n x n                                     */
  double* dvariance;        /* (kernel, X) form only: d lml / d kernel parameters and inputs       */
  double* dinv_lengthscale; /* n_scale                                                             */
  double* dlinear_c;
  double* dX;               /* point-major n x D                                                   */
  double* dcomp_variance;        /* n_components (AGP_KERNEL_SUM / AGP_KERNEL_PRODUCT only)         */
  double* dcomp_inv_lengthscale; /* n_components                                                    */
} agp_laplace_result;

/* Replaces newton_inner_loop / _newton_inner_loop (Laplace.jl:256-276, :304-307), the intermediates at
 * f_opt (:201-222), _laplace_lml (:250-254) and their reverse pass.  cache_out (optional) receives the
 * LaplaceCache at f_opt, kept on the device for posterior(la, lfx, ys) / prediction (:39-48, :425-463). */
int32_t agp_laplace_f_and_lml(agp_ctx* ctx, const agp_laplace_problem* problem, agp_laplace_result* result,
                              agp_laplace_cache** cache_out);
/* field: 0 = W, 1 = Wsqrt, 2 = d_loglik, 3 = a, 4 = f, 6 = fnew (callback view only) (n doubles each);
 *        5 = B_ch.L (n x n column-major); 7 = loglik (1 double, owned caches only);
 *        8 = Wsqrt, 9 = d_loglik of the last Newton step (owned caches; they differ from 1 / 2 only after a maxiter-stopped loop) */
int32_t agp_laplace_cache_fetch(agp_laplace_cache* cache, int32_t field, double* host_out);
int32_t agp_laplace_cache_destroy(agp_laplace_cache* cache);
/* Number of latent values n of a cache (or callback view): the length of its vector fields. */
int32_t agp_laplace_cache_n(agp_laplace_cache* cache);
/* Replaces laplace_f_cov(cache) -- Laplace.jl:376-386: the covariance of q(f), Wsqrt^-1 (I - B^-1) Wsqrt^-1, column-major
 * n x n, from an owned cache or from the view a Newton callback receives (laplace_steps, Laplace.jl:409-421).            */
int32_t agp_laplace_f_cov(agp_laplace_cache* cache, double* cov_out);
/* _laplace_lml(cache.f, cache) -- Laplace.jl:250-254 -- as LaplaceResult(fnew, cache) (:388-395) evaluates it per step. */
int32_t agp_laplace_cache_lml(agp_laplace_cache* cache, double* lml_out);
/* rrule(newton_inner_loop) -- Laplace.jl:330-369: newton_pullback(df_opt) = (Wsqrt .* (B_ch \ (df_opt ./ Wsqrt))) * d_loglik' on the
 * cache returned by agp_laplace_f_and_lml (the fields of the last Newton step, as in the reference).  u_out (n) receives the left
 * factor of the rank-1 cotangent, dK_out (column-major n x n) the dense matrix; either may be NULL.  The cotangents of the
 * likelihood and of ys are @not_implemented in the reference (:352-358) and are not produced here either.                  */
int32_t agp_laplace_newton_pullback(agp_laplace_cache* cache, const double* df_opt, double* u_out, double* dK_out);
/* frule(newton_inner_loop) -- Laplace.jl:309-328: fdot = (B_ch \ (Wsqrt .* (dK * d_loglik))) ./ Wsqrt for a tangent dK (host,
 * column-major n x n).                                                                                                    */
int32_t agp_laplace_newton_pushforward(agp_laplace_cache* cache, const double* dK, double* fdot_out);
/* Replaces the prediction methods of ApproxPosteriorGP{<:LaplaceApproximation} -- Laplace.jl:425-463
 * (_laplace_predict_intermediates, mean_and_var, mean_and_cov, mean, var, cov(f, x), cov(f, x, y)) on a cache
 * returned by agp_laplace_f_and_lml: kernel / Xtrain (host, point-major n x D) describe prior_at_x.  Any of
 * mean1_out (n1), var1_out (n1), cov_out (column-major n1 x n2; X2 == NULL -> y = x) may be NULL.             */
int32_t agp_laplace_predict(agp_laplace_cache* cache, const agp_kernel* kernel, const double* Xtrain, int32_t D,
                            const double* X1, int64_t n1, const double* X2, int64_t n2, double* mean1_out,
                            double* var1_out, double* cov_out);

#ifdef __cplusplus
}
#endif
#endif /* AGP_H_ */
