"""Host-side mirror of the reference's public interface for the hot path.

The reference's host language (Julia) is not installed in this image, so the host side above the
C ABI is Python with the same names, argument meaning and error behaviour as
src/SparseVariationalApproximationModule.jl and src/LaplaceApproximationModule.jl of the reference
(paths relative to /root/reference), so that the parity tests read like the reference's own tests:

    f   = GP(variance * with_lengthscale(SqExponentialKernel(), l))
    fz  = f(z, jitter)                                  # FiniteGP
    q   = MvNormal(m, PDMat(Cholesky(LowerTriangular(A))))
    sva = SparseVariationalApproximation(fz, q)         # SVA.jl:93-95 (NonCentered default)
    elbo(sva, f(x, sigma2), y; num_data=N)              # SVA.jl:307-317 -> :340-360
    elbo(sva, LatentGP(f, BernoulliLikelihood(), 1e-18)(x), y; quadrature=GaussHermiteExpectation(20))
    post = posterior(sva); mean_and_var(post, xnew)     # SVA.jl:115-187, :246-253
    approx_lml(LaplaceApproximation(), lfx, y)          # Laplace.jl:58-60

Only glue lives here (argument checking, packing into the C structs, numpy views).  All
arithmetic of the path runs in libagp_b200.so on the GPU; nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import _lib as L

# ---------------------------------------------------------------------------------------------
# kernels (KernelFunctions.jl names)
# ---------------------------------------------------------------------------------------------


@dataclass(frozen=True)
class _BaseKernel:
    kind: int
    c: float = 0.0

    def __rmul__(self, variance):  # variance * kernel -> ScaledKernel
        return Kernel(self.kind, float(variance), np.ones(1), self.c)

    def __mul__(self, other):  # kernel * kernel -> KernelProduct
        if isinstance(other, (_BaseKernel, Kernel)):
            return _compose(L.KERNEL_PRODUCT, self, other)
        return self.__rmul__(other)

    def __add__(self, other):  # kernel + kernel -> KernelSum
        return _compose(L.KERNEL_SUM, self, other)


def SqExponentialKernel():
    return _BaseKernel(L.KERNEL_SE)


SEKernel = SqExponentialKernel


def Matern32Kernel():
    return _BaseKernel(L.KERNEL_MATERN32)


def Matern52Kernel():
    return _BaseKernel(L.KERNEL_MATERN52)


def LinearKernel(c: float = 0.0):
    return _BaseKernel(L.KERNEL_LINEAR, float(c))


@dataclass
class Kernel:
    """``variance * (base o ScaleTransform(s))`` or ``... o ARDTransform(v)``."""

    kind: int
    variance: float = 1.0
    inv_lengthscale: np.ndarray = field(default_factory=lambda: np.ones(1))
    c: float = 0.0
    # kind == KERNEL_SUM / KERNEL_PRODUCT (``k1 + k2`` / ``k1 * k2``): ``variance * ((c_1 (+|*) c_2 ...) o Transform(inv_lengthscale))`` with
    # components ``(kind_i, variance_i, inv_lengthscale_i)``: stationary kernels with scalar lengthscales under one shared outer transform
    components: tuple = ()

    def __post_init__(self):
        self.inv_lengthscale = np.ascontiguousarray(np.atleast_1d(self.inv_lengthscale), dtype=np.float64)
        self.components = tuple((int(q[0]), float(q[1]), float(q[2])) for q in self.components)

    def __rmul__(self, variance):
        return Kernel(self.kind, self.variance * float(variance), self.inv_lengthscale, self.c, self.components)

    def __mul__(self, other):
        if isinstance(other, (_BaseKernel, Kernel)):
            return _compose(L.KERNEL_PRODUCT, self, other)
        return self.__rmul__(other)

    def __add__(self, other):
        return _compose(L.KERNEL_SUM, self, other)


def _compose(op: int, a, b) -> "Kernel":
    """``KernelSum`` / ``KernelProduct`` of stationary kernels with scalar ScaleTransforms (each term keeps its own variance and
    lengthscale; an outer ``variance * (...)`` / ``with_lengthscale(..., l)`` / ``ARDTransform`` then applies to the whole sum or product).
    Nested sums of sums (products of products) flatten; anything else is not a device kernel (use the matrix-form Laplace interface)."""
    terms = []
    for k in (a, b):
        k = _as_kernel(k)
        if k.kind == op and k.variance == 1.0 and np.all(k.inv_lengthscale == 1.0):
            terms.extend(k.components)
        elif k.kind in (L.KERNEL_SE, L.KERNEL_MATERN32, L.KERNEL_MATERN52) and k.inv_lengthscale.size == 1:
            terms.append((k.kind, k.variance, float(k.inv_lengthscale[0])))
        else:
            raise ValueError("ArgumentError: kernel sums / products on the device take stationary kernels (SqExponential, Matern32, Matern52) with "
                             "scalar lengthscales; apply ARDTransform / variance to the whole sum or product")
    if len(terms) > L.MAX_COMPONENTS:
        raise ValueError(f"ArgumentError: at most {L.MAX_COMPONENTS} components")
    return Kernel(op, 1.0, np.ones(1), 0.0, tuple(terms))


def KernelSum(*ks) -> "Kernel":
    out = ks[0]
    for k in ks[1:]:
        out = _compose(L.KERNEL_SUM, out, k)
    return _as_kernel(out)


def KernelProduct(*ks) -> "Kernel":
    out = ks[0]
    for k in ks[1:]:
        out = _compose(L.KERNEL_PRODUCT, out, k)
    return _as_kernel(out)


def agp_kernel_struct(k: "Kernel"):
    """(ctypes agp_kernel, keep-alive objects) of a host-mirror kernel."""
    ils = np.ascontiguousarray(k.inv_lengthscale, dtype=np.float64)
    kk = L.AgpKernel(k.kind, ils.size, k.variance, L.dptr(ils), k.c, 0, None)
    keep = [ils]
    if k.components:
        arr = (L.AgpKernelComponent * len(k.components))(*[L.AgpKernelComponent(q[0], q[1], q[2]) for q in k.components])
        kk.n_components = len(k.components)
        kk.components = C.cast(arr, C.POINTER(L.AgpKernelComponent))
        keep.append(arr)
    return kk, keep


def _as_kernel(k) -> Kernel:
    if isinstance(k, Kernel):
        return k
    if isinstance(k, _BaseKernel):
        return Kernel(k.kind, 1.0, np.ones(1), k.c)
    raise TypeError(f"unsupported kernel {k!r}")


def with_lengthscale(k, lengthscale) -> Kernel:
    """``with_lengthscale(k, l) = k o ScaleTransform(1/l)`` (vector l -> ARDTransform(1 ./ l))."""
    k = _as_kernel(k)
    return Kernel(k.kind, k.variance, k.inv_lengthscale * (1.0 / np.atleast_1d(np.asarray(lengthscale, dtype=np.float64))), k.c, k.components)


def ScaleTransform(k, s) -> Kernel:
    """``k o ScaleTransform(s)``."""
    k = _as_kernel(k)
    return Kernel(k.kind, k.variance, k.inv_lengthscale * float(s), k.c, k.components)


def ARDTransform(k, v) -> Kernel:
    """``k o ARDTransform(v)``."""
    k = _as_kernel(k)
    return Kernel(k.kind, k.variance, k.inv_lengthscale * np.asarray(v, dtype=np.float64), k.c, k.components)


# ---------------------------------------------------------------------------------------------
# likelihoods / expectation methods (GPLikelihoods.jl names)
# ---------------------------------------------------------------------------------------------


@dataclass
class GaussianLikelihood:
    sigma2: float = 1e-6
    kind: int = L.LIK_GAUSSIAN


class LogisticLink:
    """``GPLikelihoods.LogisticLink()`` (the default link of ``BernoulliLikelihood``)."""


class ProbitLink:
    """``GPLikelihoods.ProbitLink()``: ``p = normcdf(f)``."""


class BernoulliLikelihood:
    """``BernoulliLikelihood(l=logistic)``: ``Bernoulli(l(f))`` with the logistic (default) or the probit link."""

    sigma2 = 0.0

    def __init__(self, link=None):
        if link is None or isinstance(link, LogisticLink) or link is LogisticLink or link == "logistic":
            self.kind = L.LIK_BERNOULLI_LOGIT
        elif isinstance(link, ProbitLink) or link is ProbitLink or link in ("probit", "normcdf"):
            self.kind = L.LIK_BERNOULLI_PROBIT
        else:
            raise ValueError(f"ArgumentError: unsupported link {link!r} (no CPU fallback)")

    def __repr__(self):
        return f"BernoulliLikelihood({'ProbitLink' if self.kind == L.LIK_BERNOULLI_PROBIT else 'LogisticLink'}())"


@dataclass
class PoissonLikelihood:
    sigma2: float = 0.0
    kind: int = L.LIK_POISSON_EXP


@dataclass
class ExponentialLikelihood:
    """``ExponentialLikelihood()`` with the exp link: ``Exponential(scale = exp(f))``."""

    sigma2: float = 0.0
    kind: int = L.LIK_EXPONENTIAL_EXP


class GammaLikelihood:
    """``GammaLikelihood(alpha)`` with the exp link: ``Gamma(alpha, scale = exp(f))``.  The shape travels in the ABI's scalar
    likelihood-parameter slot (``sigma2``); its gradient comes back in ``ELBOGradient.lik_sigma2``."""

    kind = L.LIK_GAMMA_EXP

    def __init__(self, alpha: float = 1.0):
        self.alpha = float(alpha)

    @property
    def sigma2(self) -> float:
        return self.alpha


@dataclass
class DefaultExpectationMethod:
    pass


@dataclass
class AnalyticExpectation:
    pass


@dataclass
class GaussHermiteExpectation:
    n_points: int = 20

    def nodes_weights(self):
        # FastGaussQuadrature.gausshermite(n): physicists' weight exp(-x^2)
        return np.polynomial.hermite.hermgauss(self.n_points)


@dataclass
class MonteCarloExpectation:
    """``GPLikelihoods.MonteCarloExpectation(n_samples)``: reparameterised samples ``mu + sigma * randn()``.  The normal
    variates come from a counter-based Philox stream keyed by ``seed`` (the reference uses Julia's task-local RNG, so only the
    distribution, not the stream, is shared with it)."""

    n_samples: int = 20
    seed: int = 0


# ---------------------------------------------------------------------------------------------
# GP containers (AbstractGPs.jl names)
# ---------------------------------------------------------------------------------------------


def _points(x) -> np.ndarray:
    """(N, D) C-contiguous float64: the point-major layout of ColVecs(D x N) / Vector."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    return np.ascontiguousarray(x)


class GP:
    """``GP(kernel)`` (ZeroMean) or ``GP(c, kernel)`` (ConstMean)."""

    def __init__(self, *args):
        if len(args) == 1:
            self.mean_const, k = 0.0, args[0]
        elif len(args) == 2:
            self.mean_const, k = float(args[0]), args[1]
        else:
            raise TypeError("GP(kernel) or GP(mean_const, kernel)")
        self.kernel = _as_kernel(k)

    def __call__(self, x, sigma2=1e-18):
        return FiniteGP(self, x, sigma2)


class FiniteGP:
    """``f(x, sigma2)``; ``sigma2`` may be a scalar (isotropic) or a vector (heteroscedastic)."""

    def __init__(self, f: GP, x, sigma2=1e-18):
        self.f = f
        self.x = x if isinstance(x, DeviceData) else _points(x)
        self.Sigma_y = sigma2

    def __len__(self):
        return len(self.x)


class LatentGP:
    """``LatentGP(f, lik, jitter)``."""

    def __init__(self, f: GP, lik, jitter=1e-18):
        self.f, self.lik, self.Sigma_y = f, lik, jitter

    def __call__(self, x):
        return LatentFiniteGP(FiniteGP(self.f, x, self.Sigma_y), self.lik)


@dataclass
class LatentFiniteGP:
    fx: FiniteGP
    lik: Any


class MvNormal:
    """``MvNormal(m, S)``.  ``MvNormal(m, chol_lower=A)`` is the reference's
    ``MvNormal(m, PDMat(Cholesky(LowerTriangular(A))))`` (the factor is used as given, utils.jl:18);
    a full covariance is factorised once on construction, exactly like ``PDMat(S)`` does."""

    def __init__(self, m, S=None, chol_lower=None):
        self.m = np.ascontiguousarray(m, dtype=np.float64)
        if chol_lower is not None:
            self.Lq = np.tril(np.asarray(chol_lower, dtype=np.float64))
        else:
            S = np.asarray(S, dtype=np.float64)
            self.Lq = np.linalg.cholesky(0.5 * (S + S.T))  # PDMat(S): LinearAlgebra.cholesky at construction
        if self.Lq.shape != (self.m.size, self.m.size):
            raise ValueError("DimensionMismatch: covariance / mean sizes differ")


class Centered:
    pass


class NonCentered:
    pass


class SparseVariationalApproximation:
    """``SparseVariationalApproximation([Centered()|NonCentered()], fz, q)`` -- SVA.jl:59-95."""

    def __init__(self, *args):
        if len(args) == 2:
            param, (fz, q) = NonCentered(), args  # SVA.jl:93-95
        elif len(args) == 3:
            param, fz, q = args
        else:
            raise TypeError("SparseVariationalApproximation([parametrization,] fz, q)")
        if isinstance(param, type):
            param = param()
        if not isinstance(param, (Centered, NonCentered)):
            raise TypeError("parametrization must be Centered() or NonCentered()")
        if not isinstance(fz, FiniteGP) or not isinstance(q, MvNormal):
            raise TypeError("fz must be a FiniteGP and q an MvNormal")
        if np.ndim(fz.Sigma_y) != 0:
            raise ValueError("the inducing-point jitter fz.Sigma_y must be a scalar")
        if q.m.size != len(fz.x):
            raise ValueError("DimensionMismatch: q and fz have different lengths")
        self.parametrization, self.fz, self.q = param, fz, q

    @property
    def centered(self) -> bool:
        return isinstance(self.parametrization, Centered)


def SVGP(*args):  # src/deprecations.jl:1
    return SparseVariationalApproximation(Centered(), *args)


# ---------------------------------------------------------------------------------------------
# device plumbing
# ---------------------------------------------------------------------------------------------


class Context:
    """One ``agp_ctx`` (device + stream + workspaces)."""

    def __init__(self, device: int = 0):
        self.lib = L.load_library()
        h = C.c_void_p()
        L.check(self.lib.agp_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.agp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self) -> int:
        n = C.c_int64()
        L.check(self.lib.agp_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def profile(self, enable: bool = True):
        """Switch the per-kernel-class CUDA-event timers on / off."""
        L.check(self.lib.agp_ctx_profile(self.h, 1 if enable else 0))

    def profile_read(self) -> dict:
        """{class name: (milliseconds, launch groups)} accumulated since the last read."""
        n = C.c_int32()
        ms = (C.c_double * 64)()
        cnt = (C.c_int64 * 64)()
        L.check(self.lib.agp_ctx_profile_read(self.h, 64, ms, cnt, C.byref(n)))
        return {self.lib.agp_profile_class_name(i).decode(): (ms[i], cnt[i]) for i in range(n.value)}

    def fp64_peak(self) -> dict:
        """Achieved TFLOP/s of register-resident DMMA.8x8x4 and DFMA chains on this device, measured now (agp_fp64_peak)."""
        out = {}
        v = C.c_double()
        for i, name in enumerate(("dmma", "dfma")):
            L.check(self.lib.agp_fp64_peak(self.h, i, C.byref(v)))
            out[name] = v.value
        return out

    def stream(self) -> int:
        s = C.c_void_p()
        L.check(self.lib.agp_ctx_stream(self.h, C.byref(s)))
        return s.value or 0

    # data-parallel communicator (NCCL); `unique_id` = 128 bytes from rank 0
    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        L.check(self.lib.agp_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        L.check(self.lib.agp_comm_init(self.h, nranks, rank, buf))


_default_ctx: dict[int, Context] = {}


def kernelmatrix(k, x, y=None, *, ctx: "Context | None" = None) -> np.ndarray:
    """``KernelFunctions.kernelmatrix(k, x[, y])`` = ``cov(GP(k), x[, y])`` evaluated on the device (agp_kernel_matrix)."""
    ctx = ctx or default_context()
    k = _as_kernel(k)
    x = _points(x)
    yy = None if y is None else _points(y)
    n1, n2 = len(x), (len(x) if yy is None else len(yy))
    kk, _keep = agp_kernel_struct(k)
    out = np.zeros((n1, n2), order="F")
    L.check(ctx.lib.agp_kernel_matrix(ctx.h, C.byref(kk), x.shape[1], L.dptr(x), n1, L.dptr(yy), n2, L.dptr(out)))
    return out


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


_YTYPES = {np.dtype(np.float64): L.Y_F64, np.dtype(np.float32): L.Y_F32, np.dtype(np.int64): L.Y_I64,
           np.dtype(np.uint8): L.Y_U8, np.dtype(np.bool_): L.Y_U8}


class DeviceData:
    """Device-resident ``(x, y)`` (an ``agp_dataset``): upload once, evaluate many minibatches."""

    def __init__(self, x=None, y=None, *, capacity=None, D=None, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        if x is not None:
            if not (isinstance(x, np.ndarray) and x.dtype == np.float32):  # Float32 points are uploaded as they are
                x = _points(x)
            capacity = capacity or x.shape[0]
            D = 1 if x.ndim == 1 else x.shape[1]
        h = C.c_void_p()
        L.check(self.ctx.lib.agp_dataset_create(self.ctx.h, int(capacity), int(D), C.byref(h)))
        self.h, self.D, self.N = h, int(D), 0
        if x is not None:
            self.upload(x, y)

    def upload(self, x, y=None):
        x32 = isinstance(x, np.ndarray) and x.dtype == np.float32
        x = np.ascontiguousarray(x.reshape(len(x), -1)) if x32 else _points(x)
        yp, yt = None, L.Y_F64
        if y is not None:
            y = np.ascontiguousarray(y)
            if y.dtype not in _YTYPES:
                y = y.astype(np.float64)
            yt = _YTYPES[y.dtype]
            yp = y.ctypes.data_as(C.c_void_p)
        up = self.ctx.lib.agp_dataset_upload_f32 if x32 else self.ctx.lib.agp_dataset_upload  # Float32 inputs travel as Float32
        L.check(up(self.h, x.ctypes.data_as(C.c_void_p), x.shape[0], x.shape[1], L.POINT_MAJOR, yp, yt, L.HOST))
        self.N = x.shape[0]

    def upload_device(self, x_ptr: int, n: int, y_ptr: int | None, layout=L.POINT_MAJOR, ldx=0, ytype=L.Y_F64):
        """Fill from device pointers (e.g. ``torch.Tensor.data_ptr()`` of synthetic data)."""
        L.check(self.ctx.lib.agp_dataset_upload(self.h, C.c_void_p(x_ptr), int(n), int(ldx), layout, C.c_void_p(y_ptr) if y_ptr else None, ytype, L.DEVICE))
        self.N = int(n)

    def __len__(self):
        return self.N

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.agp_dataset_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# packing
# ---------------------------------------------------------------------------------------------


def _compute_dtype(dtype) -> int:
    """None / "f64" / np.float64 -> AGP_COMPUTE_F64; "f32" / np.float32 -> AGP_COMPUTE_F32 (the Float32 fast mode); "f64emu" ->
    AGP_COMPUTE_F64_EMU (Float64 tolerance, the reverse pass's point-sum product as an FP64-accurate INT8-slice product)."""
    if dtype is None or dtype in ("f64", "float64", np.float64):
        return L.COMPUTE_F64
    if dtype in ("f32", "float32", np.float32):
        return L.COMPUTE_F32
    if dtype in ("f32_tc_solve",):
        return L.COMPUTE_F32_TC_SOLVE
    if dtype in ("f64emu", "f64_emu"):
        return L.COMPUTE_F64_EMU
    raise ValueError(f"ArgumentError: unsupported compute dtype {dtype!r}")


class _Packed:
    """Keeps the numpy buffers alive next to the ctypes struct that points into them."""

    def __init__(self, sva: SparseVariationalApproximation, lik=None, quadrature=None, dtype=None):
        k = sva.fz.f.kernel
        Z = _points(sva.fz.x)
        M, D = Z.shape
        if k.inv_lengthscale.size not in (1, D):
            raise ValueError("ARDTransform length must equal the input dimension")
        self.Z, self.m = Z, np.ascontiguousarray(sva.q.m, dtype=np.float64)
        self.Lq = np.asfortranarray(sva.q.Lq, dtype=np.float64)  # column-major
        p = L.AgpSvgpParams()
        p.kernel, self._kernel_keep = agp_kernel_struct(k)
        self.ils = self._kernel_keep[0]
        self.n_comp = len(k.components)
        p.mean_const = sva.fz.f.mean_const
        p.M, p.D = M, D
        p.Z, p.jitter = L.dptr(Z), float(sva.fz.Sigma_y)
        p.m, p.Lq, p.ldLq = L.dptr(self.m), L.dptr(self.Lq), M
        p.parametrization = L.CENTERED if sva.centered else L.NONCENTERED
        lik = lik or GaussianLikelihood(1.0)
        p.lik = L.AgpLikelihood(lik.kind, float(lik.sigma2))
        q = quadrature or DefaultExpectationMethod()
        if isinstance(q, GaussHermiteExpectation) or (isinstance(q, DefaultExpectationMethod) and lik.kind in (L.LIK_BERNOULLI_LOGIT, L.LIK_BERNOULLI_PROBIT)):
            gh = q if isinstance(q, GaussHermiteExpectation) else GaussHermiteExpectation(20)
            xs, ws = gh.nodes_weights()
            self.xs, self.ws = np.ascontiguousarray(xs), np.ascontiguousarray(ws)
            p.expect = L.AgpExpectation(L.EXPECT_GAUSS_HERMITE, len(xs), L.dptr(self.xs), L.dptr(self.ws), 0)
        elif isinstance(q, AnalyticExpectation):
            p.expect = L.AgpExpectation(L.EXPECT_ANALYTIC, 0, None, None, 0)
        elif isinstance(q, DefaultExpectationMethod):
            p.expect = L.AgpExpectation(L.EXPECT_DEFAULT, 0, None, None, 0)
        elif isinstance(q, MonteCarloExpectation):
            p.expect = L.AgpExpectation(L.EXPECT_MONTE_CARLO, int(q.n_samples), None, None, int(q.seed))
        else:
            raise ValueError(f"unsupported expectation method {q!r}")
        p.compute_dtype = _compute_dtype(dtype)
        self.p, self.M, self.D = p, M, D


PackedParams = _Packed  # public name: the agp_svgp_params struct of one (sva, likelihood, quadrature) plus the buffers it points into


@dataclass
class ELBOGradient:
    """Structural tangent of ``elbo`` (what the new ``rrule`` returns)."""

    m: np.ndarray
    Lq: np.ndarray  # lower triangular; the cotangent of the PDMat Cholesky factor
    Z: np.ndarray
    variance: float
    inv_lengthscale: np.ndarray
    linear_c: float
    mean_const: float
    lik_sigma2: float
    comp_variance: np.ndarray = field(default_factory=lambda: np.zeros(0))        # kernel sums / products: per component
    comp_inv_lengthscale: np.ndarray = field(default_factory=lambda: np.zeros(0))


def _resolve_lik(sva, l_fx):
    """FiniteGP -> GaussianLikelihood(fx.Sigma_y[1]) (SVA.jl:307-317); heteroscedastic -> error (:319-327)."""
    if isinstance(l_fx, LatentFiniteGP):
        fx, lik = l_fx.fx, l_fx.lik
    elif isinstance(l_fx, FiniteGP):
        fx = l_fx
        if np.ndim(fx.Sigma_y) != 0:
            raise RuntimeError(
                "The observation noise fx.Σy must be homoscedastic.\nTo avoid this error, construct fx using: "
                "f = GP(kernel); fx = f(x, σ²), where σ² is a positive Real."
            )
        lik = GaussianLikelihood(float(fx.Sigma_y))
    else:
        raise TypeError("expected a FiniteGP or LatentFiniteGP")
    if sva.fz.f is not fx.f:  # SVA.jl:347-351
        raise ValueError("ArgumentError: (Latent)FiniteGP prior is not consistent with SparseVariationalApproximation's")
    if not hasattr(lik, "kind"):
        raise ValueError(f"unsupported likelihood {lik!r}")
    return fx, lik


def _dataset_for(fx: FiniteGP, y, ctx: Context):
    if isinstance(fx.x, DeviceData):
        return fx.x, False
    y = np.asarray(y)
    if len(y) != len(fx.x):
        raise ValueError("DimensionMismatch: x and y have different lengths")
    return DeviceData(fx.x, y, ctx=ctx), True


def elbo(sva, l_fx, y=None, *, num_data=None, quadrature=None, ctx: Context | None = None, offset=0, count=None, global_batch=0, dtype=None) -> float:
    """``AbstractGPs.elbo(sva, fx | lfx, y; num_data, quadrature)`` -- SVA.jl:307-360.

    ``offset`` / ``count`` select a minibatch view of a device-resident data set.  With a communicator attached to ``ctx``
    (``attach_communicator``) every rank passes its own shard and ``global_batch`` = the number of points over all ranks; the
    partial sums are all-reduced once inside the library and every rank returns the same value.  ``dtype="f32"`` selects the Float32
    fast mode (what a Float32 GP is in the type-generic reference): the GEMM-shaped sweep stages run as 3xTF32 split products on
    the tcgen05 tensor cores, parity target 1e-4 against the Float64 result."""
    fx, lik = _resolve_lik(sva, l_fx)  # argument errors first: they never cross the ABI
    pk = _Packed(sva, lik, quadrature, dtype)
    ctx = ctx or default_context()
    ds, own = _dataset_for(fx, y, ctx)
    try:
        count = len(ds) - offset if count is None else count
        out = C.c_double()
        L.check(ctx.lib.agp_svgp_elbo(ctx.h, ds.h, offset, count, C.byref(pk.p), float(num_data or 0), int(global_batch), C.byref(out)))
        return out.value
    finally:
        if own:
            ds.close()


def elbo_and_gradient(sva, l_fx, y=None, *, num_data=None, quadrature=None, ctx: Context | None = None, offset=0, count=None, global_batch=0, dtype=None):
    """Value and gradient of ``elbo``: the forward + pullback of the new ``ChainRulesCore.rrule`` (same keywords as ``elbo``)."""
    fx, lik = _resolve_lik(sva, l_fx)
    pk = _Packed(sva, lik, quadrature, dtype)
    ctx = ctx or default_context()
    ds, own = _dataset_for(fx, y, ctx)
    try:
        count = len(ds) - offset if count is None else count
        M, D = pk.M, pk.D
        g = ELBOGradient(np.zeros(M), np.zeros((M, M), order="F"), np.zeros((M, D)), 0.0, np.zeros(pk.ils.size), 0.0, 0.0, 0.0)
        sc = np.zeros(4)
        G = L.AgpSvgpGrads(L.dptr(g.m), L.dptr(g.Lq), L.dptr(g.Z), sc[0:1].ctypes.data_as(L.c_double_p), L.dptr(g.inv_lengthscale),
                           sc[1:2].ctypes.data_as(L.c_double_p), sc[2:3].ctypes.data_as(L.c_double_p), sc[3:4].ctypes.data_as(L.c_double_p))
        if pk.n_comp:
            g.comp_variance, g.comp_inv_lengthscale = np.zeros(pk.n_comp), np.zeros(pk.n_comp)
            G.dcomp_variance, G.dcomp_inv_lengthscale = L.dptr(g.comp_variance), L.dptr(g.comp_inv_lengthscale)
        out = C.c_double()
        L.check(ctx.lib.agp_svgp_elbo_grad(ctx.h, ds.h, offset, count, C.byref(pk.p), float(num_data or 0), int(global_batch), C.byref(out), C.byref(G)))
        g.variance, g.linear_c, g.mean_const, g.lik_sigma2 = (float(v) for v in sc)
        return out.value, g
    finally:
        if own:
            ds.close()


class FlatELBO:
    """``flatten`` / ``unflatten`` of the trainable parameters of ``elbo(sva, l_fx, y)`` and the objective on the flat vector
    (the role of ``ParameterHandling.flatten`` + ``Optim`` in examples/b-classification/script.jl:102-142): ``x0`` is the current
    parameter vector, ``value_and_gradient(x)`` evaluates ELBO and gradient through ``agp_svgp_elbo_grad_flat`` (one pointer in,
    one out), ``unflatten(x)`` gives back a dict of named views.  Layout: include/agp.h ``agp_svgp_elbo_grad_flat``."""

    def __init__(self, sva, l_fx, y=None, *, num_data=None, quadrature=None, ctx: Context | None = None, offset=0, count=None, global_batch=0,
                 resident=True):
        """``resident=True`` evaluates through an optimiser-step handle (``agp_svgp_stepper_*``): parameter / result buffers,
        workspace and quadrature table stay on the device side between calls and a small problem (a minibatch of the
        a-regression example) costs ONE kernel launch; ``resident=False`` calls ``agp_svgp_elbo_grad_flat`` every time."""
        self.ctx = ctx or default_context()
        fx, lik = _resolve_lik(sva, l_fx)
        self._ds, self._own = _dataset_for(fx, y, self.ctx)
        self._pk = _Packed(sva, lik, quadrature)
        self._count = int(count if count is not None else len(self._ds) - offset)
        self._offset, self._num_data, self._gb = int(offset), float(num_data or 0.0), int(global_batch)
        pk = self._pk
        self.n_scale, self.M, self.D = pk.ils.size, pk.M, pk.D
        n = C.c_int64()
        L.check(self.ctx.lib.agp_svgp_flat_size(C.byref(pk.p), C.byref(n)))
        self.size = int(n.value)
        k = sva.fz.f.kernel
        self.n_comp = pk.n_comp
        self.x0 = np.concatenate([[k.variance], pk.ils, [k.c, sva.fz.f.mean_const, float(lik.sigma2)], pk.Z.ravel(), pk.m, pk.Lq.ravel(order="F"),
                                  [q[1] for q in k.components], [q[2] for q in k.components]])
        assert self.x0.size == self.size
        self._stepper = None
        if resident:
            h = C.c_void_p()
            L.check(self.ctx.lib.agp_svgp_stepper_create(self.ctx.h, C.byref(pk.p), C.byref(h)))
            self._stepper = h

    def unflatten(self, x) -> dict:
        x = np.asarray(x)
        ns, M, D = self.n_scale, self.M, self.D
        o = 4 + ns
        e = o + M * D + M + M * M
        nc = self.n_comp
        return dict(variance=x[0], inv_lengthscale=x[1:1 + ns], linear_c=x[1 + ns], mean_const=x[2 + ns], lik_param=x[3 + ns],
                    Z=x[o:o + M * D].reshape(M, D), m=x[o + M * D:o + M * D + M], Lq=x[o + M * D + M:e].reshape(M, M, order="F"),
                    comp_variance=x[e:e + nc], comp_inv_lengthscale=x[e + nc:e + 2 * nc])

    def value_and_gradient(self, x, want_grad=True, offset=None, count=None, out_grad=None):
        """ELBO and gradient at the flat vector ``x`` over points ``[offset, offset+count)`` of the resident data set (default: the
        range given at construction -- pass another one per call for minibatching, examples/a-regression/script.jl:176-194)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape == (self.size,)
        out = C.c_double()
        g = (out_grad if out_grad is not None else np.zeros(self.size)) if want_grad else None
        off = self._offset if offset is None else int(offset)
        cnt = self._count if count is None else int(count)
        if self._stepper is not None:
            L.check(self.ctx.lib.agp_svgp_stepper_eval(self._stepper, self._ds.h, off, cnt, L.dptr(x), self._num_data, self._gb, C.byref(out), L.dptr(g)))
        else:
            L.check(self.ctx.lib.agp_svgp_elbo_grad_flat(self.ctx.h, self._ds.h, off, cnt, C.byref(self._pk.p), L.dptr(x), self._num_data,
                                                         self._gb, C.byref(out), L.dptr(g)))
        return out.value, g

    def path_counts(self) -> tuple:
        """(evaluations that took the one-launch small-problem kernel, evaluations that took the throughput path)."""
        if self._stepper is None:
            return (0, 0)
        a, b = C.c_int64(), C.c_int64()
        L.check(self.ctx.lib.agp_svgp_stepper_counts(self._stepper, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def __call__(self, x):
        return self.value_and_gradient(x, want_grad=False)[0]

    def close(self):
        if getattr(self, "_stepper", None) is not None:
            self.ctx.lib.agp_svgp_stepper_destroy(self._stepper)
            self._stepper = None
        if self._own and self._ds is not None:
            self._ds.close()
        self._ds = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def approx_lml(approx, l_fx, ys=None, **kwargs):
    """``API.approx_lml`` -- SVA.jl:276-280 (alias of elbo) / Laplace.jl:58-60."""
    if isinstance(approx, SparseVariationalApproximation):
        return elbo(approx, l_fx, ys, **kwargs)
    if isinstance(approx, LaplaceApproximation):
        from .laplace_api import laplace_approx_lml

        return laplace_approx_lml(approx, l_fx, ys, **kwargs)
    raise TypeError(f"approx_lml: unsupported approximation {approx!r}")


def _prior_kl(sva, ctx: Context | None = None) -> float:
    """``_prior_kl(sva)`` -- SVA.jl:362-373."""
    ctx = ctx or default_context()
    pk = _Packed(sva)
    out = C.c_double()
    L.check(ctx.lib.agp_svgp_prior_kl(ctx.h, C.byref(pk.p), C.byref(out)))
    return out.value


class ApproxPosteriorGP:
    """``posterior(sva)`` -- SVA.jl:115-187: ``data = (Kuu = chol, B, alpha)``, computed on device."""

    def __init__(self, approx, prior: GP, ctx: Context):
        self.approx, self.prior, self.ctx = approx, prior, ctx
        self._pk = _Packed(approx)
        self._data = None

    @property
    def data(self):
        if self._data is None:
            M = self._pk.M
            Lk, B, alpha = np.zeros((M, M), order="F"), np.zeros((M, M), order="F"), np.zeros(M)
            L.check(self.ctx.lib.agp_svgp_posterior(self.ctx.h, C.byref(self._pk.p), L.dptr(Lk), L.dptr(B), L.dptr(alpha)))
            self._data = dict(Kuu_L=Lk, B=B, alpha=alpha)
        return self._data


def posterior(approx, l_fx=None, ys=None, ctx: Context | None = None):
    """``posterior(sva)`` / ``posterior(sva, fx, y)`` / ``posterior(sva, lfx, y)`` -- SVA.jl:115-201;
    ``posterior(la, lfx, ys)`` -- Laplace.jl:39-48."""
    ctx = ctx or default_context()
    if isinstance(approx, LaplaceApproximation):
        from .laplace_api import laplace_posterior

        return laplace_posterior(approx, l_fx, ys, ctx)
    if l_fx is not None:
        fx = l_fx.fx if isinstance(l_fx, LatentFiniteGP) else l_fx
        assert approx.fz.f is fx.f  # SVA.jl:192,199
    return ApproxPosteriorGP(approx, approx.fz.f, ctx)


def mean_and_var(post, x):
    """``StatsBase.mean_and_var(f_post, x)`` -- SVA.jl:246-253 / Laplace.jl:433-437."""
    from .laplace_api import LaplacePosterior

    if isinstance(post, LaplacePosterior):
        return post.mean_and_var(x)
    x = _points(x)
    mu, var = np.zeros(len(x)), np.zeros(len(x))
    L.check(post.ctx.lib.agp_svgp_mean_and_var(post.ctx.h, C.byref(post._pk.p), L.dptr(x), len(x), L.dptr(mu), L.dptr(var)))
    return mu, var


def mean_and_cov(post, x):
    """``StatsBase.mean_and_cov(f_post, x)`` -- SVA.jl:237-244 / Laplace.jl:439-443."""
    from .laplace_api import LaplacePosterior

    if isinstance(post, LaplacePosterior):
        return post.mean_and_cov(x)
    x = _points(x)
    mu, cov_ = np.zeros(len(x)), np.zeros((len(x), len(x)), order="F")
    L.check(post.ctx.lib.agp_svgp_mean_and_cov(post.ctx.h, C.byref(post._pk.p), L.dptr(x), len(x), None, 0, L.dptr(mu), L.dptr(cov_)))
    return mu, cov_


def cov(post, x, y=None):
    """``Statistics.cov(f_post, x)`` -- SVA.jl:223-228 -- and the cross-covariance ``cov(f_post, x, y)`` -- :255-264
    (Laplace.jl:453-463 for the Laplace posterior)."""
    from .laplace_api import LaplacePosterior

    if isinstance(post, LaplacePosterior):
        return post.cov(x, y)
    if y is None:
        return mean_and_cov(post, x)[1]
    x, y = _points(x), _points(y)
    out = np.zeros((len(x), len(y)), order="F")
    L.check(post.ctx.lib.agp_svgp_mean_and_cov(post.ctx.h, C.byref(post._pk.p), L.dptr(x), len(x), L.dptr(y), len(y), None, L.dptr(out)))
    return out


def marginals(post, x, jitter: float = 1e-18):
    """``marginals(f_post(x))`` (AbstractGPs; used at SVA.jl:354): the parameters ``(mu, sigma)`` of
    ``Normal.(mean, sqrt.(var .+ jitter))`` with AbstractGPs' default ``f_post(x)`` jitter of 1e-18."""
    mu, v = mean_and_var(post, x)
    if np.any(v + jitter < 0):
        from ._lib import DomainError, ERR_DOMAIN

        raise DomainError(ERR_DOMAIN, "DomainError: sqrt of a negative marginal variance")
    return mu, np.sqrt(v + jitter)


def mean(post, x):
    return mean_and_var(post, x)[0]


def var(post, x):
    return mean_and_var(post, x)[1]


def inducing_points(post: ApproxPosteriorGP):
    return post.approx.fz.x


class LaplaceApproximation:
    """``LaplaceApproximation(; newton_kwargs...)`` -- Laplace.jl:26-30 (f_init, maxiter, callback)."""

    def __init__(self, **newton_kwargs):
        self.newton_kwargs = newton_kwargs
