"""ctypes binding of ``libagp_b200.so`` (the C ABI of include/agp.h).

This is the exact analogue of the Julia ``ccall`` shim shown in INTEGRATION.md: plain pointers
and sizes, no torch types.  There is no CPU fallback -- if the shared library (or a CUDA device)
is missing every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGP_B200_LIB") or os.path.join(HERE, "libagp_b200.so")  # same override the Julia shim reads

# status codes (include/agp.h)
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_NOT_PD, ERR_DOMAIN, ERR_CUDA, ERR_NCCL, ERR_ALLOC = range(8)

KERNEL_SE, KERNEL_MATERN32, KERNEL_MATERN52, KERNEL_LINEAR, KERNEL_SUM, KERNEL_PRODUCT = range(6)
MAX_COMPONENTS = 4
LIK_GAUSSIAN, LIK_BERNOULLI_LOGIT, LIK_POISSON_EXP, LIK_EXPONENTIAL_EXP, LIK_GAMMA_EXP, LIK_BERNOULLI_PROBIT = range(6)
EXPECT_DEFAULT, EXPECT_ANALYTIC, EXPECT_GAUSS_HERMITE, EXPECT_MONTE_CARLO = range(4)
NONCENTERED, CENTERED = 0, 1
POINT_MAJOR, FEATURE_MAJOR = 0, 1
Y_F64, Y_F32, Y_I64, Y_U8 = range(4)
HOST, DEVICE = 0, 1
COMPUTE_F64, COMPUTE_F32, COMPUTE_F32_TC_SOLVE, COMPUTE_F64_EMU = 0, 1, 2, 3

c_double_p = C.POINTER(C.c_double)


class AgpKernelComponent(C.Structure):
    _fields_ = [("kind", C.c_int32), ("variance", C.c_double), ("inv_lengthscale", C.c_double)]


class AgpKernel(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n_scale", C.c_int32),
        ("variance", C.c_double),
        ("inv_lengthscale", c_double_p),
        ("linear_c", C.c_double),
        ("n_components", C.c_int32),
        ("components", C.POINTER(AgpKernelComponent)),
    ]


class AgpLikelihood(C.Structure):
    _fields_ = [("kind", C.c_int32), ("sigma2", C.c_double)]


class AgpExpectation(C.Structure):
    _fields_ = [("method", C.c_int32), ("n_points", C.c_int32), ("nodes", c_double_p), ("weights", c_double_p), ("seed", C.c_uint64)]


class AgpSvgpParams(C.Structure):
    _fields_ = [
        ("kernel", AgpKernel),
        ("mean_const", C.c_double),
        ("M", C.c_int32),
        ("D", C.c_int32),
        ("Z", c_double_p),
        ("jitter", C.c_double),
        ("m", c_double_p),
        ("Lq", c_double_p),
        ("ldLq", C.c_int32),
        ("parametrization", C.c_int32),
        ("lik", AgpLikelihood),
        ("expect", AgpExpectation),
        ("compute_dtype", C.c_int32),
    ]


class AgpSvgpGrads(C.Structure):
    _fields_ = [
        ("dm", c_double_p),
        ("dLq", c_double_p),
        ("dZ", c_double_p),
        ("dvariance", c_double_p),
        ("dinv_lengthscale", c_double_p),
        ("dlinear_c", c_double_p),
        ("dmean_const", c_double_p),
        ("dlik_sigma2", c_double_p),
        ("dcomp_variance", c_double_p),
        ("dcomp_inv_lengthscale", c_double_p),
    ]


NEWTON_CALLBACK = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.c_void_p)


class AgpLaplaceProblem(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("K", c_double_p),
        ("kernel", C.POINTER(AgpKernel)),
        ("X", c_double_p),
        ("D", C.c_int32),
        ("jitter", C.c_double),
        ("y", c_double_p),
        ("lik", AgpLikelihood),
        ("f_init", c_double_p),
        ("maxiter", C.c_int32),
        ("callback", NEWTON_CALLBACK),
        ("user", C.c_void_p),
    ]


class AgpLaplaceResult(C.Structure):
    _fields_ = [
        ("f_opt", c_double_p),
        ("lml", C.c_double),
        ("steps", C.c_int32),
        ("converged", C.c_int32),
        ("dK", c_double_p),
        ("dvariance", c_double_p),
        ("dinv_lengthscale", c_double_p),
        ("dlinear_c", c_double_p),
        ("dX", c_double_p),
        ("dcomp_variance", c_double_p),
        ("dcomp_inv_lengthscale", c_double_p),
    ]


# every symbol include/agp.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    "agp_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(_vp)]),
    "agp_ctx_destroy": (C.c_int32, [_vp]),
    "agp_last_error_string": (C.c_char_p, []),
    "agp_last_error_info": (C.c_int32, []),
    "agp_build_arch": (C.c_int32, []),
    "agp_ctx_stream": (C.c_int32, [_vp, C.POINTER(_vp)]),
    "agp_ctx_launch_count": (C.c_int32, [_vp, C.POINTER(C.c_int64)]),
    "agp_ctx_profile": (C.c_int32, [_vp, C.c_int32]),
    "agp_ctx_profile_read": (C.c_int32, [_vp, C.c_int32, c_double_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "agp_profile_class_name": (C.c_char_p, [C.c_int32]),
    "agp_comm_unique_id": (C.c_int32, [_vp]),
    "agp_comm_init": (C.c_int32, [_vp, C.c_int32, C.c_int32, _vp]),
    "agp_comm_destroy": (C.c_int32, [_vp]),
    "agp_dataset_create": (C.c_int32, [_vp, C.c_int64, C.c_int32, C.POINTER(_vp)]),
    "agp_dataset_upload": (C.c_int32, [_vp, _vp, C.c_int64, C.c_int64, C.c_int32, _vp, C.c_int32, C.c_int32]),
    "agp_dataset_upload_f32": (C.c_int32, [_vp, _vp, C.c_int64, C.c_int64, C.c_int32, _vp, C.c_int32, C.c_int32]),
    "agp_dataset_size": (C.c_int32, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "agp_dataset_destroy": (C.c_int32, [_vp]),
    "agp_svgp_elbo_grad": (
        C.c_int32,
        [_vp, _vp, C.c_int64, C.c_int64, C.POINTER(AgpSvgpParams), C.c_double, C.c_int64, c_double_p, C.POINTER(AgpSvgpGrads)],
    ),
    "agp_svgp_elbo": (C.c_int32, [_vp, _vp, C.c_int64, C.c_int64, C.POINTER(AgpSvgpParams), C.c_double, C.c_int64, c_double_p]),
    "agp_svgp_flat_size": (C.c_int32, [C.POINTER(AgpSvgpParams), C.POINTER(C.c_int64)]),
    "agp_svgp_elbo_grad_flat": (
        C.c_int32,
        [_vp, _vp, C.c_int64, C.c_int64, C.POINTER(AgpSvgpParams), c_double_p, C.c_double, C.c_int64, c_double_p, c_double_p],
    ),
    "agp_svgp_stepper_create": (C.c_int32, [_vp, C.POINTER(AgpSvgpParams), C.POINTER(_vp)]),
    "agp_svgp_stepper_flat_size": (C.c_int32, [_vp, C.POINTER(C.c_int64)]),
    "agp_svgp_stepper_eval": (C.c_int32, [_vp, _vp, C.c_int64, C.c_int64, c_double_p, C.c_double, C.c_int64, c_double_p, c_double_p]),
    "agp_svgp_stepper_counts": (C.c_int32, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "agp_svgp_stepper_phase_ticks": (C.c_int32, [_vp, c_double_p]),
    "agp_svgp_stepper_destroy": (C.c_int32, [_vp]),
    "agp_svgp_sweep": (C.c_int32, [_vp, _vp, C.c_int64, C.c_int64, C.POINTER(AgpSvgpParams), C.c_double, C.c_int64, C.c_int32]),
    "agp_svgp_reduce_buffer": (C.c_int32, [_vp, C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "agp_svgp_finish": (C.c_int32, [_vp, c_double_p, C.POINTER(AgpSvgpGrads)]),
    "agp_svgp_prior_kl": (C.c_int32, [_vp, C.POINTER(AgpSvgpParams), c_double_p]),
    "agp_svgp_posterior": (C.c_int32, [_vp, C.POINTER(AgpSvgpParams), c_double_p, c_double_p, c_double_p]),
    "agp_svgp_mean_and_var": (C.c_int32, [_vp, C.POINTER(AgpSvgpParams), c_double_p, C.c_int64, c_double_p, c_double_p]),
    "agp_svgp_mean_and_cov": (C.c_int32, [_vp, C.POINTER(AgpSvgpParams), c_double_p, C.c_int64, c_double_p, C.c_int64, c_double_p, c_double_p]),
    "agp_kernel_matrix": (C.c_int32, [_vp, C.POINTER(AgpKernel), C.c_int32, c_double_p, C.c_int64, c_double_p, C.c_int64, c_double_p]),
    "agp_fp64_peak": (C.c_int32, [_vp, C.c_int32, c_double_p]),
    "agp_laplace_predict": (
        C.c_int32,
        [_vp, C.POINTER(AgpKernel), c_double_p, C.c_int32, c_double_p, C.c_int64, c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p],
    ),
    "agp_laplace_f_and_lml": (C.c_int32, [_vp, C.POINTER(AgpLaplaceProblem), C.POINTER(AgpLaplaceResult), C.POINTER(_vp)]),
    "agp_laplace_cache_fetch": (C.c_int32, [_vp, C.c_int32, c_double_p]),
    "agp_laplace_cache_destroy": (C.c_int32, [_vp]),
    "agp_laplace_cache_n": (C.c_int32, [_vp]),
    "agp_laplace_f_cov": (C.c_int32, [_vp, c_double_p]),
    "agp_laplace_newton_pullback": (C.c_int32, [_vp, c_double_p, c_double_p, c_double_p]),
    "agp_laplace_newton_pushforward": (C.c_int32, [_vp, c_double_p, c_double_p]),
    "agp_laplace_cache_lml": (C.c_int32, [_vp, c_double_p]),
}


class AgpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"agp error {code}: {msg}")
        self.code = code


class PosDefException(AgpError):
    """cholesky(Kuu) / cholesky(B) failed (LinearAlgebra.PosDefException in the reference)."""


class DomainError(AgpError):
    """sqrt of a negative W / variance (Base.DomainError in the reference)."""


_lib = None


def load_library(path: str | None = None):
    """dlopen the C-ABI library and attach prototypes.  Needs neither torch nor a GPU."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status == OK:
        return
    msg = load_library().agp_last_error_string().decode("utf-8", "replace")
    if status == ERR_NOT_PD:
        e = PosDefException(status, msg)
        e.info = int(load_library().agp_last_error_info())  # LinearAlgebra.PosDefException(info): the failing column
        raise e
    if status == ERR_DOMAIN:
        raise DomainError(status, msg)
    if status in (ERR_INVALID, ERR_UNSUPPORTED):
        raise ValueError(f"agp error {status}: {msg}")  # ArgumentError in the Julia shim
    raise AgpError(status, msg)


def dptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.float64 and (a.flags.c_contiguous or a.flags.f_contiguous), "float64, contiguous"
    return a.ctypes.data_as(c_double_p)
