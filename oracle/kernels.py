"""Kernel matrices with KernelFunctions.jl / Distances.jl semantics (oracle; test infrastructure).

The reference never computes a kernel itself: ``cov(f.prior, z, x)``
(src/SparseVariationalApproximationModule.jl:216), ``cov(fz)`` (src/utils.jl:17) and
``cov(fx)`` (src/LaplaceApproximationModule.jl:174) lower to
``KernelFunctions.kernelmatrix`` which lowers to ``Distances.pairwise``.  Those
packages are not vendored under /root/reference and carry no pinned version
(Project.toml:15), so this file restates their published algorithm
(SURVEY.md Appendix A):

* ``SqExponentialKernel``: exp(-d2/2);  ``Matern32``: (1+sqrt3 d) exp(-sqrt3 d);
  ``Matern52``: (1+sqrt5 d+5 d2/3) exp(-sqrt5 d);  ``LinearKernel(c)``: x.y + c.
* ``k o ScaleTransform(s)`` / ``ARDTransform(v)`` multiply the inputs by s / v first;
  ``with_lengthscale(k, l) = k o ScaleTransform(1/l)``; ``variance * k`` is ``ScaledKernel``.
* vector-of-vector inputs (D > 1): ``pairwise(SqEuclidean())`` uses the GEMM form
  ``max(|x|^2 + |y|^2 - 2 x.y, 0)``; ``Euclidean`` is its sqrt.  ``Vector{<:Real}`` inputs
  (D == 1) broadcast ``(x - y)^2`` directly.  The one-argument ``pairwise(d, x)`` is exactly
  symmetric with an exactly-zero diagonal.

Inputs here are row-major ``(n_points, D)`` float64 arrays.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SE, MATERN32, MATERN52, LINEAR = "se", "matern32", "matern52", "linear"
SUM, PRODUCT = "sum", "product"
KINDS = (SE, MATERN32, MATERN52, LINEAR, SUM, PRODUCT)

_SQRT3 = np.sqrt(3.0)
_SQRT5 = np.sqrt(5.0)


@dataclass
class Kernel:
    """``variance * (base o ScaleTransform(inv_lengthscale))`` (or ``ARDTransform`` when
    ``inv_lengthscale`` has D entries).  ``c`` is LinearKernel's offset."""

    kind: str = SE
    variance: float = 1.0
    inv_lengthscale: np.ndarray = field(default_factory=lambda: np.ones(1))
    c: float = 0.0
    # kind == SUM / PRODUCT (KernelFunctions ``k1 + k2`` / ``k1 * k2`` of stationary kernels under a shared outer transform):
    # ``variance * ((c_1 (+|*) c_2 ...) o Transform(inv_lengthscale))``, c_i = ``v_i * (base_i o ScaleTransform(s_i))`` given as
    # (kind_i, v_i, s_i) tuples.  One scaled squared distance u serves all components: c_i = v_i kappa_i(s_i^2 u).
    components: tuple = ()

    def __post_init__(self):
        assert self.kind in KINDS, self.kind
        self.inv_lengthscale = np.atleast_1d(np.asarray(self.inv_lengthscale, dtype=np.float64))
        if self.kind in (SUM, PRODUCT):
            assert 1 <= len(self.components) <= 4 and all(q[0] in (SE, MATERN32, MATERN52) for q in self.components)
            self.components = tuple((q[0], float(q[1]), float(q[2])) for q in self.components)

    def scale_vec(self, D: int) -> np.ndarray:
        s = self.inv_lengthscale
        if s.size == 1:
            return np.full(D, s[0])
        assert s.size == D
        return s


def _as2d(X) -> np.ndarray:
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    return X


def _u_matrix(k: Kernel, X: np.ndarray, Y: np.ndarray | None):
    """The scalar fed to kappa: squared distance of the transformed inputs (stationary
    kernels) or their dot product (linear)."""
    X = _as2d(X)
    D = X.shape[1]
    s = k.scale_vec(D)
    Xs = X * s
    sym = Y is None
    Ys = Xs if sym else _as2d(Y) * s
    if k.kind == LINEAR:
        return Xs @ Ys.T
    if D == 1:
        u = (Xs[:, 0][:, None] - Ys[:, 0][None, :]) ** 2
    else:
        xn = np.sum(Xs * Xs, axis=1)
        yn = np.sum(Ys * Ys, axis=1)
        u = np.maximum(xn[:, None] + yn[None, :] - 2.0 * (Xs @ Ys.T), 0.0)
    if sym:
        u = np.triu(u, 1)
        u = u + u.T  # exactly symmetric, exactly zero diagonal
    return u


def _kappa(kind: str, u: np.ndarray, c: float) -> np.ndarray:
    if kind == SE:
        return np.exp(-u / 2.0)
    if kind == LINEAR:
        return u + c
    d = np.sqrt(u)
    if kind == MATERN32:
        return (1.0 + _SQRT3 * d) * np.exp(-_SQRT3 * d)
    return (1.0 + _SQRT5 * d + 5.0 * u / 3.0) * np.exp(-_SQRT5 * d)


def _dkappa_du(kind: str, u: np.ndarray) -> np.ndarray:
    """d kappa / d u, finite at u == 0 (SURVEY.md Appendix A last bullet of KernelFunctions)."""
    if kind == SE:
        return -0.5 * np.exp(-u / 2.0)
    if kind == LINEAR:
        return np.ones_like(u)
    d = np.sqrt(u)
    if kind == MATERN32:
        return -1.5 * np.exp(-_SQRT3 * d)
    return -(5.0 / 6.0) * (1.0 + _SQRT5 * d) * np.exp(-_SQRT5 * d)


def _F(k: Kernel, u: np.ndarray) -> np.ndarray:
    """kappa for every kind (sums / products of stationary components included)."""
    if k.kind not in (SUM, PRODUCT):
        return _kappa(k.kind, u, k.c)
    terms = [v * _kappa(kd, s * s * u, 0.0) for kd, v, s in k.components]
    out = terms[0]
    for t in terms[1:]:
        out = out + t if k.kind == SUM else out * t
    return out


def _dF(k: Kernel, u: np.ndarray):
    """(dF/du, [dF/dv_i], [dF/ds_i]) for a sum / product kernel."""
    kap = [_kappa(kd, s * s * u, 0.0) for kd, v, s in k.components]
    dkap = [_dkappa_du(kd, s * s * u) for kd, v, s in k.components]
    n = len(k.components)
    if k.kind == SUM:
        dFdu = sum(v * s * s * dkap[i] for i, (kd, v, s) in enumerate(k.components))
        dv = [kap[i] for i in range(n)]
        ds = [v * dkap[i] * u * 2.0 * s for i, (kd, v, s) in enumerate(k.components)]
        return dFdu, dv, ds
    dFdu = 0.0
    dv, ds = [], []
    for i, (kd, v, s) in enumerate(k.components):
        oth = 1.0
        for j, (kdj, vj, sj) in enumerate(k.components):
            if j != i:
                oth = oth * (vj * kap[j])
        dFdu = dFdu + oth * v * s * s * dkap[i]
        dv.append(oth * kap[i])
        ds.append(oth * v * dkap[i] * u * 2.0 * s)
    return dFdu, dv, ds


def _F0(k: Kernel) -> float:
    if k.kind == SUM:
        return float(sum(v for _, v, _ in k.components))
    if k.kind == PRODUCT:
        return float(np.prod([v for _, v, _ in k.components]))
    return 1.0


def kernelmatrix(k: Kernel, X, Y=None) -> np.ndarray:
    """``kernelmatrix(k, x[, y])``: (len(X), len(Y))."""
    return k.variance * _F(k, _u_matrix(k, X, Y))


def kernelmatrix_diag(k: Kernel, X) -> np.ndarray:
    """``kernelmatrix_diag(k, x)`` == ``var(GP(k), x)``."""
    X = _as2d(X)
    if k.kind == LINEAR:
        Xs = X * k.scale_vec(X.shape[1])
        return k.variance * (np.sum(Xs * Xs, axis=1) + k.c)
    return np.full(X.shape[0], k.variance * _F0(k))


@dataclass
class KernelGrad:
    variance: float = 0.0
    inv_lengthscale: np.ndarray | None = None
    c: float = 0.0
    comp_variance: np.ndarray = field(default_factory=lambda: np.zeros(0))
    comp_inv_lengthscale: np.ndarray = field(default_factory=lambda: np.zeros(0))

    def add(self, o: "KernelGrad") -> "KernelGrad":
        def _a(x, y):
            return y if x.size == 0 else (x if y.size == 0 else x + y)

        return KernelGrad(self.variance + o.variance, self.inv_lengthscale + o.inv_lengthscale, self.c + o.c,
                          _a(self.comp_variance, o.comp_variance), _a(self.comp_inv_lengthscale, o.comp_inv_lengthscale))


def kernelmatrix_pullback(k: Kernel, X, Y, Kbar: np.ndarray):
    """Reverse pass of ``kernelmatrix(k, X, Y)`` for cotangent ``Kbar``.

    Returns (Xbar, Ybar, KernelGrad).  ``Y is None`` means the symmetric one-argument form;
    then ``Xbar`` already holds the contributions of both arguments and ``Ybar`` is None.
    Differentiates through u (never through d = sqrt(u)), which is the exact derivative of the
    reference's forward pass wherever the ``max(., 0)`` clamp is inactive.
    """
    X = _as2d(X)
    D = X.shape[1]
    s = k.scale_vec(D)
    sym = Y is None
    Yr = X if sym else _as2d(Y)
    u = _u_matrix(k, X, None if sym else Yr)
    kap = _F(k, u)
    g = KernelGrad(float(np.sum(Kbar * kap)), np.zeros_like(k.inv_lengthscale), 0.0)
    if k.kind in (SUM, PRODUCT):
        dFdu, dv, ds = _dF(k, u)
        g.comp_variance = np.array([float(np.sum(Kbar * k.variance * t)) for t in dv])
        g.comp_inv_lengthscale = np.array([float(np.sum(Kbar * k.variance * t)) for t in ds])
        W = Kbar * (k.variance * dFdu)
    else:
        W = Kbar * (k.variance * _dkappa_du(k.kind, u))  # cotangent of u
    if sym and k.kind != LINEAR:
        W = W.copy()
        np.fill_diagonal(W, 0.0)  # the diagonal of u is the constant 0 for stationary kernels
    rs = W.sum(axis=1)
    cs = W.sum(axis=0)
    if k.kind == LINEAR:
        g.c = float(k.variance * np.sum(Kbar))
        # u = sum_d s_d^2 x_d y_d
        Xbar = (W @ Yr) * s**2
        Ybar = (W.T @ X) * s**2
        sbar = 2.0 * s * np.einsum("ij,id,jd->d", W, X, Yr)
    else:
        # u = sum_d s_d^2 (x_d - y_d)^2
        WY = W @ Yr
        WtX = W.T @ X
        Xbar = 2.0 * s**2 * (X * rs[:, None] - WY)
        Ybar = 2.0 * s**2 * (Yr * cs[:, None] - WtX)
        sbar = 2.0 * s * (
            np.einsum("i,id->d", rs, X * X) - 2.0 * np.einsum("id,id->d", WY, X) + np.einsum("j,jd->d", cs, Yr * Yr)
        )
    if k.inv_lengthscale.size == 1:
        g.inv_lengthscale = np.array([np.sum(sbar)])
    else:
        g.inv_lengthscale = sbar
    if sym:
        return Xbar + Ybar, None, g
    return Xbar, Ybar, g


def kernelmatrix_diag_pullback(k: Kernel, X, vbar: np.ndarray):
    """Reverse pass of ``kernelmatrix_diag``: returns (Xbar, KernelGrad)."""
    X = _as2d(X)
    D = X.shape[1]
    s = k.scale_vec(D)
    g = KernelGrad(0.0, np.zeros_like(k.inv_lengthscale), 0.0)
    if k.kind == LINEAR:
        Xs2 = np.sum((X * s) ** 2, axis=1)
        g.variance = float(np.sum(vbar * (Xs2 + k.c)))
        g.c = float(k.variance * np.sum(vbar))
        sbar = 2.0 * k.variance * s * np.einsum("i,id->d", vbar, X * X)
        g.inv_lengthscale = np.array([np.sum(sbar)]) if k.inv_lengthscale.size == 1 else sbar
        Xbar = 2.0 * k.variance * (vbar[:, None] * X) * s**2
        return Xbar, g
    g.variance = float(np.sum(vbar)) * _F0(k)
    if k.kind == SUM:
        g.comp_variance = np.full(len(k.components), k.variance * float(np.sum(vbar)))
        g.comp_inv_lengthscale = np.zeros(len(k.components))
    elif k.kind == PRODUCT:
        g.comp_variance = np.array([k.variance * float(np.sum(vbar)) * _F0(k) / v for _, v, _ in k.components])
        g.comp_inv_lengthscale = np.zeros(len(k.components))
    return np.zeros_like(X), g
