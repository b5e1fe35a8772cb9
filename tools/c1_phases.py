#!/usr/bin/env python
"""Phase time stamps of the one-launch small-problem kernel (library built with -DAGP_SMALL_TIMING; AGP_B200_LIB points at it)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import agp_b200 as agp  # noqa: E402
from agp_b200 import _lib as L  # noqa: E402
from bench import c1_problem  # noqa: E402

NAMES = ["params", "Kuu", "chol+inv", "whiten", "tile: xs/Kuf/A/C", "tile: per-point", "tile: Ab/Kb", "tile: G/g", "tile: kgrad sums", "KL", "data dZ/theta",
         "P1/Asum/V/Bbar", "dLq/Lbar/Phi/Kuubar", "Kuu kgrad", "outputs"]
for M in (20, 50):
    x, y, z = c1_problem(M)
    ctx = agp.default_context()
    f = agp.GP(1.3 * agp.with_lengthscale(agp.SqExponentialKernel(), 0.3))
    sva = agp.SparseVariationalApproximation(f(z, 1e-5), agp.MvNormal(np.zeros(M), chol_lower=np.eye(M)))
    ds = agp.DeviceData(x, y, ctx=ctx)
    fo = agp.FlatELBO(sva, agp.FiniteGP(f, ds, 0.3), None, num_data=1e4, ctx=ctx)
    for _ in range(5):
        fo.value_and_gradient(fo.x0, offset=0, count=100)
    t = np.zeros(16)
    L.check(ctx.lib.agp_svgp_stepper_phase_ticks(fo._stepper, L.dptr(t)))
    d = np.diff(t) / 1e3
    print(f"M={M}: kernel {(t[15] - t[0]) / 1e3:.1f} us  " + "  ".join(f"{n}={v:.1f}" for n, v in zip(NAMES, d)))
