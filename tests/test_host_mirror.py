"""CPU tests of the host mirror's kernel composition rules (no device needed)."""
import numpy as np
import pytest


def test_kernel_sum_product_composition_rules():
    """`k1 + k2` / `k1 * k2` of the host mirror: components keep their own variance and scalar lengthscale, outer variance / transform apply to
    the whole sum or product, nested sums flatten, unsupported compositions raise (ArgumentError)."""
    import agp_b200 as agp
    from agp_b200 import _lib as L

    k = 2.0 * agp.with_lengthscale(0.5 * agp.with_lengthscale(agp.SqExponentialKernel(), 2.0) + agp.Matern32Kernel() + 3.0 * agp.Matern52Kernel(), [1.0, 4.0])
    assert k.kind == L.KERNEL_SUM and k.variance == 2.0 and np.allclose(k.inv_lengthscale, [1.0, 0.25])
    assert k.components == ((L.KERNEL_SE, 0.5, 0.5), (L.KERNEL_MATERN32, 1.0, 1.0), (L.KERNEL_MATERN52, 3.0, 1.0))
    kp = agp.KernelProduct(agp.SqExponentialKernel(), agp.with_lengthscale(agp.Matern52Kernel(), 0.5))
    assert kp.kind == L.KERNEL_PRODUCT and kp.components == ((L.KERNEL_SE, 1.0, 1.0), (L.KERNEL_MATERN52, 1.0, 2.0))
    kk, keep = agp.agp_kernel_struct(k)
    assert kk.n_components == 3 and kk.components[2].variance == 3.0 and kk.n_scale == 2
    with pytest.raises(ValueError):
        agp.LinearKernel() + agp.SqExponentialKernel()
    with pytest.raises(ValueError):
        agp.with_lengthscale(agp.SqExponentialKernel(), [1.0, 2.0]) + agp.SqExponentialKernel()
    with pytest.raises(ValueError):
        (agp.SqExponentialKernel() + agp.Matern32Kernel()) * agp.SqExponentialKernel()
