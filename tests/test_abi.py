"""CPU: the C-ABI shared library loads (no GPU needed) and exports every symbol include/agp.h declares;
the ctypes prototypes cover the same set; the product path fails loudly without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "agp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^(?:int32_t|const char\*)\s+(agp_\w+)\s*\(", src, flags=re.M)
    assert len(names) >= 25
    return names


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    import agp_b200

    return agp_b200.load_library()


def test_every_declared_symbol_is_exported(lib):
    import agp_b200

    declared = _declared_symbols()
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/agp.h but not exported"
    assert set(declared) == set(agp_b200.SYMBOLS), set(declared) ^ set(agp_b200.SYMBOLS)
    assert lib.agp_build_arch() == 100


def test_library_is_self_contained(lib):
    """The product .so must not link torch (Julia hosts dlopen it directly); NCCL is dlopen'ed lazily."""
    import subprocess

    import agp_b200

    out = subprocess.run(["ldd", agp_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libc10" not in out and "nccl" not in out, out


def test_sm100a_code_is_embedded(lib):
    import subprocess

    import agp_b200

    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", agp_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry point is an error (never a silent CPU path)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import agp_b200 as agp

    with pytest.raises(agp.AgpError) as e:
        agp.Context(0)
    assert "no CPU fallback" in str(e.value)
    f = agp.GP(agp.SqExponentialKernel())
    sva = agp.SparseVariationalApproximation(f(np.zeros((3, 1)) + np.arange(3)[:, None], 1e-6), agp.MvNormal(np.zeros(3), chol_lower=np.eye(3)))
    with pytest.raises(agp.AgpError):
        agp.elbo(sva, f(np.arange(5.0), 0.1), np.zeros(5))


def test_argument_errors_never_cross_the_abi(lib):
    """SVA.jl:347-351 (ArgumentError), :319-327 (ErrorException) are raised by the host mirror before any device call."""
    import agp_b200 as agp

    f = agp.GP(agp.SqExponentialKernel())
    other = agp.GP(agp.SqExponentialKernel())
    sva = agp.SparseVariationalApproximation(f(np.arange(3.0), 1e-6), agp.MvNormal(np.zeros(3), chol_lower=np.eye(3)))
    with pytest.raises(ValueError, match="ArgumentError"):
        agp.elbo(sva, other(np.arange(5.0), 0.1), np.zeros(5))
    with pytest.raises(RuntimeError, match="homoscedastic"):
        agp.elbo(sva, f(np.arange(5.0), np.full(5, 0.1)), np.zeros(5))
    with pytest.raises(ValueError):
        agp.SparseVariationalApproximation(f(np.arange(3.0), 1e-6), agp.MvNormal(np.zeros(4), chol_lower=np.eye(4)))
    assert agp.SVGP(f(np.arange(3.0), 1e-6), agp.MvNormal(np.zeros(3), chol_lower=np.eye(3))).centered  # src/deprecations.jl:1
    assert not sva.centered  # NonCentered is the default, SVA.jl:93-95
