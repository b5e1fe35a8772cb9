import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Write the measured GPU-vs-oracle errors collected by tests/_cases.record_parity (GPU sessions only)."""
    cases = sys.modules.get("_cases")
    rec = getattr(cases, "PARITY_ERRORS", None)
    if not rec:
        return
    import json

    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    worst = {}
    for errs in rec.values():
        for k, v in errs.items():
            if isinstance(v, float) and k != "tol" and v == v:
                worst[k] = max(worst.get(k, 0.0), v)
    with open(os.path.join(out, "parity_errors.json"), "w") as fh:
        json.dump({"what": "max-abs error of the CUDA path against the NumPy oracle, relative to the max-abs entry of the oracle's array "
                           "(scalars: relative); every case is one GPU test through the C ABI", "worst_per_quantity": worst, "cases": rec}, fh, indent=1)
