"""The two tcgen05 GEMM engines on the device, through their stand-alone test programs (built by __graft_entry__.build()):
tools/tf32x3_test.cu (3xTF32 split products, the Float32 mode's engine) and tools/i8emu_test.cu (INT8-slice FP64 emulation)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tool):
    exe = os.path.join(ROOT, "build", tool)
    if not os.path.exists(exe):
        pytest.skip(f"build/{tool} not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    rows = [json.loads(line) for line in out.stdout.splitlines() if line.startswith("{")]
    assert rows, out.stdout
    return rows


def test_tf32x3_engine():
    rows = _run("tf32x3_test")
    assert rows[-1]["ok"] is True and rows[-1]["worst_rel"] < 1e-5  # FP32-level products (22-bit operands, FP64-carried accumulation)


def test_i8emu_engine():
    rows = _run("i8emu_test")
    cases = [r for r in rows if "rel_to_max" in r]
    assert len(cases) >= 5
    for r in cases:
        # error against the long-double reference, relative to the largest entry of the product: FP64-level
        assert r["rel_to_max"] < 2e-13, r
    short = [r for r in rows if "rel_to_max_4slices" in r]
    assert short and all(r["rel_to_max_4slices"] < 1e-7 for r in short)  # the 28-bit configuration
    assert rows[-1]["ok"] is True
