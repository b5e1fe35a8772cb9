"""CPU: the parts of bench.py that run without a GPU -- synthetic data generation, and the `--impl reference` arm (the CPU
restatement timed on the host cores), whose JSON line must carry the contract's keys."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_synthetic_rows_are_independent_of_the_slice():
    import bench

    w = dict(bench.WORKLOADS["c4"])
    wvec, Z, m, A = bench.make_params(w)
    Xa, ya = bench.gen_rows(w, 0, 3000, wvec)
    Xb, yb = bench.gen_rows(w, 1000, 2000, wvec)
    assert np.array_equal(Xa[1000:2000], Xb) and np.array_equal(ya[1000:2000], yb)
    lo = bench.BLOCK - 5  # a slice that straddles a generation block
    Xc, yc = bench.gen_rows(w, lo, lo + 10, wvec)
    Xd, _ = bench.gen_rows(w, lo + 5, lo + 6, wvec)
    assert np.array_equal(Xc[5:6], Xd) and Xc.shape == (10, 8) and np.all(yc >= 0)
    assert Z.shape == (1024, 8) and np.all(np.diag(A) > 0) and np.allclose(A, np.tril(A))
    assert bench.flops_per_point(1024, 8) == 6 * 1024 * 1024 + 6 * 1024 * 8
    assert abs(sum(bench.class_flops_per_point(1024, 8).values()) + bench.SAVED_BY_ALGEBRA(1024, 8) - bench.flops_per_point(1024, 8)) < 1e-6


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OPENBLAS_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and d["dtype"] == "f64" and d["data"] == "synthetic"
