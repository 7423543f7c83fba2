"""CPU-side checks of bench.py: the reference arm (`--impl reference`, the oracle port timed on the host cores) and
the cpu_baseline leg run without a GPU, so their JSON contract is checked here; the GPU arm is exercised on the
B200 by the driver."""
import json
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--nx", "40", "--ny", "30", "--nobs", "1500"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "columns/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_cpu_baseline_leg_reports_scan_and_cell_grid_figures():
    import torch
    sys.path.insert(0, ROOT)
    import bench
    a = types.SimpleNamespace(nx=40, ny=30, nz=30, N=64, m=1500, corr=4000.0, maxlen=8000.0, gpus=1)
    d = bench.build_rank_data(a, 0, 1, torch.device("cpu"))
    out, _ = bench.cpu_baseline(a, d, 2.0)
    assert out["kind"] == "port" and out["value"] > 0 and out["unit"] == "columns/s"
    assert out["with_cell_grid"]["value"] > 0
    assert np.isfinite(out["value"]) and "assimilation.F90:3745-3757" in out["sample"]
