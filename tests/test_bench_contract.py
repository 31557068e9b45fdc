"""bench.py's contract on a box without a GPU: the reference arm prints one JSON line with the agreed keys and never
loads the product library; the B200 arm refuses to run (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_arm_prints_the_contract_line_and_never_loads_the_product():
    env = dict(os.environ, PYTHONPATH=ROOT)
    # LD_DEBUG=files lists every shared object the process maps: libem2b200.so must not be among them
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=dict(env, LD_DEBUG="files"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "libem2b200" not in r.stderr and "ExpressionMatrix2.cpython" not in r.stderr
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "cell-pairs/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("cell-pairs/sec") and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"] == "c1" and line["config"]["cells"] == 10_000
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["cpu_baseline"]["value"] == line["value"] and "pair loop" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == dict(value=line["value"], unit="cell-pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None


def test_reference_arm_does_nothing_on_other_ranks():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
