"""GPU: the bench.py contract - both arms print ONE JSON line with the keys the driver reads, the same `config`, and
plausible values (a short run; the numbers themselves are the driver's business)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def run(*args, timeout=600):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_both_arms_print_the_contract_line():
    ours = run("--gpus", "1", "--steps", "3", "--warmup", "3", "--headline-only", "--no-cpu-baseline")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in ours, k
    assert ours["unit"] == "draft tokens/s" and ours["n_gpus"] == 1 and ours["steps"] == 3 and ours["dtype"] == "bf16"
    assert ours["higher_is_better"] is True and ours["scaling"] == "weak" and ours["vs_baseline"] is None
    assert 1e5 < ours["value"] < 1e6 and 1e5 < ours["e2e"]["value"] <= ours["value"] * 1.05
    assert ours["e2e"]["h2d_bytes_per_step"] == 64 * 8 + 64 * 4096 * 2 and ours["e2e"]["d2h_bytes_per_step"] == 64 * 3 * 8
    assert ours["gpu_launches"] == 3 * 10                      # rms + 3 x (gate, cell, head) per step
    rf = ours["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and 0.5 < rf["frac"] < 1.1
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["traffic"] is None or 0.9 < rf["traffic"] / rf["algorithmic_bytes"] < 1.2     # no wasted re-reads
    assert "workload" in ours["config"] and "model" in ours["config"]
    assert set(ours["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert 0.9 < ours["kv_write"]["frac_of_hbm_peak"] < 1.1 and ours["kv_write"]["rope_fused"]["frac_of_hbm_peak"] > 0.7

    ref = run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1")
    assert ref["impl"] == "reference" and ref["config"] == ours["config"]
    assert ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["e2e"] == {"value": ref["value"], "unit": ref["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert ref["cpu_baseline"]["kind"] in ("reference", "port") and ref["cpu_baseline"]["cores"] >= 1
    assert 10 < ref["value"] < ours["value"]
