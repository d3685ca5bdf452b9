"""bench.py --impl reference runs on the host cores only (the oracle port is the CPU baseline there): the
contract of its JSON line can be checked without a GPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "iq_msamples_per_s_iqbaseband_fmdemod" and d["unit"] == "Msamples/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("c2")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


import pytest


@pytest.mark.gpu
def test_bench_json_line_contract():
    """The headline arm on one GPU: one JSON line with the keys the driver reads."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "3", "--passes", "8", "--no-secondary",
                        "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "burst", "c5_bank", "timed_region_s"):
        assert k in d, k
    assert d["steps"] == 4 and d["n_gpus"] == 1 and d["dtype"] == "f32"
    assert d["config"]["passes_per_step"] == 8 and d["gpu_launches"] == 4 * 8 * 2       # accumulate + finalize per pass
    assert abs(d["value"] - d["config"]["samples_per_step_per_gpu"] / d["ms_per_step"] / 1e3) < 1e-6 * d["value"]
    c5 = d["c5_bank"]
    assert c5["c5_check"]["bit_exact"] is True and c5["c5_check"]["channels"] >= 8 and c5["timed_region_s"] >= 0.2
    assert c5["value"] > 0 and c5["n_gpus"] == 1
    ro = d["roofline"]
    assert ro["bound"] == "hbm" and ro["unit"] == "GB/s" and ro["kernel_launches"] >= 1
    assert abs(ro["frac"] - ro["achieved"] / ro["peak"]) < 1e-9 and 0.2 < ro["frac"] < 1.2
    e = d["e2e"]
    per_pass = d["config"]["buffer_size"] * d["config"]["buffers_per_pass"]
    assert e["h2d_bytes_per_step"] == per_pass * 8 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]
