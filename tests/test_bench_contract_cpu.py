"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints exactly ONE JSON line on stdout
with the keys the driver reads; everything else (library banners, progress) goes to stderr."""
import json
import os
import subprocess
import sys

import pytest

from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not (pyoracle.have_ref() or pyoracle.have_oracle()), reason="CPU checker not built")
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "runmicro cell-hours/sec" and d["unit"] == "cell-hours/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("runmicro_big synthetic 8192x8192")


def test_native_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert not r.stdout.strip()  # nothing that could be mistaken for a result


def test_committed_bench_line_has_the_contract_keys():
    """The bench line committed under profiles/ (written by `python bench.py` on a B200) carries every key of the contract."""
    path = os.path.join(ROOT, "profiles", "r01_bench_final_with_pageable.json")
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert k in d, k
    assert d["metric"] == "runmicro cell-hours/sec" and d["dtype"] == "f64" and d["scaling"] == "weak" and d["warmup"] >= 3
    assert d["gpu_launches"] > 0 and d["config"]["workload"].startswith("runmicro_big synthetic 8192x8192")
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"] > 0
    assert r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.7          # the stash no longer doubles the DRAM traffic
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    full = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_final.json")))
    assert full["cpu_baseline"]["kind"] == "reference" and full["cpu_baseline"]["cores"] >= 1
    ref = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_final_reference_arm.json")))
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["config"] == d["config"]
