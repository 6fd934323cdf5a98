"""The committed bench lines (profiles/r02_bench_*.json, written by bench.py on a B200) carry every key of the measurement
contract and are internally consistent.  CPU only: it reads the artefacts, it does not run the bench."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = ["r02_bench_B400.json", "r02_bench_FB_B1200.json", "r02_bench_CP_B52.json", "r02_bench_FB_B156.json"]


def _load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", LINES)
def test_own_arm_line_has_the_contract_keys(name):
    d = _load(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["unit"] == "evals/s" and d["dtype"] == "f64" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["warmup"] >= 3 or name != "r02_bench_B400.json"
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = evaluations of the whole job per second of device time
    assert d["value"] == pytest.approx(d["config"]["evals_per_step"] / (d["ms_per_step"] * 1e-3), rel=1e-9)
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0.5 * d["value"] < e["value"] <= 1.001 * d["value"]  # copies inside the timed region: never faster than resident
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-12)
    assert 0.5 < r["frac"] < 1.0
    c = d["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"]


def test_default_line_carries_cpu_baseline_and_traffic():
    d = _load("r02_bench_B400.json")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == "evals/s" and "sample" in cb
    assert len(cb["arrangements"]) == 2  # cores x 1 BLAS thread, and 1 x all-threads BLAS
    assert d["roofline"]["traffic"] is not None and d["roofline"]["traffic"] > 0
    assert sum(d["config"]["info_histogram"].values()) == 400  # how many of the 400 GPs retried with jitter


def test_reference_arm_line():
    d = _load("r02_bench_reference_arm.json")
    own = _load("r02_bench_B400.json")
    assert d["impl"] == "reference" and d["metric"] == own["metric"] and d["unit"] == own["unit"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
