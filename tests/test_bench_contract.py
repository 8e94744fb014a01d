"""The JSON lines bench.py printed on the B200 boxes (kept under profiles/) carry every key the measurement
contract names -- a guard against breaking the line's shape while changing what is measured."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def last_line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(name + " not recorded")
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def check_common(d):
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e"):
        assert key in d, key
    assert d["metric"] == "particle-likelihoods/sec" and d["unit"] == "particle-likelihoods/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in d["e2e"], key


def test_own_arm_one_gpu():
    d = last_line("r02_bench_1gpu.json")
    check_common(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"]                      # host copies inside the timed region cost something
    assert abs(d["value"] - d["config"]["particles_total"] * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] in ("hbm", "tensor", "issue") and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in c, key
    assert c["kind"] == "port" and c["cores"] >= 1
    k = d["clocks"]
    assert k["sm_mhz"] > 0 and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm():
    d = last_line("r02_bench_reference.json")
    check_common(d)
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 == d["e2e"]["d2h_bytes_per_step"]
    own = last_line("r02_bench_1gpu.json")
    assert d["metric"] == own["metric"] and d["unit"] == own["unit"] and d["config"]["workload"] == own["config"]["workload"]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_own_arm_several_gpus(n):
    d = last_line("r02_bench_%dgpu.json" % n)
    check_common(d)
    assert d["n_gpus"] == n and d["config"]["particles_total"] == n * d["config"]["particles_per_gpu"]
    assert d["gather"]["mode"] in ("fused", "collective")
