"""bench.py's JSON contract, checked on CPU against the committed line of the last GPU run (profiles/r1c_bench_tc.json)
and against BASELINE.json; the helpers that need no GPU are called."""
import json
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_has_every_contract_key():
    line = json.load(open(os.path.join(ROOT, "profiles", "r1c_bench_tc.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["metric"] == bench.METRIC and line["unit"] == bench.UNIT and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"] and "l2" in line["config"]
    assert line["gpu_launches"] > 0
    e2e = line["e2e"]
    assert e2e["unit"] == bench.UNIT and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert 0 < e2e["value"] <= line["value"] * 1.02           # host buffers cost something; never faster than device-resident
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["traffic"] is None or r["traffic"] > 0
    c = line["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value is B * frames * steps / time
    assert abs(line["value"] - 32 * 128 / (line["ms_per_step"] / 1e3)) / line["value"] < 1e-6


def test_metric_is_baselines_metric():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert base["metric"].replace("→", "->").startswith(bench.METRIC)


def test_roofline_inputs_exist():
    traffic, src = bench.ncu_traffic()                          # bytes per launch from the newest committed ncu capture
    assert traffic > 1e6 and os.path.exists(os.path.join(ROOT, src))
    pk = bench.peaks()
    assert pk["tflops"] > 100 and pk["hbm"] > 1000
    # algorithmic FLOPs of config 2: 32 * (4.677 + 50 * 2 * 1.2678 + 3.899) G = 4.33 TFLOP
    alg = 32 * (bench.E_COND + bench.S_STEPS * 2 * bench.E_TRUNK + bench.E_DEC)
    assert abs(alg - 4.331e12) / 4.331e12 < 1e-3
