"""The driver-facing JSON line of bench.py: the committed line of the round's measurement run
(profiles/r2/bench_default.json, written by `python bench.py` on a B200) carries every key of the
contract with consistent values.  A static check -- the line itself is produced on the GPU box."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", "r2", name)) as f:
        lines = [l for l in f.read().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "bench.py prints exactly one JSON line on stdout"
    return json.loads(lines[0])


def test_headline_line_has_the_contract_keys():
    d = _line("bench_default.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "workloads"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "brick10_guadalupe_twirl" and "model" not in d["config"]
    assert d["steps"] >= 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= 1.05 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["achieved"] <= r["peak"] and r["traffic"] is not None and "dm_sweep_tma_kernel" in r["kernel"]
    # achieved = algorithmic bytes of the sweep launches / their CUDA-event time, and the sweeps are ~99 % of the step
    assert abs(r["bytes_per_launch"] * r["launches_per_step"] / (d["ms_per_step"] * 1e-3 * r["sweep_share_of_step"]) / 1e9 - r["achieved"]) < 0.03 * r["achieved"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["max_abs_diff_vs_gpu"] <= 1e-10
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    w = d["workloads"]
    assert set(w) == {"tfim4_lima_zne", "tfim14_dm", "tfim30_sv"}
    for name, v in w.items():
        assert "error" not in v, (name, v.get("error"))
        assert v["value"] > 0 and v["e2e"] > 0 and v["roofline"]["frac"] > 0 and v["max_abs_diff_vs_cpu"] <= 1e-10, name
    assert w["tfim14_dm"]["roofline"]["frac"] >= 0.70 and r["frac"] >= 0.70


def test_reference_arm_line():
    d = _line("bench_reference.json")
    assert d["impl"] == "reference" and d["metric"] and d["unit"] == "circuits/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
