"""CPU check of the bench-line contract on the committed lines under profiles/ (written by bench.py on a B200):
every key the driver reads is present and self-consistent."""
import json
import os

import pytest

PROFILES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def _line(name):
    with open(os.path.join(PROFILES, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,gpus", [("r01_bench_default.json", 1), ("r01_bench_2gpu.json", 2),
                                       ("r01_bench_4gpu.json", 4), ("r01_bench_8gpu.json", 8),
                                       ("r02_bench_default.json", 1), ("r02_bench_2gpu.json", 2),
                                       ("r02_bench_4gpu.json", 4), ("r02_bench_8gpu.json", 8)])
def test_committed_bench_lines_follow_the_contract(name, gpus):
    d = _line(name)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert key in d, key
    assert d["n_gpus"] == gpus and d["unit"] == "LP/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] == 2 * d["steps"]
    # value = units all ranks processed / max-over-ranks step time
    assert abs(d["value"] - d["config"]["global_batch"] / d["ms_per_step"] * 1e3) <= 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == "LP/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0.5 * d["value"] < e["value"] < d["value"]          # copies inside the timed region cost something
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1 and r["unit"] == "TFLOP/s"
    c = d["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert c["sm_mhz"] >= 0.9 * c["sm_max_mhz"]
    if gpus == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
        assert r["traffic"] and r["traffic"] > 0


def test_round2_arms_share_the_configuration_and_report_strong_scaling():
    """The B200 arm and the reference (CPU) arm print the same config object (workload string, batch, seed); the reference
    line carries what the driver needs to form the ratio; every B200 line also reports the fixed-global-batch side
    measurement, and the strong-scaling mode says so in `scaling`."""
    gpu, ref = _line("r02_bench_default.json"), _line("r02_bench_reference.json")
    assert ref["impl"] == "reference" and ref["config"] == gpu["config"]
    assert ref["metric"] == gpu["metric"] and ref["unit"] == gpu["unit"] and ref["higher_is_better"] is True
    assert ref["e2e"]["value"] == ref["value"] and ref["e2e"]["h2d_bytes_per_step"] == 0
    cb = ref["cpu_baseline"]
    assert cb["value"] == ref["value"] and cb["cores"] >= 1 and len(cb["serial"]["passes_s"]) >= 3
    assert cb["multiprocess"]["workers"] >= 1 and len(cb["multiprocess"]["passes_s"]) >= 3
    assert gpu["value"] > 1000 * ref["value"]
    for name, gpus in (("r02_bench_default.json", 1), ("r02_bench_2gpu.json", 2), ("r02_bench_4gpu.json", 4),
                       ("r02_bench_8gpu.json", 8)):
        side = _line(name)["strong_side"]
        assert side["global_batch"] == 1024 and side["batch_per_gpu"] * gpus == 1024
        assert abs(side["value"] - 1024 / side["ms_per_step"] * 1e3) <= 1e-6 * side["value"]
        assert side["gpu_ms_per_step"] <= side["ms_per_step"] * 1.001   # device time vs host-clocked step: the split per N
    strong = _line("r02_bench_2gpu_strong.json")
    assert strong["scaling"] == "strong" and strong["config"]["global_batch"] == 1024 and strong["n_gpus"] == 2
    tm = _line("r02_bench_default.json")
    assert tm["problem"]["tmem"] == 1 and tm["roofline"]["kernel"] == "intopt_forward_tmem_kernel"
