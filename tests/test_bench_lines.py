"""The bench contract, checked on the JSON lines the final runs of the round left in profiles/ (no GPU needed): every line
carries the keys the driver reads, the roofline fraction is achieved / peak of the stated bound, parity against the oracle
was green at every GPU count, and the whole-job rate is consistent with ms_per_step."""
import json
from pathlib import Path

import pytest

PROFILES = Path(__file__).resolve().parents[1] / "profiles"
LINES = sorted(PROFILES.glob("r02_scale_n*.json")) + sorted(PROFILES.glob("r02_bench_c*.json"))


@pytest.mark.parametrize("path", LINES, ids=[p.name for p in LINES])
def test_committed_bench_line_follows_the_contract(path):
    d = json.loads(path.read_text().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "roofline", "e2e", "gpu_launches", "clocks", "parity"):
        assert key in d, key
    assert d["unit"] == "it/s" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                      # BASELINE.md holds no published number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    # value = iterations of all timed steps / their time
    its = d["steps"] * d["config"]["iters_per_step"]
    assert abs(d["value"] - its / (d["steps"] * d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["ms_per_launch"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"]
    assert "traffic_source" in r
    if r["traffic"] is not None:  # (the N = 1 line of c3 was taken minutes before the round's ncu capture was filed: null there)
        assert 0.9 <= r["traffic"] / r["algorithmic_bytes_per_launch"] <= 1.1   # no wasted re-reads
    if d.get("cpu_baseline"):     # rank 0 at N = 1, unless --no-cpu-baseline
        assert d["n_gpus"] == 1 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= 1.02 * d["value"]
    assert d["parity"]["ok"] is True
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not (bad & set(d["clocks"]["reasons"]))
    if d["n_gpus"] > 1:
        assert d["scaling"] == "strong" and d["config"]["comm_error"] == 0
        assert len(d["config"]["spmv_ms_per_launch_by_rank"]) == d["n_gpus"]
