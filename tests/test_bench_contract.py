"""The bench line the driver parses: the committed B200 records under profiles/ carry every key of the contract, and
the reference arm (oracle port on the host cores) emits the same shape here on the CPU."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def _line(path):
    return json.loads(Path(path).read_text().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("round2_bench.json", 1), ("round2_bench_2gpu.json", 2), ("round2_bench_4gpu.json", 4),
                                    ("round2_bench_8gpu.json", 8)])
def test_committed_b200_records_follow_the_contract(name, n):
    o = _line(ROOT / "profiles" / name)
    assert BASE_KEYS <= set(o) and o["n_gpus"] == n and o["metric"] == "relaxed_proposals_per_sec"
    assert o["higher_is_better"] is True and o["scaling"] == "weak" and o["vs_baseline"] is None and o["data"] == "synthetic"
    assert "workload" in o["config"] and "model" not in o["config"]
    assert o["value"] > 0 and abs(o["ms_per_step"] * 1e-3 * o["value"] - 128 * n) < 1e-6 * 128 * n      # value = chains / step time
    e2e = o["e2e"]
    assert e2e["unit"] == o["unit"] and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert 0 < e2e["value"] and e2e["value"] != o["value"]
    assert o["gpu_launches"] > 0 and 0 < o["graph_launches"] < o["gpu_launches"]
    r = o["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] <= 1
    assert all(0 < (c.get("frac_executed_of_pipe_peak") or c.get("frac_of_hbm_peak")) <= 1 for c in r["classes"].values())
    ck = o["clocks"]
    assert ck["sm_mhz"] > 0.9 * ck["sm_max_mhz"] and not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert o["config"]["coverage"]["adsorbates_per_chain"]["mean"] > 25          # the burnt-in regime, not the pristine slab
    if n == 1:
        cb = o["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == o["unit"] and cb["sample"]
        assert o["reference_equivalent_gpu"]["value"] > cb["value"]
    for w in o["workloads"]:
        assert {"metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "dtype", "config", "e2e", "gpu_launches", "roofline"} <= set(w)
        assert w["n_gpus"] == n and w["value"] > 0 and w["e2e"]["value"] > 0 and "workload" in w["config"]
        assert w["roofline"]["frac"] is None or 0 < w["roofline"]["frac"] <= 1


def test_reference_arm_line_on_cpu():
    """`bench.py --impl reference` on the smallest workload (GaN Tersoff, 2 proposals): same keys, impl = reference, e2e = value."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "gan_tersoff", "--steps", "2",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    o = json.loads(r.stdout.strip().splitlines()[-1])
    assert BASE_KEYS <= set(o) and o["impl"] == "reference" and o["value"] > 0
    assert o["e2e"] == {"value": o["value"], "unit": o["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert o["cpu_baseline"]["kind"] == "port" and o["cpu_baseline"]["value"] == o["value"] and o["cpu_baseline"]["cores"] >= 1
