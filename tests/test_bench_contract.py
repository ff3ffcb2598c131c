"""The reference arm of bench.py (`--impl reference`) runs on the CPU: check here that it prints ONE
JSON line with the keys the driver reads, and that the non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra):
    env = dict(os.environ, DFB_BENCH_REF_SAMPLE="64", **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--gpus", env_extra.get("WORLD_SIZE", "1")],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    return [ln for ln in r.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_json_line():
    lines = run({})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rk_steps_per_sec" and d["unit"] == "RK steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "RK steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


@pytest.mark.gpu
def test_native_arm_json_line_has_every_contract_key():
    env = dict(os.environ, DFB_BENCH_SKIP_DDE="1", DFB_BENCH_SKIP_OTHERS="1", DFB_BENCH_CPU_SECONDS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3",
                        "--members", "65536"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "strong"
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and "workload" in d["config"]
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert d["e2e"]["h2d_bytes_per_step"] == 65536 * 6 * 8 and d["e2e"]["d2h_bytes_per_step"] == 65536 * 4 * 8
    assert 0 < d["e2e"]["value"] <= d["value"] * 1.001
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert 0.0 < d["roofline"]["frac"] <= 1.0 and d["roofline"]["unit"] == "TFLOP/s"
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["gpu_launches"] == 2 and set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_dde_leg_report_combines_ranks_and_survives_a_failed_rank():
    # the collective part of the secondary (DDE) leg, on one rank: the arithmetic of the report and the
    # error path (a failed leg must yield an error entry, never an exception: the headline still prints)
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    n = bench.DDE_MEMBERS
    K = 3
    local = {"total_ms": 18.5 * K, "e2e_s": 0.020 * K, "err": 1e-11, "steps": 5001 * n, "n_failed": 0,
             "h2d": n * 40, "d2h": n * 16, "phases": {}, "n_local": n}
    rep = bench.dde_report(local, None, n, K, "strong")
    assert rep["n_gpus"] == 1 and rep["rk_steps_per_member"] == 5001 and rep["n_failed"] == 0
    assert rep["members_total"] == n and rep["scaling"] == "strong"
    assert abs(rep["value"] - 5001 * n / 18.5e-3) < 1.0
    rf = rep["roofline"]
    assert rf["bound"] == "hbm" and abs(rf["achieved"] - bench.DDE_BYTES_PER_STEP * rep["value"] / 1e9) < 1e-6
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert abs(rep["e2e"]["value"] - 5001 * n / 0.020) < 1.0 and rep["e2e"]["h2d_bytes_per_step"] == n * 40
    assert bench.dde_report({"error": "boom"}, None, n, K, "strong") == {"error": "boom"}


def test_shards_of_the_strong_scaling_run_tile_the_baseline_ensemble():
    sys.path.insert(0, ROOT)
    import importlib
    import numpy as np
    bench = importlib.import_module("bench")
    total = 1000
    y_all, p_all = bench.lorenz_inputs(total, 0, total)
    for world in (1, 2, 3, 8):
        rows = []
        for r in range(world):
            lo, hi = bench.shard_bounds(total, r, world)
            rows.append(bench.lorenz_inputs(total, lo, hi)[1])
        assert np.array_equal(np.concatenate(rows), p_all)
    # both arms print the same config dict (the driver compares them)
    assert bench.make_config("strong", 8) == bench.make_config("strong", 8, "fast")
    assert bench.make_config("strong", 8)["members_total"] == bench.N_MEMBERS
    assert bench.make_config("weak", 8)["members_total"] == 8 * bench.N_MEMBERS
