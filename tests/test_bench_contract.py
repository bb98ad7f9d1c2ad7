"""CPU checks of bench.py's bookkeeping: the algorithmic-byte formulas of SURVEY.md section 8(d), the committed ncu
traffic figure the roofline block cites, the workloads BASELINE.json names, and the last committed bench line."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_match_the_survey(bench):
    n, b = 1 << 20, 12771                                   # C2: N1 = N2 = Nc, 12 771 dense buckets
    assert bench.alg_bytes(n, n, b, n) == 104 * n + 36 * n + 24 * b + 40 * n == 189050184      # 189 MB
    assert bench.nn_alg_bytes(n, n, b) == 36 * n + 32 * n + 12 * b == 71456420                 # the search launch's share
    # the stage shares add up to the whole: T = 48 N1, G = 20 N1 + 12 B, N as above, A = 4 N2 + 40 Nc
    assert 48 * n + (20 * n + 12 * b) + bench.nn_alg_bytes(n, n, b) + (4 * n + 40 * n) == bench.alg_bytes(n, n, b, n)


def test_committed_ncu_traffic_is_readable_and_below_the_algorithmic_bytes(bench):
    t = bench.ncu_traffic("k_nn_search_hull", "c2")
    assert isinstance(t, int) and 30e6 < t < bench.nn_alg_bytes(1 << 20, 1 << 20, 12771)    # L2 hits: no wasted re-reads
    assert bench.ncu_traffic("k_nn_search_hull", "c1") is None                                 # captured on C2 only
    assert bench.ncu_traffic("no_such_kernel", "c2") is None


def test_workloads_are_the_baseline_configs(bench):
    cfg = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "points/sec per ICP iteration" in bench.__doc__ or "points" in cfg["metric"]
    assert set(bench.WORKLOADS) == {"c1", "c2", "c3"}
    assert bench.WORKLOADS["c2"][2] == 1.0 and bench.WORKLOADS["c1"][2] == 0.5 and bench.WORKLOADS["c3"][2] == 0.25


def test_last_committed_bench_line_carries_the_contract_keys():
    line = json.loads(open(os.path.join(ROOT, "profiles", "r4z_bench.json")).read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    r = line["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["algorithmic_bytes_per_launch"] == 71456420 and r["traffic"] is not None
    assert line["e2e"]["h2d_bytes_per_step"] == 40 * 2 * (1 << 20) and line["e2e"]["value"] < line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["gpu_launches"] == 10 * line["steps"] and not line["clocks"]["reasons"]
