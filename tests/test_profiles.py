"""profiles/traffic.json feeds bench.py's roofline object (measured DRAM bytes, issue-slot utilisation, SIMT efficiency of
the stage kernels, from the committed ncu captures): every bench workload must have its dominant kernels and its per-step
totals there, or the bench line silently reports nulls."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_json_covers_every_bench_workload():
    import bench
    d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    need = {"cbox": ["k_trace_closest", "k_scatter", "k_trace_fused"], "tess20m": ["k_trace_closest", "k_scatter", "k_trace_fused"],
            "inst10k": ["k_trace_closest", "k_scatter", "k_trace_fused"], "smoke": ["k_trace_closest", "k_scatter", "k_medium_sample"]}
    assert set(need) == set(bench.WORKLOADS)
    for wl, kernels in need.items():
        step = d[wl]["__step__"]
        assert step["dram_bytes_per_step"] > 1e9 and step["launches_per_step"] >= 10 and step["kernel_ms_per_step_serialised"] > 1
        for k in kernels:
            e = bench.load_ncu(wl, k)
            assert e, (wl, k)
            assert 0 < e["issue_active_pct"] <= 100 and 0 < e["simt_efficiency"] <= 1 and e["dram_bytes_per_launch"] > 0, (wl, k, e)
            assert "issue" in e["bound"]


def test_committed_bench_lines_carry_the_contract_keys():
    for wl in ("cbox", "tess20m", "inst10k", "smoke"):
        line = json.load(open(os.path.join(ROOT, "profiles", f"r02_bench_{wl}_r2j.json")))
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                    "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "sustained"):
            assert key in line, (wl, key)
        assert line["config"]["workload"] and line["roofline"]["frac"] <= 1.0 and line["clocks"]["reasons"] == []
        assert line["e2e"]["d2h_bytes_per_step"] == line["config"]["width"] * line["config"]["height"] * 16
