#!/usr/bin/env python3
"""Summarise Nsight Compute output for profiles/ (run here, no GPU needed).

  tools/ncu_summary.py launches gpurun_out/launches.csv        -> per-kernel launch counts / time shares
  tools/ncu_summary.py report   gpurun_out/prof.ncu-rep        -> key metrics of every captured launch
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)


def traffic(out, *reps):
    """profiles/traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over
    the captured launches) of each kernel in the given `ncu --set full` reports; bench.py reads it."""
    import json
    import re
    res = {}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = re.sub(r"^(void )?(krr::)?", "", d["Kernel Name"]).split("<")[0].split("(")[0]
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(d[k].replace(",", "")) * mult[units[hdr.index(k)]]
            e = res.setdefault(name, {"launches": 0, "bytes": 0.0, "source": rep.split("/")[-1]})
            e["launches"] += 1
            e["bytes"] += tot
    out_d = {k: {"dram_bytes_per_launch": v["bytes"] / v["launches"], "launches_captured": v["launches"], "source": v["source"],
                 "note": "first launches of the stage (depth 0/1) of the bench workload rendered as one band (full-frame launches, as in bench.py's profiling pass); ncu replays run cold-cache"} for k, v in res.items()}
    json.dump(out_d, open(out, "w"), indent=1)
    print(json.dumps(out_d, indent=1))


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name", "gpu__time_duration.sum") != "gpu__time_duration.sum":
            continue  # (lists taken with the DRAM byte counters as well: tools/ncu_summary.py counters reads those)
        name = row["Kernel Name"].split("(")[0]
        t = to_us(row["Metric Value"], row["Metric Unit"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {total / 1e3:.3f} ms of kernel time (cold-cache, serialised)")
    print(f"{'kernel':60s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:60]:60s} {n:8d} {t / 1e3:10.3f} {t / n:10.1f} {100 * t / total:6.1f}%")


def report(path):
    out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## {d.get('Kernel Name')}  (launch id {d.get('ID')})")
        for k in KEYS:
            if k in d:
                print(f"  {k:85s} {d[k]:>18s} {units[hdr.index(k)]}")
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        print()


def short_name(full):
    import re
    return re.sub(r"^(void )?(krr::)?", "", full).split("<")[0].split("(")[0]


def counters(out, *pairs):
    """profiles/traffic.json, per workload: `tools/ncu_summary.py counters profiles/traffic.json cbox=a.ncu-rep tess20m=b.ncu-rep
    inst10k@step=launches.csv ...`.  <workload>=<report>: per kernel of an `ncu --set full` capture (mean over the
    captured launches) DRAM bytes per launch, DRAM % of peak, issue-slot utilisation, warps active, SIMT efficiency
    (active threads per executed instruction / 32), L2 hit rate, registers; <workload>@step=<csv>: a launch list with
    dram__bytes_read.sum, dram__bytes_write.sum and gpu__time_duration.sum of EVERY launch of a short bench run ->
    DRAM bytes per step (all path kernels, divided by the number of render calls = k_begin_frame launches).
    bench.py reads the file (roofline.traffic / issue_active_pct / ... / dram_bytes_per_step)."""
    import json
    import os
    res = json.load(open(out)) if os.path.exists(out) else {}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    want = {"issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
            "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio", "dram_pct_of_peak": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l2_hit_pct": "lts__t_sector_hit_rate.pct", "registers": "launch__registers_per_thread",
            "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "stall_no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"}
    for pair in pairs:
        key, path = pair.split("=", 1)
        if key.endswith("@step"):
            wl = key[:-5]
            lines = [l for l in open(path) if not l.startswith("==")]
            by_id = collections.OrderedDict()
            for row in csv.DictReader(lines):
                e = by_id.setdefault(row["ID"], {"name": short_name(row["Kernel Name"])})
                v = float(row["Metric Value"].replace(",", ""))
                if row["Metric Name"].startswith("dram__bytes"):
                    e["bytes"] = e.get("bytes", 0.0) + v * mult.get(row["Metric Unit"], 1)
                elif row["Metric Name"] == "gpu__time_duration.sum":
                    e["us"] = to_us(row["Metric Value"], row["Metric Unit"])
            path_kernels = [e for e in by_id.values() if e["name"].startswith(("k_trace", "k_scatter", "k_generate", "k_resolve", "k_film", "k_handle", "k_begin_frame",
                                                                                 "k_fold", "k_medium", "k_tail", "k_sort", "k_ray_keys"))]
            renders = max(1, sum(1 for e in path_kernels if e["name"] == "k_begin_frame"))
            res.setdefault(wl, {})["__step__"] = {"dram_bytes_per_step": sum(e.get("bytes", 0.0) for e in path_kernels) / renders,
                                                  "kernel_ms_per_step_serialised": sum(e.get("us", 0.0) for e in path_kernels) / renders / 1e3,
                                                  "launches_per_step": len(path_kernels) / renders, "renders_captured": renders, "source": os.path.basename(path)}
            continue
        # an .ncu-rep, or its `--page raw --csv` export (tools/round2.sh exports on the GPU box: the reports are too large to bring back)
        txt = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        acc = {}
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = short_name(d["Kernel Name"])
            a = acc.setdefault(name, {"n": 0, "bytes": 0.0, "us": 0.0, **{k: 0.0 for k in want}})
            a["n"] += 1
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                a["bytes"] += float(d[k].replace(",", "")) * mult[units[hdr.index(k)]]
            a["us"] += to_us(d["gpu__time_duration.sum"], units[hdr.index("gpu__time_duration.sum")])
            for k, m in want.items():
                a[k] += float(d[m].replace(",", ""))
        for name, a in acc.items():
            n = a["n"]
            e = {"dram_bytes_per_launch": a["bytes"] / n, "launch_us_under_ncu": a["us"] / n, "launches_captured": n, "source": os.path.basename(path)}
            for k in want:
                e[k] = round(a[k] / n, 3)
            e["simt_efficiency"] = round(e.pop("threads_per_inst") / 32.0, 3)
            e["bound"] = ("issue/latency: %.0f %% of the issue slots busy, %.0f %% of the lanes active per instruction, DRAM at %.1f %% of its peak"
                          % (e["issue_active_pct"], 100 * e["simt_efficiency"], e["dram_pct_of_peak"]))
            res.setdefault(key, {})[name] = e
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: (list(v) if isinstance(v, dict) and k in ("cbox", "tess20m", "inst10k", "smoke") else "...") for k, v in res.items()}))


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic, "counters": counters}[sys.argv[1]](*sys.argv[2:])
