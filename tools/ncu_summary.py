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
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
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


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
