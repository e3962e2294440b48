#!/usr/bin/env python3
"""Aggregates tools/ncu_lines.py output (all lines) of a trace kernel into code regions of bvh.cuh / motion.cuh /
wavefront_kernels.cuh: share of the executed warp instructions, share of the stall samples, and the average number of
active lanes per executed instruction (SIMT efficiency per phase).  usage: tools/ncu_regions.py <lines.txt>"""
import collections
import re
import sys

agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for ln in open(sys.argv[1]):
    m = re.match(r"(\S+):(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", ln)
    if not m:
        continue
    f, l, i, s, lanes = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))
    if f == "bvh.cuh":
        k = ("nodeStep" if 349 <= l <= 443 else "enterInstance" if 316 <= l <= 347 else "triangle test" if 97 <= l <= 151 else
             "triStep/tryHit" if 444 <= l <= 516 else "trip (votes)" if 536 <= l <= 562 else "setSpace/begin/push/pop" if 244 <= l <= 314 else "bvh other")
    elif f == "motion.cuh":
        k = "motion (SRT chain)"
    elif f == "krr_math.cuh":
        k = "krr_math (exact ops: triangle test, SRT chain, ray transform)"
    elif f == "wavefront_kernels.cuh":
        k = "stage body (refill, finalise, routing)"
    else:
        k = f
    a = agg[k]
    a[0] += i; a[1] += s; a[2] += i * lanes
print(f"{'region':64s} {'inst%':>6s} {'samp%':>6s} {'lanes':>6s}")
for k, (i, s, il) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:64s} {i:6.2f} {s:6.2f} {il / max(i, 1e-9):6.2f}")
