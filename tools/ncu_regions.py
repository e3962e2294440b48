#!/usr/bin/env python3
"""Aggregates tools/ncu_lines.py output (all lines) of a trace kernel into the traversal PHASES of bvh.cuh (located by
their function names in the current source), the SRT chain (motion.cuh), the exact-arithmetic helpers (krr_math.cuh) and
the stage body (wavefront_kernels.cuh): share of the executed warp instructions, share of the stall samples, and the average
number of active lanes per executed instruction (SIMT efficiency per phase).
usage: tools/ncu_regions.py <lines.txt>   (lines.txt = `tools/ncu_lines.py <capture> <kernel> [cubin] 3000`)"""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARKS = [("triIntersectE(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float tmax, float &t, float &u, float &v) {", "triangle test"),
         ("KRR_HD bool betterHit", "triStep / tryHit"), ("void movingRay", "enter instance (probe, park, transform)"),
         ("KRR_DEV void setSpace", "ray-space set-up, begin, push, pop"), ("KRR_DEV int probeInstance", "enter instance (probe, park, transform)"),
         ("KRR_DEV void nodeStep", "node test (nodeStep)"), ("KRR_DEV bool tryHit", "triStep / tryHit"), ("KRR_DEV void runToEnd", "bvh other"),
         ("KRR_DEV bool trip(", "trip (phase votes)")]


def bvh_regions():
    src = open(os.path.join(ROOT, "kiraray_b200", "csrc", "bvh.cuh")).read().splitlines()
    starts = []
    for i, ln in enumerate(src, 1):
        for pat, name in MARKS:
            if pat in ln:
                starts.append((i, name))
    starts.sort()
    return starts


def main():
    starts = bvh_regions()

    def bvh_region(l):
        name = "bvh other"
        for s, n in starts:
            if l >= s - 1:
                name = n
        return name

    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for ln in open(sys.argv[1]):
        m = re.match(r"(\S+):(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", ln)
        if not m:
            continue
        f, l, i, s, lanes = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))
        k = (bvh_region(l) if f == "bvh.cuh" else "SRT chain (motion.cuh)" if f == "motion.cuh" else
             "exact arithmetic (krr_math.cuh: triangle test, SRT chain, ray transform)" if f == "krr_math.cuh" else
             "stage body (refill, finalise, routing)" if f == "wavefront_kernels.cuh" else f)
        a = agg[k]
        a[0] += i; a[1] += s; a[2] += i * lanes
    print(f"{'region':76s} {'inst%':>6s} {'samp%':>6s} {'lanes':>6s}")
    for k, (i, s, il) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:76s} {i:6.2f} {s:6.2f} {il / max(i, 1e-9):6.2f}")


if __name__ == "__main__":
    main()
