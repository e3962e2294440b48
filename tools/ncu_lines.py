#!/usr/bin/env python3
"""Per-source-line view of an `ncu --set full --import-source on` capture (run here, no GPU).

  tools/ncu_lines.py <report.ncu-rep> <mangled kernel name> [cubin] [top N]

ncu's CSV export of the source page has SASS rows only, so this joins them (by instruction offset) with
`nvdisasm --print-line-info` of the cubin the library was built from (kiraray_b200/lib/obj/api.o) and
aggregates executed warp instructions and stall samples per source line."""
import collections, csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_map(cubin, func):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    m, cur, on = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            on = ln.strip() == f".text.{func}:"
            continue
        if not on:
            continue
        mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if mm:
            m[int(mm.group(1), 16)] = (cur, mm.group(2).strip())
    return m


def main():
    rep, func = sys.argv[1], sys.argv[2]
    cubin = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3].endswith(".cubin") else None
    top = int(sys.argv[-1]) if sys.argv[-1].isdigit() else 40
    if not cubin:
        os.makedirs("/tmp/cub", exist_ok=True)
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "kiraray_b200/lib/obj/api.o")], cwd="/tmp/cub", capture_output=True)
        cubin = "/tmp/cub/api.sm_100a.cubin"
    lm = line_map(cubin, func)
    # an .ncu-rep, or its `--page source --csv --print-source sass` export
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    # one section per captured launch ("Kernel Name" line, header, SASS rows): take the first launch of the wanted kernel
    want = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3].startswith("launch=") else None
    secs, cur = [], None
    for ln in txt.splitlines():
        if ln.startswith('"Kernel Name"'):
            cur = []
            secs.append(cur)
        elif cur is not None:
            cur.append(ln)
    sec = secs[int(want.split("=")[1])] if want else secs[0]
    rows = list(csv.reader(io.StringIO("\n".join(sec))))
    hdr = rows[0]
    ia, isrc, iinst, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ithr = hdr.index("Predicated-On Thread Instructions Executed")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(rows[1][ia], 16)
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), 0])
    tot_i = tot_s = mismatch = 0
    for r in rows[1:]:
        off = int(r[ia], 16) - base
        loc, text = lm.get(off, (None, ""))
        if text.split(" ")[0].split(".")[0].lstrip("@!UP0123456789 ") != r[isrc].strip().split(" ")[0].split(".")[0].lstrip("@!UP0123456789 "):
            mismatch += 1
        a = agg[loc]
        a[0] += int(r[iinst]); a[1] += int(r[isamp]); a[3] += int(r[ithr])
        for i in stalls:
            if int(r[i]): a[2][hdr[i][6:]] += int(r[i])
        tot_i += int(r[iinst]); tot_s += int(r[isamp])
    print(f"# {func}: {tot_i} warp instructions, {tot_s} samples, {len(rows) - 1} SASS rows, opcode mismatches vs cubin: {mismatch}")
    print(f"{'file:line':34s} {'inst%':>6s} {'samp%':>6s} {'lanes':>6s}  top stalls")
    for loc, (ni, ns, st, nt) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        name = f"{loc[0]}:{loc[1]}" if loc else "?"
        print(f"{name:34s} {100 * ni / tot_i:6.2f} {100 * ns / max(tot_s, 1):6.2f} {nt / max(ni, 1):6.2f}  " + " ".join(f"{k}={v}" for k, v in st.most_common(3)))


if __name__ == "__main__":
    main()
