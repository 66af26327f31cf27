#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep (needs -lineinfo + --import-source on): samples and
warp-instructions per CUDA source line of one launch.  Usage: python scripts/ncu_lines.py <rep> [launch] [min_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "File Path":
        cur = {"file": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]) and r[2] == "-":
        cur["rows"].append(r)
# blocks come per (file, launch); group by launch order per file
files = {}
for b in blocks:
    files.setdefault(b["file"], []).append(b)
tot_s = tot_i = 0
sel = []
for f, bl in files.items():
    if which < len(bl):
        b = bl[which]
        ix = {h: i for i, h in enumerate(b["hdr"])}
        for r in b["rows"]:
            s, n = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
            if s or n:
                sel.append((f.split("/")[-1], int(r[0]), s, n, r[ix["Avg. Threads Executed"]], r[1].strip()))
                tot_s += s; tot_i += n
print(f"total samples {tot_s}, warp-instructions {tot_i}")
for f, ln, s, n, thr, src in sel:
    if 100.0 * s / max(tot_s, 1) >= minpct or 100.0 * n / max(tot_i, 1) >= minpct:
        print(f"{f[:18]:18s} {ln:4d} {100.0 * s / tot_s:5.1f}% smp {100.0 * n / tot_i:5.1f}% ins thr {thr:>5s}  {src[:100]}")
