#!/usr/bin/env python
"""Summarises an .ncu-rep here (no GPU needed): headline metrics per launch, stall reasons,
opcode mix, and the hottest source lines.  Usage: python scripts/ncu_summary.py <rep> [launch-index]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max']
for k in keys:
    for i, h in enumerate(hdr):
        if h == k:
            print(f"{k} [{units[i]}]: " + " | ".join(r[i][:70] for r in rows[2:]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
b = blocks[which]
ix = {h: i for i, h in enumerate(b["hdr"])}
stalls = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); mix = collections.Counter(); samp = inst = 0
for r in b["rows"]:
    samp += int(r[ix["# Samples"]]); n = int(r[ix["Instructions Executed"]]); inst += n
    for s in stalls:
        tot[s] += int(r[ix[s]])
    op = r[ix["Source"]].strip().split()
    if op[0].startswith("@"):
        op = op[1:]
    mix[op[0].rstrip(";")] += n
print(f"\nlaunch {which}: {b['name'][:80]}\nsamples {samp}  warp-instructions {inst}")
print("stalls: " + ", ".join(f"{s[6:]} {100 * v / samp:.1f}%" for s, v in tot.most_common(8)))
print("opcode mix: " + ", ".join(f"{k} {100 * v / inst:.1f}%" for k, v in mix.most_common(22)))
if "--sass" in sys.argv:
    for r in b["rows"]:
        print(f"{int(r[ix['# Samples']]):6d} {int(r[ix['Instructions Executed']]):10d} {r[ix['Avg. Threads Executed']]:>5s}  {r[ix['Source']].strip()[:90]}")
