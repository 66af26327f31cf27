#!/usr/bin/env python
"""profiles/<tag>_ncu_short16_kernel_<workload>.txt (scripts/ncu_summary.py output of one step's DP
launches under `ncu --set full`) -> profiles/dram_traffic.json, which bench.py reports as
roofline.traffic.  Usage: python scripts/dram_traffic.py <tag>"""
import json, re, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out_path = ROOT / "profiles" / "dram_traffic.json"
out = json.loads(out_path.read_text()) if out_path.exists() else {}
for wl, n in (("short8", 1_000_000), ("long16", 1_000_000)):
    f = ROOT / "profiles" / f"{tag}_ncu_short16_kernel_{wl}.txt"
    if not f.exists():
        continue
    tot, launches = 0.0, 0
    for line in f.read_text().splitlines():
        m = re.match(r"dram__bytes_(read|write)\.sum \[(\w+)\]: (.*)", line)
        if m:
            vals = [float(x) for x in m.group(3).split("|")]
            tot += sum(vals) * UNIT[m.group(2)]
            launches = len(vals)
    out[wl] = {**out.get(wl, {}), "dram_bytes_per_step": tot, "launches": launches, "pairs_per_gpu": n,
               "source": f"profiles/{f.name}"}        # (keeps the hand-added warm-L2 / gathered figures and the note)
out_path.write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
