#!/usr/bin/env python
"""The UNMODIFIED reference driver (main_banded.cpp linked against libbsw_b200.so: oracle/_ref/bsw_main_b200) with
its own habit -t T -b 512 (scripts/run-cpu.sh:30): pairs/s as the driver itself reports them ("Overall SW cycles"),
with the shim's call coalescing on (default) and off (BSW_SHIM_COALESCE=0).
python scripts/driver_probe.py [pairs] > profiles/<tag>_driver_probe.txt"""
import os, re, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import genomicsbench_b200 as gb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
BIN = ROOT / "oracle" / "_ref" / "bsw_main_b200"
cfg = gb.gen_named_config("small", host_only=True)
pairs, ref, qer = gb.gen_pairs(cfg, 0, n, host_only=True)
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "pairs.txt")
    gb.write_pairs_file(path, pairs, ref, qer) if False else gb.load_host_library().bsw_write_pairs_file(
        path.encode(), pairs.ctypes.data, len(pairs), ref.ctypes.data, qer.ctypes.data)
    for label, env in (("coalescing on ", {}), ("coalescing off", {"BSW_SHIM_COALESCE": "0"})):
        for T in (1, 4, 8, 16, 32):
            for b in (512,):
                res = subprocess.run([str(BIN), "-pairs", path, "-t", str(T), "-b", str(b)], capture_output=True, text=True,
                                     env=dict(os.environ, **env), timeout=600)
                m = re.search(r"Overall SW cycles = \d+, ([\d.]+) s", res.stdout)
                secs = float(m.group(1)) if m else float("nan")
                print(f"{label}  -t {T:2d} -b {b}: {secs:7.3f} s  {n / secs / 1e6 if secs > 0 else 0:7.2f} M pairs/s", flush=True)
