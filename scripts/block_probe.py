"""Kernel-only time of the named configs with the packed kernel's block size forced (BSW_SHORT_BLOCK) or chosen per class."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, genomicsbench_b200 as gb
name = sys.argv[1]
cfg = gb.gen_named_config(name); pairs, ref, qer = gb.gen_pairs(cfg, 0, 1_000_000)
with gb.Engine() as eng:
    a = pairs.copy(); eng.stage(a, ref, qer, 100)
    ts = []
    for _ in range(6):
        eng.run_staged(); ts.append(eng.stats()["ms_kernel"])
    eng.fetch(a)
    print(name, "BSW_SHORT_BLOCK=%s" % os.environ.get("BSW_SHORT_BLOCK", "auto"), "kernel ms %.3f" % min(ts[1:]), "checksum", int(a["score"].sum()))
