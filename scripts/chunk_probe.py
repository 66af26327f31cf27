"""e2e time of a named config on pinned buffers under the chunk-schedule knobs BSW_CHUNK / BSW_RAMPDOWN."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, genomicsbench_b200 as gb
name = sys.argv[1]
cfg = gb.gen_named_config(name); pairs, ref, qer = gb.gen_pairs(cfg, 0, 1_000_000)
with gb.Engine() as eng:
    pp, pr, pq = gb.pinned_copy(pairs), gb.pinned_copy(ref), gb.pinned_copy(qer)
    e = []
    for _ in range(6):
        t0 = time.perf_counter(); eng.extend(pp, pr, pq, 100); e.append((time.perf_counter() - t0) * 1e3)
    print(name, "BSW_CHUNK=%s BSW_RAMPDOWN=%s" % (os.environ.get("BSW_CHUNK", "-"), os.environ.get("BSW_RAMPDOWN", "-")),
          "e2e pinned ms %.3f" % min(e[1:]), "checksum", int(pp["score"].sum()))
