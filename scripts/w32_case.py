"""The sweep pairs at w = 32 on the resident path (the packed kernel's circular rows), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genomicsbench_b200 as gb
pairs, ref, qer = gb.gen_pairs(gb.gen_named_config("sweep"), 0, 1_000_000)
with gb.Engine() as eng:
    eng.stage(pairs, ref, qer, 32)
    for _ in range(3):
        eng.run_staged()
    print(eng.stats()["ms_kernel"], eng.stats()["kernel_launches"])
