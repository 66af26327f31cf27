#!/usr/bin/env python
"""Throughput of bsw_extend_chains (seed -> pair construction + GPU extension) next to
the reference's own mem_chain2aln on one host thread.  Workload: the chains of the golden case
tests/golden/chain/chain_default.npz (made by the reference itself), tiled T times; results are checked
against the tiled golden regions.  Usage: python scripts/chain_bench.py [tiles=400] [reps=5]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import genomicsbench_b200 as gb
from test_chain import load_chain_case, _build_batch

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 400
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
c = load_chain_case("chain_default")
eng = gb.Engine(end_bonus=c["clip5"], **c["P"])
chains, seeds, query, ref = _build_batch(gb, eng, c)
n, ns = len(chains), len(seeds)
big_ch = np.tile(chains, tiles); big_sd = np.tile(seeds, tiles)
big_ch["seed_first"] += np.repeat(np.arange(tiles, dtype=np.int64) * ns, n)
# every tile reads the same read / window bytes (offsets unchanged): the engine's gather still moves them per pair
t_best, st = 1e9, None
out = (np.zeros(len(big_sd), dtype=gb.ALNREG_DTYPE), np.zeros(len(big_ch), dtype=np.int32))   # caller-owned, reused like a C caller's
for r in range(reps + 1):
    t0 = time.perf_counter()
    regs, count = eng.extend_chains(big_ch, big_sd, query, ref, c["w"], c["clip5"], c["clip3"], 2, out=out)
    dt = time.perf_counter() - t0
    if r:
        t_best = min(t_best, dt)
    st = eng.stats()
assert np.array_equal(count, np.tile(c["reg_n"], tiles))
got = np.concatenate([regs[int(f): int(f) + int(k)] for f, k in zip(big_ch["seed_first"][:n], count[:n])])
gm = np.stack([got[f] for f in gb.ALNREG_FIELDS], axis=1).astype(np.int64)
assert np.array_equal(gm, c["regs"]), "tile 0 differs from the reference's regions"
last = np.concatenate([regs[int(f): int(f) + int(k)] for f, k in zip(big_ch["seed_first"][-n:], count[-n:])])
assert np.array_equal(np.stack([last[f] for f in gb.ALNREG_FIELDS], axis=1).astype(np.int64), c["regs"])
line = {"metric": "chains_per_sec", "chains": int(n * tiles), "seeds": int(ns * tiles), "regions": int(count.sum()),
        "extensions": int(st["pairs"]), "seconds": t_best, "value": n * tiles / t_best,
        "extensions_per_sec": st["pairs"] / t_best, "gcups_effective": st["cells_effective"] / t_best / 1e9,
        "gpu_launches": int(st["kernel_launches"]), "buffers": "pageable numpy arrays"}
# the reference's own mem_chain2aln (oracle/_ref/libbwamemref.so), one thread, its stdout dump sent to /dev/null
try:
    import ctypes as C
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    import make_golden_chain as mg
    lib = C.CDLL(str(mg.SO)); libc = C.CDLL(None)
    lib.mem_opt_init.restype = C.POINTER(mg.MemOpt)
    lib.bwa_fill_scmat.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int8)]
    lib.mem_chain2aln.argtypes = [C.POINTER(mg.MemOpt), C.POINTER(mg.BntSeq), C.c_void_p, C.c_int, C.c_void_p,
                                  C.POINTER(mg.MemChain), C.POINTER(mg.MemAlnRegV)]
    lib.mem_chain2aln.restype = None
    libc.free.argtypes = [C.c_void_p]
    L = c["l_pac"]; G = c["D"][:L]
    pac = np.zeros(L // 4 + 1, dtype=np.uint8)
    for sh in range(4):
        part = G[sh::4]; pac[:len(part)] |= (part << ((~sh & 3) << 1)).astype(np.uint8)
    ann = mg.BntAnn(0, L, 0, 0, 0, b"chr", b""); bns = mg.BntSeq(L, 1, 11, C.pointer(ann), 0, None, None)
    opt = lib.mem_opt_init(); lib.bwa_fill_scmat(opt.contents.a, opt.contents.b, opt.contents.mat)
    jobs = []
    for k in range(n):
        f, m = int(chains["seed_first"][k]), int(chains["n_seeds"][k])
        arr = (mg.MemSeed * m)(*[mg.MemSeed(int(s["rbeg"]), int(s["qbeg"]), int(s["len"]), int(s["score"])) for s in seeds[f: f + m]])
        q = np.ascontiguousarray(query[int(chains["query_off"][k]): int(chains["query_off"][k]) + int(chains["l_query"][k])])
        jobs.append((arr, mg.MemChain(m, m, 0, 0, 0, 0.0, 0, arr), q))
    sys.stdout.flush(); saved = os.dup(1); dn = os.open(os.devnull, os.O_WRONLY); os.dup2(dn, 1)
    t0 = time.perf_counter(); done = 0
    while time.perf_counter() - t0 < 3.0:
        for arr, ch, q in jobs:
            av = mg.MemAlnRegV(0, 0, None)
            lib.mem_chain2aln(opt, C.byref(bns), pac.ctypes.data, len(q), q.ctypes.data, C.byref(ch), C.byref(av))
            if av.a: libc.free(av.a)
        done += n
    libc.fflush(None); dt = time.perf_counter() - t0
    os.dup2(saved, 1); os.close(dn)
    line["cpu_baseline"] = {"kind": "reference", "what": "mem_chain2aln (tools/bwa, unmodified), called per chain through ctypes, "
                            "its per-extension stdout dump sent to /dev/null", "cores": 1, "value": done / dt, "unit": "chains/s"}
except Exception as e:                       # oracle/_ref absent
    line["cpu_baseline"] = {"unavailable": repr(e)}
print(json.dumps(line))
eng.close()
