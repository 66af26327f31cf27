"""Golden vectors for the seed -> pair construction / chain extension, produced by
EXECUTING THE REFERENCE's own mem_chain2aln (tools/bwa/bwamem.c:632-808) from
oracle/_ref/libbwamemref.so -- the unmodified tools/bwa sources compiled where they lie
(oracle/Makefile) -- on synthetic genomes, reads and seed chains.  Run in the build container:
    python tests/golden/make_golden_chain.py
Inputs of a case: a random genome G (the reference fetches windows of the doubled sequence
D = G + revcomp(G) through bns_fetch_seq, bntseq.c:427-452), reads cut from D with substitutions and
indels, and per read one chain of seeds = the exact-match runs of its true alignment plus a few
contained / off-diagonal extras that drive the containment test (bwamem.c:664-700).
Outputs: every mem_alnreg_t the reference pushed, in order (rb re qb qe score truesc w seedcov seedlen0).
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
OUT = Path(__file__).resolve().parent / "chain"
SO = ROOT / "oracle" / "_ref" / "libbwamemref.so"


class MemOpt(C.Structure):                      # mem_opt_t, bwamem.h:32-69
    _fields_ = [(n, C.c_int) for n in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "pen_unpaired", "pen_clip5",
                                       "pen_clip3", "w", "zdrop")] + \
               [("max_mem_intv", C.c_uint64)] + \
               [(n, C.c_int) for n in ("T", "flag", "min_seed_len", "min_chain_weight", "max_chain_extend")] + \
               [("split_factor", C.c_float)] + \
               [(n, C.c_int) for n in ("split_width", "max_occ", "max_chain_gap", "n_threads", "chunk_size")] + \
               [(n, C.c_float) for n in ("mask_level", "drop_ratio", "XA_drop_ratio", "mask_level_redun", "mapQ_coef_len")] + \
               [(n, C.c_int) for n in ("mapQ_coef_fac", "max_ins", "max_matesw", "max_XA_hits", "max_XA_hits_alt")] + \
               [("mat", C.c_int8 * 25)]


class BntAnn(C.Structure):                      # bntann1_t, bntseq.h:41-48
    _fields_ = [("offset", C.c_int64), ("len", C.c_int32), ("n_ambs", C.c_int32), ("gi", C.c_uint32),
                ("is_alt", C.c_int32), ("name", C.c_char_p), ("anno", C.c_char_p)]


class BntSeq(C.Structure):                      # bntseq_t, bntseq.h:56-64
    _fields_ = [("l_pac", C.c_int64), ("n_seqs", C.c_int32), ("seed", C.c_uint32), ("anns", C.POINTER(BntAnn)),
                ("n_holes", C.c_int32), ("ambs", C.c_void_p), ("fp_pac", C.c_void_p)]


class MemSeed(C.Structure):                     # mem_seed_t, bwamem.c:168-172
    _fields_ = [("rbeg", C.c_int64), ("qbeg", C.c_int32), ("len", C.c_int32), ("score", C.c_int)]


class MemChain(C.Structure):                    # mem_chain_t, bwamem.c:174-180
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("first", C.c_int), ("rid", C.c_int), ("bits", C.c_uint32),
                ("frac_rep", C.c_float), ("pos", C.c_int64), ("seeds", C.POINTER(MemSeed))]


class MemAlnReg(C.Structure):                   # mem_alnreg_t, bwamem.h:71-91
    _fields_ = [("rb", C.c_int64), ("re", C.c_int64)] + \
               [(n, C.c_int) for n in ("qb", "qe", "rid", "score", "truesc", "sub", "alt_sc", "csub", "sub_n", "w",
                                       "seedcov", "secondary", "secondary_all", "seedlen0", "bits")] + \
               [("frac_rep", C.c_float), ("hash", C.c_uint64)]


class MemAlnRegV(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(MemAlnReg))]


REG_FIELDS = ("rb", "re", "qb", "qe", "score", "truesc", "w", "seedcov", "seedlen0")

# name -> dict(seed, genome length, reads, error rate, overrides of mem_opt_t)
CASES = {
    "chain_default":  dict(seed=0xB5B20301, L=60_000, reads=500, err=(0.02, 0.06), opt={}),
    "chain_w8_div":   dict(seed=0xB5B20302, L=40_000, reads=350, err=(0.08, 0.12), opt=dict(w=8)),
    "chain_gaps_e2":  dict(seed=0xB5B20303, L=40_000, reads=250, err=(0.03, 0.08),
                           opt=dict(o_del=4, e_del=2, o_ins=5, e_ins=1, zdrop=60)),
    # several chains per read pushing into ONE mem_alnreg_v, as mem_align1_core does (bwamem.c:1105-1112): the
    # containment test of a later chain sees the regions of the earlier ones
    "chain_multi":    dict(seed=0xB5B20304, L=40_000, reads=400, err=(0.03, 0.09), opt={}, split=True),
}


def make_reads(rng, D, L, n_reads, err):
    """-> list of (read codes, seeds[(rbeg, qbeg, len)])."""
    out = []
    for k in range(n_reads):
        rl = int(rng.integers(70, 251))
        strand = int(rng.integers(0, 2))
        edge = rng.random()
        lo, hi = strand * L, (strand + 1) * L
        if edge < 0.08:
            start = lo + int(rng.integers(0, 40))                 # at the strand's first bases
        elif edge < 0.16:
            start = hi - rl - int(rng.integers(0, 40))            # at its last bases (next to the strand boundary)
        else:
            start = int(rng.integers(lo, hi - rl - 40))
        r = float(rng.uniform(*err))
        read, runs = [], []
        run_q = run_r = run_len = 0
        pos = start
        while len(read) < rl and pos < hi:
            u = rng.random()
            if u < r / 3:                                          # substitution
                if run_len: runs.append((run_r, run_q, run_len)); run_len = 0
                read.append(int((D[pos] + 1 + rng.integers(0, 3)) % 4)); pos += 1
            elif u < 2 * r / 3:                                    # insertion into the read
                if run_len: runs.append((run_r, run_q, run_len)); run_len = 0
                read.append(int(rng.integers(0, 4)))
            elif u < r:                                            # deletion from the read
                if run_len: runs.append((run_r, run_q, run_len)); run_len = 0
                pos += 1
            else:
                if not run_len: run_q, run_r = len(read), pos
                read.append(int(D[pos])); pos += 1; run_len += 1
        if run_len: runs.append((run_r, run_q, run_len))
        read = np.array(read[:rl], dtype=np.uint8)
        runs = [(rb, qb, min(ln, len(read) - qb)) for rb, qb, ln in runs if qb < len(read)]
        seeds = [s for s in runs if s[2] >= 19]
        if not seeds:
            best = max(runs, key=lambda s: s[2])
            if best[2] < 8:
                continue
            seeds = [best]
        extra = []
        for rb, qb, ln in seeds:
            v = rng.random()
            if ln >= 30 and v < 0.25:                              # contained, same diagonal
                off = int(rng.integers(1, ln - 20))
                extra.append((rb + off, qb + off, int(rng.integers(15, ln - off + 1))))
            elif ln >= 24 and v < 0.40:                            # overlapping, shifted diagonal
                sh = int(rng.choice([-3, -2, -1, 1, 2, 3]))
                if lo <= rb + sh and rb + sh + ln <= hi:
                    extra.append((rb + sh, qb, ln - int(rng.integers(0, 3))))
        seeds = sorted(set(seeds + extra), key=lambda s: (s[1], s[0]))
        if rng.random() < 0.1:                                     # a few N in the read
            for p in rng.integers(0, len(read), size=int(rng.integers(1, 4))):
                read[p] = 4
        out.append((read, seeds))
    return out


def main():
    OUT.mkdir(exist_ok=True)
    lib = C.CDLL(str(SO))
    libc = C.CDLL(None)
    lib.mem_opt_init.restype = C.POINTER(MemOpt)
    lib.bwa_fill_scmat.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int8)]
    lib.mem_chain2aln.argtypes = [C.POINTER(MemOpt), C.POINTER(BntSeq), C.c_void_p, C.c_int, C.c_void_p,
                                  C.POINTER(MemChain), C.POINTER(MemAlnRegV)]
    lib.mem_chain2aln.restype = None
    libc.free.argtypes = [C.c_void_p]
    # mem_chain2aln dumps every extension to stdout (the benchmark's input format, bwamem.c:741-745): drop it
    sys.stdout.flush()
    saved = os.dup(1)
    for name, spec in CASES.items():
        rng = np.random.default_rng(spec["seed"])
        L = spec["L"]
        G = rng.integers(0, 4, size=L, dtype=np.uint8)
        D = np.concatenate([G, (3 - G[::-1]).astype(np.uint8)])
        pac = np.zeros(L // 4 + 1, dtype=np.uint8)
        for sh in range(4):                                        # _set_pac, bntseq.c
            part = G[sh::4]
            pac[:len(part)] |= (part << ((~sh & 3) << 1)).astype(np.uint8)
        ann = BntAnn(0, L, 0, 0, 0, b"chr", b"")
        bns = BntSeq(L, 1, 11, C.pointer(ann), 0, None, None)
        opt = lib.mem_opt_init()
        for k, v in spec["opt"].items():
            setattr(opt.contents, k, v)
        lib.bwa_fill_scmat(opt.contents.a, opt.contents.b, opt.contents.mat)
        reads = make_reads(rng, D, L, spec["reads"], spec["err"])
        q_all, q_off, l_query, seed_rows, chain_first, chain_n, reg_rows, reg_n, chain_read = [], [], [], [], [], [], [], [], []
        qpos = 0
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            for rid, (read, seeds) in enumerate(reads):
                q = np.ascontiguousarray(read)
                parts = [seeds]
                if spec.get("split") and len(seeds) >= 2:
                    cut = rng.integers(0, 3)
                    parts = [seeds[0::2], seeds[1::2]] if cut == 0 else [seeds[:len(seeds) // 2], seeds[len(seeds) // 2:]] \
                        if cut == 1 else [seeds[len(seeds) // 2:], seeds[:len(seeds) // 2]]
                av = MemAlnRegV(0, 0, None)
                for part in parts:
                    arr = (MemSeed * len(part))(*[MemSeed(rb, qb, ln, ln * opt.contents.a) for rb, qb, ln in part])
                    ch = MemChain(len(part), len(part), 0, 0, 0, 0.0, 0, arr)
                    before = av.n
                    lib.mem_chain2aln(opt, C.byref(bns), pac.ctypes.data, len(q), q.ctypes.data, C.byref(ch), C.byref(av))
                    chain_first.append(len(seed_rows)); chain_n.append(len(part)); chain_read.append(rid)
                    seed_rows += [(rb, qb, ln, ln * opt.contents.a) for rb, qb, ln in part]
                    reg_n.append(av.n - before)
                    for i in range(before, av.n):
                        reg_rows.append([getattr(av.a[i], f) for f in REG_FIELDS])
                    q_off.append(qpos); l_query.append(len(q))
                if av.a:
                    libc.free(av.a)
                q_all.append(q); qpos += len(q)
            libc.fflush(None)
        finally:
            os.dup2(saved, 1)
            os.close(devnull)
        o = opt.contents
        params = np.array([o.a, o.b, o.o_del, o.e_del, o.o_ins, o.e_ins, o.zdrop, o.pen_clip5, o.pen_clip3, o.w], dtype=np.int32)
        regs = np.array(reg_rows, dtype=np.int64).reshape(-1, len(REG_FIELDS))
        np.savez_compressed(OUT / f"{name}.npz", genome=G, query=np.concatenate(q_all), query_off=np.array(q_off, np.int64),
                            l_query=np.array(l_query, np.int32), seeds=np.array(seed_rows, np.int64),
                            chain_first=np.array(chain_first, np.int64), chain_n=np.array(chain_n, np.int32),
                            regs=regs, reg_n=np.array(reg_n, np.int32), params=params,
                            chain_read=np.array(chain_read, np.int32))
        retried = int((regs[:, 6] > o.w).sum())
        to_end = int(((regs[:, 2] == 0)).sum())
        print(f"{name}: reads={len(reads)} chains={len(chain_n)} seeds={len(seed_rows)} regs={len(regs)} "
              f"skipped seeds={len(seed_rows) - len(regs)} band-retried regs={retried} qb==0 regs={to_end}")
        libc.free(opt)


if __name__ == "__main__":
    main()
