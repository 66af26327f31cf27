"""Golden vectors for the banded global alignment + CIGAR, produced by EXECUTING THE
REFERENCE's own ksw_global2 (tools/bwa/ksw.c:502-606, oracle/_ref/libkswref.so = the unmodified ksw.c).
Run in the build container:  python tests/golden/make_golden_global.py
Pairs: a random query and a target derived from it by substitutions / insertions / deletions (what
bwa_gen_cigar2, bwa.c, hands to ksw_global2: the two sides of an alignment region), band w >= |len diff|."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import KswReference, make_params   # noqa: E402

OUT = Path(__file__).resolve().parent / "global"
# cases only the CPU emulation of the second kernel runs (tests/test_g2_emulation.py): regimes beyond the four above --
# bands far wider than the indels need (many launch classes, rows of up to 250 slots), heavy scores with a positive
# score against N, free gap opens
OUT_EMU = Path(__file__).resolve().parent / "global_emu"

CASES = {
    "global_default": dict(seed=0xB5B20401, n=700, qlen=(20, 260), err=(0.0, 0.10), wextra=(0, 30), params={}),
    "global_gaps_e2": dict(seed=0xB5B20402, n=400, qlen=(10, 200), err=(0.02, 0.15), wextra=(0, 8),
                           params=dict(o_del=4, e_del=2, o_ins=5, e_ins=1, match=2, mismatch=3)),
    "global_tiny":    dict(seed=0xB5B20403, n=300, qlen=(1, 12), err=(0.0, 0.4), wextra=(0, 3), params={}),
    "global_long":    dict(seed=0xB5B20404, n=60, qlen=(600, 1500), err=(0.01, 0.06), wextra=(5, 60), params={}),
}


CASES_EMU = {
    "wide_bands":   dict(seed=0xB5B20411, n=300, qlen=(30, 300), err=(0.0, 0.12), wextra=(0, 120), params={}),
    "heavy_scores": dict(seed=0xB5B20412, n=300, qlen=(10, 400), err=(0.02, 0.2), wextra=(0, 20),
                         params=dict(o_del=11, e_del=3, o_ins=9, e_ins=4, match=7, mismatch=13, ambig=2)),
    "free_opens":   dict(seed=0xB5B20413, n=300, qlen=(5, 150), err=(0.05, 0.3), wextra=(0, 10),
                         params=dict(o_del=0, e_del=2, o_ins=0, e_ins=1, match=1, mismatch=2, ambig=-1)),
}


def make_pair(rng, qlen, err):
    q = rng.integers(0, 4, qlen).astype(np.uint8)
    t = []
    for b in q:
        u = rng.random()
        if u < err / 3: t.append((int(b) + 1 + int(rng.integers(0, 3))) % 4)
        elif u < 2 * err / 3: t.append(int(b)); t.append(int(rng.integers(0, 4)))
        elif u < err: pass
        else: t.append(int(b))
    if not t:
        t = [int(rng.integers(0, 4))]
    t = np.array(t, dtype=np.uint8)
    if rng.random() < 0.08:
        q[int(rng.integers(0, len(q)))] = 4
    if rng.random() < 0.05:
        t[int(rng.integers(0, len(t)))] = 4
    return q, t


def main():
    OUT.mkdir(exist_ok=True); OUT_EMU.mkdir(exist_ok=True)
    K = KswReference()
    only_emu = "--emu-only" in sys.argv                  # leave the committed GPU cases as they are
    for name, spec, out_dir in [(n, s, OUT) for n, s in CASES.items() if not only_emu] + [(n, s, OUT_EMU) for n, s in CASES_EMU.items()]:
        rng = np.random.default_rng(spec["seed"])
        P = make_params(**spec["params"])
        qs, ts, ws, scores, cig, cig_n = [], [], [], [], [], []
        for _ in range(spec["n"]):
            q, t = make_pair(rng, int(rng.integers(spec["qlen"][0], spec["qlen"][1] + 1)), float(rng.uniform(*spec["err"])))
            w = abs(len(q) - len(t)) + int(rng.integers(spec["wextra"][0], spec["wextra"][1] + 1))
            w = max(w, 1)
            sc, cg = K.global_align(P, q, t, w)
            qs.append(q); ts.append(t); ws.append(w); scores.append(sc); cig.append(cg); cig_n.append(len(cg))
        np.savez_compressed(out_dir / f"{name}.npz", query=np.concatenate(qs), target=np.concatenate(ts),
                            len2=np.array([len(x) for x in qs], np.int32), len1=np.array([len(x) for x in ts], np.int32),
                            w=np.array(ws, np.int32), score=np.array(scores, np.int32), cigar=np.concatenate(cig),
                            n_cigar=np.array(cig_n, np.int32),
                            params=np.array([P.o_del, P.e_del, P.o_ins, P.e_ins, P.match, P.mismatch, P.ambig], np.int32))
        print(f"{name}: n={spec['n']} ops={sum(cig_n)} max ops/pair={max(cig_n)} score range [{min(scores)}, {max(scores)}]")


if __name__ == "__main__":
    main()
