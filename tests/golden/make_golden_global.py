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

CASES = {
    "global_default": dict(seed=0xB5B20401, n=700, qlen=(20, 260), err=(0.0, 0.10), wextra=(0, 30), params={}),
    "global_gaps_e2": dict(seed=0xB5B20402, n=400, qlen=(10, 200), err=(0.02, 0.15), wextra=(0, 8),
                           params=dict(o_del=4, e_del=2, o_ins=5, e_ins=1, match=2, mismatch=3)),
    "global_tiny":    dict(seed=0xB5B20403, n=300, qlen=(1, 12), err=(0.0, 0.4), wextra=(0, 3), params={}),
    "global_long":    dict(seed=0xB5B20404, n=60, qlen=(600, 1500), err=(0.01, 0.06), wextra=(5, 60), params={}),
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
    OUT.mkdir(exist_ok=True)
    K = KswReference()
    for name, spec in CASES.items():
        rng = np.random.default_rng(spec["seed"])
        P = make_params(**spec["params"])
        qs, ts, ws, scores, cig, cig_n = [], [], [], [], [], []
        for _ in range(spec["n"]):
            q, t = make_pair(rng, int(rng.integers(spec["qlen"][0], spec["qlen"][1] + 1)), float(rng.uniform(*spec["err"])))
            w = abs(len(q) - len(t)) + int(rng.integers(spec["wextra"][0], spec["wextra"][1] + 1))
            w = max(w, 1)
            sc, cg = K.global_align(P, q, t, w)
            qs.append(q); ts.append(t); ws.append(w); scores.append(sc); cig.append(cg); cig_n.append(len(cg))
        np.savez_compressed(OUT / f"{name}.npz", query=np.concatenate(qs), target=np.concatenate(ts),
                            len2=np.array([len(x) for x in qs], np.int32), len1=np.array([len(x) for x in ts], np.int32),
                            w=np.array(ws, np.int32), score=np.array(scores, np.int32), cigar=np.concatenate(cig),
                            n_cigar=np.array(cig_n, np.int32),
                            params=np.array([P.o_del, P.e_del, P.o_ins, P.e_ins, P.match, P.mismatch, P.ambig], np.int32))
        print(f"{name}: n={spec['n']} ops={sum(cig_n)} max ops/pair={max(cig_n)} score range [{min(scores)}, {max(scores)}]")


if __name__ == "__main__":
    main()
