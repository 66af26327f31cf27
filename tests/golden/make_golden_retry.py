"""Golden vectors for the band-doubling retry (tools/bwa/bwamem.c:630,723-753,770-800), produced by
EXECUTING THE REFERENCE's own ksw_extend2 (tools/bwa/ksw.c:380-479, oracle/_ref/libkswref.so) inside
the loop exactly as mem_chain2aln writes it.  Run in the build container:
    python tests/golden/make_golden_retry.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import genomicsbench_b200 as gb                         # noqa: E402
from oracle.pyoracle import KswReference, make_params   # noqa: E402
from make_golden import make_cfg                        # noqa: E402

OUT = Path(__file__).resolve().parent / "retry"

# name -> (generator spec, n, w, max_try, prev rule)
CASES = {
    "left_w8":    (dict(named="sweep", seed=0xB5B20201), 400, 8, 2, "left"),     # prev = -1 (bwamem.c:707)
    "right_w8":   (dict(named="sweep", seed=0xB5B20202), 400, 8, 2, "right"),    # prev = h0 (bwamem.c:767)
    "left_w4_t3": (dict(named="sweep", seed=0xB5B20203), 300, 4, 3, "left"),
    "left_w100":  (dict(named="sweep", seed=0xB5B20204), 300, 100, 2, "left"),
}


def main():
    OUT.mkdir(exist_ok=True)
    ksw = KswReference()
    P = make_params()
    for name, (spec, n, w, max_try, rule) in CASES.items():
        pairs, seq_ref, seq_qer = gb.gen_pairs(make_cfg(spec), 0, n)
        prev = None if rule == "left" else pairs["h0"].astype(np.int32)
        got = pairs.copy()
        band = ksw.band_retry(P, got, seq_ref, seq_qer, w, max_try, prev)
        fields = np.stack([got[f] for f in gb.RESULT_FIELDS], axis=1).astype(np.int32)
        np.savez_compressed(OUT / f"{name}.npz", len1=pairs["len1"], len2=pairs["len2"], h0=pairs["h0"],
                            idr=pairs["idr"], idq=pairs["idq"], seq_ref=seq_ref, seq_qer=seq_qer,
                            w=np.int32(w), max_try=np.int32(max_try), prev=(np.zeros(0, np.int32) if prev is None else prev),
                            expect=fields, band=band)
        print(f"{name}: n={n} w={w} retried={(band > w).sum()} max band {band.max()}")


if __name__ == "__main__":
    main()
