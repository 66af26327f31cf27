"""Generates the golden vectors under tests/golden/ by EXECUTING THE REFERENCE.

Run in the build container (needs /root/reference -> oracle/_ref/libbswref.so):
    python tests/golden/make_golden.py
For every case below it draws seeded synthetic pairs with the repo's generator, runs the
reference's own AVX2 getScores16 (benchmarks/bsw/bandedSWA.cpp:1124-1148) and its
scalarBandedSWAWrapper (:254-272) on them, and stores inputs + both outputs in one .npz.
`expect` is getScores16's batch output; any pair whose batch output differs from the same
pair run alone in its SIMD group is a reference lane-interaction artifact (SURVEY.md
Appendix B, Q4) and gets the solo output instead (count stored in `q4_artifacts`).
The reference repo has no bsw test vectors of its own, so these are
the pin for oracle/ksw_oracle.c and for the CUDA kernels.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import genomicsbench_b200 as gb                      # noqa: E402  (host-only generator)
from oracle.pyoracle import Reference, make_params  # noqa: E402

OUT = Path(__file__).resolve().parent

# name -> (generator overrides, n pairs, scoring params, w)
CASES = {
    "small_151bp":   (dict(named="small"), 256, dict(), 100),
    "short8":        (dict(named="short8"), 512, dict(), 100),
    "long16_250bp":  (dict(named="long16"), 96, dict(), 100),
    "large_mix":     (dict(named="large"), 256, dict(), 100),
    "sweep_w32":     (dict(named="sweep"), 256, dict(zdrop=100), 32),
    "sweep_w500_zoff": (dict(named="sweep"), 128, dict(zdrop=32767), 500),
    "with_N":        (dict(named="small", n_rate=0.01, seed=0xB5B20101), 256, dict(), 100),
    "asym_gaps":     (dict(named="large", seed=0xB5B20102), 192, dict(o_del=6, e_del=1, o_ins=8, e_ins=1), 100),
    "gap_e2_e3_z20": (dict(named="large", seed=0xB5B20103), 192, dict(o_del=5, e_del=2, o_ins=7, e_ins=3, zdrop=20), 100),
    "a2_b3_o4_e2":   (dict(named="sweep", seed=0xB5B20104, n_rate=0.02), 192,
                      dict(match=2, mismatch=3, o_del=4, e_del=2, o_ins=4, e_ins=2, zdrop=50), 500),
    "tiny_w3":       (dict(qlen_min=1, qlen_max=8, tail_min=0, tail_max=6, h0_min=1, h0_max=10,
                           error_rate=0.1, seed=0xB5B20105), 512, dict(zdrop=5), 3),
    "w0_w1":         (dict(qlen_min=5, qlen_max=60, tail_min=0, tail_max=30, h0_min=1, h0_max=8,
                           error_rate=0.05, seed=0xB5B20106), 256, dict(), 1),
    "high_h0":       (dict(qlen_min=100, qlen_max=400, tail_min=20, tail_max=300, h0_min=200, h0_max=1000,
                           error_rate=0.03, seed=0xB5B20107), 128, dict(), 100),
    "endbonus0_z1":  (dict(named="small", seed=0xB5B20108), 128, dict(end_bonus=0, zdrop=1), 16),
    "long_1k":       (dict(qlen_min=700, qlen_max=1500, tail_min=100, tail_max=600, h0_min=19, h0_max=300,
                           error_rate=0.04, seed=0xB5B20109), 24, dict(), 100),
}


def make_cfg(spec: dict):
    spec = dict(spec)
    named = spec.pop("named", None)
    cfg = gb.gen_named_config(named) if named else gb.gen_named_config("small")
    if not named:
        cfg.max_len1 = 0
        cfg.max_score8 = 0
        cfg.n_rate = 0.0
    for k, v in spec.items():
        setattr(cfg, k, v)
    return cfg


def main():
    ref = Reference()
    for name, (spec, n, sc, w) in CASES.items():
        cfg = make_cfg(spec)
        pairs, seq_ref, seq_qer = gb.gen_pairs(cfg, 0, n)
        P = make_params(**sc)
        vec = pairs.copy()
        ref.getscores16(P, vec, seq_ref, seq_qer, w, batch=512, nthreads=1)
        sca = pairs.copy()
        ref.scalar(make_params(**dict(sc, zdrop_mode=1)), sca, seq_ref, seq_qer, w)
        expect = vec.copy()
        q4 = 0
        for k in range(n):
            solo = pairs[k:k + 1].copy()
            ref.solo(P, solo, seq_ref, seq_qer, w)
            if any(solo[f][0] != vec[f][k] for f in gb.RESULT_FIELDS):
                q4 += 1
                for f in gb.RESULT_FIELDS:
                    expect[f][k] = solo[f][0]
        fields = np.stack([expect[f] for f in gb.RESULT_FIELDS], axis=1).astype(np.int32)
        sfields = np.stack([sca[f] for f in gb.RESULT_FIELDS], axis=1).astype(np.int32)
        np.savez_compressed(
            OUT / f"{name}.npz",
            len1=pairs["len1"], len2=pairs["len2"], h0=pairs["h0"], idr=pairs["idr"], idq=pairs["idq"],
            seq_ref=seq_ref, seq_qer=seq_qer, w=np.int32(w),
            params=np.array([P.o_del, P.e_del, P.o_ins, P.e_ins, P.zdrop, P.end_bonus, P.match, P.mismatch,
                             P.ambig], dtype=np.int32),
            expect=fields, scalar=sfields, q4_artifacts=np.int32(q4), fields=np.array(gb.RESULT_FIELDS))
        ndiff = int((fields != sfields).any(axis=1).sum())
        print(f"{name}: n={n} w={w} q4_artifacts={q4} vec!=scalar pairs={ndiff} "
              f"mean score {fields[:, 0].mean():.1f}")


if __name__ == "__main__":
    main()
