"""Pins the oracle on the reference itself: the unmodified bandedSWA.cpp compiled into
oracle/_ref/libbswref.so (only where /root/reference is mounted; skipped on the GPU box).
Includes the Q4 triage protocol of SURVEY.md finding 0.5."""
import numpy as np
import pytest

import genomicsbench_b200 as gb
from conftest import results_matrix
from oracle.pyoracle import make_params

SWEEP = [
    # (named config, overrides, scoring, w, n)
    ("small", {}, {}, 100, 3000),
    ("short8", {}, {}, 100, 6000),
    ("long16", {}, {}, 100, 1200),
    ("large", {}, {}, 100, 3000),
    ("sweep", {}, {"zdrop": 100}, 32, 3000),
    ("sweep", {}, {"zdrop": 32767}, 100, 2000),
    ("sweep", {}, {"zdrop": 32767}, 500, 1500),
    ("large", {"n_rate": 0.02}, {}, 100, 2000),
    ("large", {}, {"o_del": 5, "e_del": 2, "o_ins": 7, "e_ins": 3, "zdrop": 20}, 100, 2000),
    ("large", {}, {"match": 2, "mismatch": 3, "o_del": 4, "e_del": 2, "o_ins": 4, "e_ins": 2, "zdrop": 50}, 500, 1500),
    ("small", {}, {"end_bonus": 0, "zdrop": 10}, 16, 2000),
    ("small", {}, {"end_bonus": 10, "zdrop": 1000}, 8, 2000),
]


@pytest.mark.parametrize("idx", range(len(SWEEP)))
def test_oracle_equals_reference(oracle, reference, lib, idx):
    name, over, sc, w, n = SWEEP[idx]
    cfg = gb.gen_named_config(name)
    cfg.seed = 0xC0FFEE00 + idx
    for k, v in over.items():
        setattr(cfg, k, v)
    pairs, ref, qer = gb.gen_pairs(cfg, 0, n)
    a, b, c, d = pairs.copy(), pairs.copy(), pairs.copy(), pairs.copy()
    oracle.batch(make_params(**sc), a, ref, qer, w)
    reference.getscores16(make_params(**sc), b, ref, qer, w, batch=512, nthreads=2)
    ga, gb_ = results_matrix(a), results_matrix(b)
    bad = np.nonzero((ga != gb_).any(axis=1))[0]
    # triage: a mismatch is tolerated only if the reference run SOLO agrees with the oracle (Q4)
    for k in bad:
        solo = pairs[k:k + 1].copy()
        reference.solo(make_params(**sc), solo, ref, qer, w)
        assert np.array_equal(results_matrix(solo)[0], ga[k]), f"pair {k}: oracle != reference (not a lane artifact)"
    if w >= 16:
        assert bad.size == 0
    oracle.batch(make_params(**sc, zdrop_mode=1), c, ref, qer, w)
    reference.scalar(make_params(**sc, zdrop_mode=1), d, ref, qer, w)
    assert np.array_equal(results_matrix(c), results_matrix(d))


@pytest.mark.parametrize("w", [5, 100, 400])
def test_h0_zero_is_in_the_domain(oracle, reference, w):
    """ksw_extend2 and getScores16 accept h0 == 0 (score 0, every end 0; gscore 0 / gtle 1 when the first row
    reaches the query's end, else -1 / 0): the oracle agrees with the reference's vector and scalar code."""
    cfg = gb.gen_named_config("large")
    pairs, ref, qer = gb.gen_pairs(cfg, 31, 6000)
    pairs["h0"][::3] = 0
    pairs["h0"][1::7] = 1
    a, b, c, d = pairs.copy(), pairs.copy(), pairs.copy(), pairs.copy()
    oracle.batch(make_params(), a, ref, qer, w)
    reference.getscores16(make_params(), b, ref, qer, w, batch=512, nthreads=2)
    assert np.array_equal(results_matrix(a), results_matrix(b))
    z = pairs["h0"] == 0
    assert (a["score"][z] == 0).all() and (a["qle"][z] == 0).all() and (a["tle"][z] == 0).all()
    oracle.batch(make_params(zdrop_mode=1), c, ref, qer, w)
    reference.scalar(make_params(zdrop_mode=1), d, ref, qer, w)
    assert np.array_equal(results_matrix(c), results_matrix(d))


def test_reference_layout(reference):
    assert reference.lib.ref_sizeof_seqpair() == gb.SEQPAIR_DTYPE.itemsize == 72
