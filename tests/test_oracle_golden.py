"""The CPU restatement (oracle/ksw_oracle.c) against the golden vectors that were produced by
executing the reference's AVX2 getScores16 and its scalarBandedSWA (tests/golden/make_golden.py).
Bit-exact on all six SeqPair result fields."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, results_matrix
from oracle.pyoracle import make_params


def test_golden_present():
    assert len(GOLDEN_CASES) >= 12


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_matches_getscores16(oracle, case):
    pairs, ref, qer, w, params, expect, _ = load_golden(case)
    oracle.batch(make_params(**params, zdrop_mode=0), pairs, ref, qer, w, nthreads=2)
    got = results_matrix(pairs)
    bad = np.nonzero((got != expect).any(axis=1))[0]
    assert bad.size == 0, f"{case}: {bad.size} pairs differ, first {bad[:3]}: got {got[bad[:3]]} want {expect[bad[:3]]}"


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_scalar_rule_matches_scalarBandedSWA(oracle, case):
    pairs, ref, qer, w, params, _, scalar = load_golden(case)
    oracle.batch(make_params(**params, zdrop_mode=1), pairs, ref, qer, w, nthreads=2)
    got = results_matrix(pairs)
    assert np.array_equal(got, scalar), f"{case}: {(got != scalar).any(axis=1).sum()} pairs differ"


def test_pair_entry_point_and_cells(oracle):
    pairs, ref, qer, w, params, expect, _ = load_golden("small_151bp")
    P = make_params(**params)
    total = 0
    for k in range(8):
        q = qer[pairs["idq"][k]: pairs["idq"][k] + pairs["len2"][k]].copy()
        t = ref[pairs["idr"][k]: pairs["idr"][k] + pairs["len1"][k]].copy()
        r = oracle.pair(P, q, t, w, int(pairs["h0"][k]))
        assert [r[f] for f in ("score", "qle", "tle", "gtle", "gscore", "max_off")] == list(expect[k])
        trips = oracle.row_trips(P, q, t, w, int(pairs["h0"][k]))
        assert trips.sum() == r["cells"] and r["cells"] > 0
        total += r["cells"]
    sub = pairs[:8].copy()
    assert oracle.batch(P, sub, ref, qer, w, nthreads=1) == total
