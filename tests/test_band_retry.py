"""SURVEY 8(f).1 -- band-doubling retry of the aligner (tools/bwa/bwamem.c:630,723-753,770-800).
Golden vectors come from the reference's own ksw_extend2 run inside that loop
(tests/golden/make_golden_retry.py)."""
from pathlib import Path

import numpy as np
import pytest

from conftest import results_matrix
from oracle.pyoracle import make_params

RETRY_DIR = Path(__file__).resolve().parent / "golden" / "retry"
CASES = sorted(p.stem for p in RETRY_DIR.glob("*.npz"))


def load_case(name):
    from genomicsbench_b200 import SEQPAIR_DTYPE
    z = np.load(RETRY_DIR / f"{name}.npz")
    n = len(z["len1"])
    pairs = np.zeros(n, dtype=SEQPAIR_DTYPE)
    for f in ("len1", "len2", "h0", "idr", "idq"):
        pairs[f] = z[f]
    prev = z["prev"] if len(z["prev"]) else None
    return (pairs, np.ascontiguousarray(z["seq_ref"]), np.ascontiguousarray(z["seq_qer"]), int(z["w"]),
            int(z["max_try"]), prev, z["expect"].astype(np.int32), z["band"].astype(np.int32))


def test_retry_goldens_present():
    assert len(CASES) >= 4


@pytest.mark.parametrize("case", CASES)
def test_oracle_band_retry_matches_reference_loop(oracle, case):
    pairs, ref, qer, w, max_try, prev, expect, band = load_case(case)
    got_band = oracle.band_retry(make_params(), pairs, ref, qer, w, max_try, prev)
    assert np.array_equal(results_matrix(pairs), expect)
    assert np.array_equal(got_band, band)


def test_oracle_band_retry_against_live_ksw_extend2(oracle):
    from oracle.pyoracle import KswReference
    if not KswReference.available():
        pytest.skip("oracle/_ref/libkswref.so not built (needs /root/reference)")
    import genomicsbench_b200 as gb
    cfg = gb.gen_named_config("sweep")
    cfg.seed = 0xB5B20299
    pairs, ref, qer = gb.gen_pairs(cfg, 0, 1500)
    a, b = pairs.copy(), pairs.copy()
    band_a = oracle.band_retry(make_params(), a, ref, qer, 6, 3)
    band_b = KswReference().band_retry(make_params(), b, ref, qer, 6, 3)
    assert np.array_equal(results_matrix(a), results_matrix(b)) and np.array_equal(band_a, band_b)
    assert (band_a > 6).any() and (band_a > 12).any()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_band_retry_matches_reference_loop(lib, case):
    pairs, ref, qer, w, max_try, prev, expect, band = load_case(case)
    with lib.Engine() as eng:
        got_band = eng.extend_retry(pairs, ref, qer, w, max_try, prev)
    assert np.array_equal(results_matrix(pairs), expect)
    assert np.array_equal(got_band, band)


@pytest.mark.gpu
def test_cuda_band_retry_matches_oracle_at_size(lib, oracle):
    cfg = lib.gen_named_config("sweep")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 60_000)
    want = pairs.copy()
    want_band = oracle.band_retry(make_params(), want, ref, qer, 16, 3, want["h0"].astype(np.int32))
    with lib.Engine() as eng:
        band = eng.extend_retry(pairs, ref, qer, 16, 3, pairs["h0"].astype(np.int32))
        st = eng.stats()
    assert np.array_equal(results_matrix(pairs), results_matrix(want))
    assert np.array_equal(band, want_band)
    assert (band > 16).any() and st["pairs"] == len(pairs)
