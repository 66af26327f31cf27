"""Size-independent properties at larger sizes, where running the oracle on everything would
be slow: order / batching invariance, idempotence, stage-run-fetch == extend, and a sampled
oracle check."""
import numpy as np
import pytest

from conftest import results_matrix
from oracle.pyoracle import make_params

pytestmark = pytest.mark.gpu


def test_order_and_batching_invariance(lib):
    cfg = lib.gen_named_config("large")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 200_000)
    with lib.Engine() as eng:
        a = pairs.copy()
        eng.extend(a, ref, qer, 100)
        perm = np.random.default_rng(3).permutation(len(pairs))
        b = pairs[perm].copy()
        eng.extend(b, ref, qer, 100)
        assert np.array_equal(results_matrix(b), results_matrix(a)[perm])       # per-pair results ignore neighbours
        c = pairs.copy()
        for lo in range(0, len(c), 37_123):                                     # ragged batches
            part = c[lo: lo + 37_123]
            eng.extend(part, ref, qer, 100)
        assert np.array_equal(results_matrix(c), results_matrix(a))
        d = pairs.copy()
        eng.stage(d, ref, qer, 100)
        eng.run_staged(); eng.run_staged()                                      # idempotent re-run
        eng.fetch(d)
        assert np.array_equal(results_matrix(d), results_matrix(a))
        assert (a["id"] == pairs["id"]).all() and (a["h0"] == pairs["h0"]).all()  # inputs untouched


def test_full_size_short8_sampled_against_oracle(lib, oracle):
    """BASELINE configs[1] at full size (1M pairs): sanity invariants on everything, oracle on a sample."""
    cfg = lib.gen_named_config("short8")
    pairs, ref, qer = lib.gen_pairs(cfg)
    assert len(pairs) == 1_000_000
    with lib.Engine() as eng:
        eng.extend(pairs, ref, qer, 100)
        st = eng.stats()
    assert (pairs["score"] >= pairs["h0"]).all()
    assert (pairs["qle"] >= 0).all() and (pairs["qle"] <= pairs["len2"]).all()
    assert (pairs["tle"] >= 0).all() and (pairs["tle"] <= pairs["len1"]).all()
    assert (pairs["gtle"] <= pairs["len1"]).all() and (pairs["gscore"] >= -1).all()
    assert (pairs["max_off"] <= np.maximum(pairs["len1"], pairs["len2"])).all()
    idx = np.random.default_rng(11).choice(len(pairs), 40_000, replace=False)
    sub = pairs[idx].copy()
    want = sub.copy()
    oracle.batch(make_params(), want, ref, qer, 100)
    assert np.array_equal(results_matrix(sub), results_matrix(want))
    assert st["cells_effective"] > 0.5 * st["cells_nominal"]


def test_multi_engine_same_device(lib):
    """One engine per caller thread, as main_banded.cpp:253-258 constructs them."""
    cfg = lib.gen_named_config("small")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 4096)
    e1, e2 = lib.Engine(), lib.Engine()
    a, b = pairs.copy(), pairs.copy()
    e1.extend(a, ref, qer, 100)
    e2.extend(b, ref, qer, 100)
    assert np.array_equal(results_matrix(a), results_matrix(b))
    e1.close(); e2.close()
