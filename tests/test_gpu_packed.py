"""bsw_extend_packed -- the packed host format through the C ABI on the GPU: bit-exact against the reference's
golden vectors and against bsw_extend on the same pairs, for every builder variant (2-bit, RAW, unordered),
both result formats, pageable and page-locked buffers, and the error paths."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, results_matrix
from oracle.pyoracle import make_params

pytestmark = pytest.mark.gpu

FIELDS = ("score", "qle", "tle", "gtle", "gscore", "max_off")


def out_matrix(out):
    return np.stack([out[f] for f in FIELDS], axis=1).astype(np.int32)


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("warp_max_pairs", [-1, 0])
def test_packed_matches_reference_golden(lib, case, warp_max_pairs):
    """warp_max_pairs = -1: the thread-per-pair kernel; 0 (default): batches this small run the warp-per-pair register kernel."""
    pairs, ref, qer, w, params, expect, _ = load_golden(case)
    b = lib.PackedBatch.from_pairs(pairs, ref, qer)
    with lib.Engine(warp_max_pairs=warp_max_pairs, **params) as eng:
        out = eng.extend_packed(b, w)
        st = eng.stats()
        out16 = eng.extend_packed(b, w, compact=True)
    assert np.array_equal(out_matrix(out), expect), case
    assert np.array_equal(out_matrix(out16), expect), case
    assert st["kernel_launches"] >= 1 and st["cells_effective"] > 0
    assert st["h2d_bytes"] >= b.nbytes() and st["d2h_bytes"] == 24 * len(pairs)


@pytest.mark.parametrize("case", ["small_151bp", "with_N", "long_1k", "a2_b3_o4_e2"])
def test_packed_all_raw_and_pinned(lib, case):
    """Every pair RAW (one byte per base: the byte-reading kernels), long_min_qlen routing, page-locked buffers."""
    pairs, ref, qer, w, params, expect, _ = load_golden(case)
    b = lib.PackedBatch.from_pairs(pairs, ref, qer, raw_min_qlen=1, pinned=True)
    assert (b.desc["flags"] & 1).all() and b.c.q2_words == 0
    with lib.Engine(**params) as eng:
        out = lib.pinned_empty(len(pairs), lib.OUTSCORE_DTYPE)
        eng.extend_packed(b, w, out=out)
        assert np.array_equal(out_matrix(out), expect)
    with lib.Engine(long_min_qlen=1, **params) as eng:                 # every pair on the warp-per-pair kernel
        assert np.array_equal(out_matrix(eng.extend_packed(b, w)), expect)
    b2 = lib.PackedBatch.from_pairs(pairs, ref, qer, pinned=True)      # 2-bit pairs, page-locked
    with lib.Engine(**params) as eng:
        assert np.array_equal(out_matrix(eng.extend_packed(b2, w)), expect)
        with lib.Engine(long_min_qlen=100, **params) as eng2:
            # 2-bit pairs the engine would route to the byte-reading long kernel: refused, not mis-read
            if (pairs["len2"] >= 100).any() and not (b2.desc["flags"][pairs["len2"] >= 100] & 1).all():
                with pytest.raises(lib.BswError) as ei:
                    eng2.extend_packed(b2, w)
                assert ei.value.code == -2
            b3 = lib.PackedBatch.from_pairs(pairs, ref, qer, raw_min_qlen=100)
            assert np.array_equal(out_matrix(eng2.extend_packed(b3, w)), expect)


def test_packed_unordered_batch(lib):
    """Descriptors in shuffled order over the same buffers (ordered = 0): all words are copied up front."""
    pairs, ref, qer, w, params, expect, _ = load_golden("large_mix")
    b = lib.PackedBatch.from_pairs(pairs, ref, qer)
    perm = np.random.default_rng(3).permutation(len(pairs))
    d = b.desc
    d[:] = d[perm]
    b.c.ordered = 0
    with lib.Engine(**params) as eng:
        assert np.array_equal(out_matrix(eng.extend_packed(b, w)), expect[perm])
        b.c.ordered = 1                                                # a false promise is detected or harmless, never UB
        try:
            out = eng.extend_packed(b, w)
            assert np.array_equal(out_matrix(out), expect[perm])
        except lib.BswError as e:
            assert e.code in (-1, -2)


def test_packed_equals_extend_on_seeded_stream(lib, oracle):
    """400 k pairs of the sweep config (10 % error, N sprinkled in): packed route == SeqPair route == oracle sample."""
    cfg = lib.gen_named_config("sweep")
    cfg.n_rate = 0.0001
    pairs, ref, qer = lib.gen_pairs(cfg, 12345, 400_000)
    b = lib.PackedBatch.from_pairs(pairs, ref, qer, pinned=True)
    assert 0 < (b.desc["flags"] & 1).sum() < 0.2 * len(pairs)
    with lib.Engine() as eng:
        out = lib.pinned_empty(len(pairs), lib.OUTSCORE_DTYPE)
        eng.extend_packed(b, 100, out=out)
        st = eng.stats()
        eng.extend(pairs, ref, qer, 100)
    assert np.array_equal(out_matrix(out), results_matrix(pairs))
    assert (st["h2d_bytes"] + st["d2h_bytes"]) / len(pairs) < 200        # 2 bits per base + 16 + 24, ~4 % RAW pairs (594 B in the byte layout)
    idx = np.random.default_rng(5).choice(len(pairs), 15000, replace=False)
    sub = pairs[idx].copy()
    oracle.batch(make_params(), sub, ref, qer, 100)
    assert np.array_equal(out_matrix(out)[idx], results_matrix(sub))


def test_packed_wide_fallback(lib, oracle):
    """2-bit pairs outside the 16-bit kernel's score domain: the chunk runs the 32-bit kernel as a whole."""
    cfg = lib.gen_named_config("large")
    pairs, ref, qer = lib.gen_pairs(cfg, 99, 40_000)
    hi = np.random.default_rng(11).random(len(pairs)) < 0.01
    pairs["h0"][hi] = 20000
    want = pairs.copy()
    oracle.batch(make_params(), want, ref, qer, 100)
    b = lib.PackedBatch.from_pairs(pairs, ref, qer)
    with lib.Engine() as eng:
        assert np.array_equal(out_matrix(eng.extend_packed(b, 100)), results_matrix(want))


def test_packed_errors_and_empty(lib):
    pairs, ref, qer, w, params, expect, _ = load_golden("short8")
    with lib.Engine(**params) as eng:
        b = lib.PackedBatch.from_pairs(pairs, ref, qer)
        b.desc["q_off"][5] = int(b.c.q2_words) + 7                     # points past the buffer
        with pytest.raises(lib.BswError) as ei:
            eng.extend_packed(b, w)
        assert ei.value.code in (-1, -2)
        b = lib.PackedBatch.from_pairs(pairs, ref, qer)
        b.desc["len1"][9] = 0
        with pytest.raises(lib.BswError) as ei:
            eng.extend_packed(b, w)
        assert ei.value.code == -2
        e = lib.PackedBatch.from_pairs(pairs[:0].copy(), ref, qer)
        assert len(eng.extend_packed(e, w)) == 0
        # the engine still works after the failures
        b = lib.PackedBatch.from_pairs(pairs, ref, qer)
        assert np.array_equal(out_matrix(eng.extend_packed(b, w)), expect)
