"""Parity tests proper: the CUDA path, called through the C ABI (and the reference-facing class
mirror), against (1) the golden vectors produced by the reference's AVX2 getScores16 and
(2) the oracle on the same seeded inputs.  Bit-exact on all six SeqPair result fields."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, results_matrix
from oracle.pyoracle import make_params

pytestmark = pytest.mark.gpu


def engine_kwargs(params, **extra):
    return dict(params, **extra)


# bsw_params.warp_max_pairs: -1 = the thread-per-pair kernel whatever the batch size, 0 = the default (calls this small
# run their queries of up to 255 bases on the warp-per-pair register kernel, bsw_warp16.cuh)
@pytest.mark.parametrize("warp_max_pairs", [-1, 0])
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_cuda_matches_reference_golden(lib, case, warp_max_pairs):
    pairs, ref, qer, w, params, expect, _ = load_golden(case)
    with lib.Engine(**engine_kwargs(params, warp_max_pairs=warp_max_pairs)) as eng:
        eng.extend(pairs, ref, qer, w)
        st = eng.stats()
    got = results_matrix(pairs)
    bad = np.nonzero((got != expect).any(axis=1))[0]
    assert bad.size == 0, f"{case}: {bad.size} pairs differ; first {bad[:3]} got {got[bad[:3]]} want {expect[bad[:3]]}"
    assert st["kernel_launches"] >= 1 and st["cells_effective"] > 0


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_wide32_kernel_matches_reference_golden(lib, case):
    """Every golden case again with the 32-bit thread-per-pair kernel for all short pairs
    (short_variant = BSW_SHORT_WIDE32; the default is the packed 16-bit kernel)."""
    pairs, ref, qer, w, params, expect, _ = load_golden(case)
    with lib.Engine(**engine_kwargs(params, short_variant=1)) as eng:
        eng.extend(pairs, ref, qer, w)
    assert np.array_equal(results_matrix(pairs), expect)


def test_packed16_domain_routing(lib, oracle):
    """Pairs whose scores leave the packed kernel's 16-bit domain ((h0 + len2*match)*(1+match) > 32767)
    must take the 32-bit kernel inside the same batch: high h0 next to ordinary pairs."""
    import genomicsbench_b200 as gb
    cfg = gb.gen_named_config("large")
    pairs, ref, qer = gb.gen_pairs(cfg, 4242, 6000)
    rng = np.random.default_rng(7)
    hi = rng.random(len(pairs)) < 0.3
    pairs["h0"][hi] = rng.integers(16000, 32000, hi.sum())        # (h0 + len2) * 2 > 32767, still h0 + len2 <= 32767
    want = pairs.copy()
    oracle.batch(make_params(), want, ref, qer, 100)
    with lib.Engine() as eng:
        eng.extend(pairs, ref, qer, 100)
    assert np.array_equal(results_matrix(pairs), results_matrix(want))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_long_kernel_matches_reference_golden(lib, case):
    """Every golden case again with all pairs routed to the warp-per-pair kernel."""
    pairs, ref, qer, w, params, expect, _ = load_golden(case)
    with lib.Engine(**engine_kwargs(params, long_min_qlen=1)) as eng:
        eng.extend(pairs, ref, qer, w)
        st = eng.stats()
    got = results_matrix(pairs)
    bad = np.nonzero((got != expect).any(axis=1))[0]
    assert bad.size == 0, f"{case}: {bad.size} pairs differ; first {bad[:3]} got {got[bad[:3]]} want {expect[bad[:3]]}"
    assert st["n_long"] == len(pairs) and st["n_short"] == 0


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_cuda_scalar_rule_matches_scalarBandedSWA(lib, case):
    pairs, ref, qer, w, params, _, scalar = load_golden(case)
    with lib.Engine(**engine_kwargs(params, zdrop_mode=1)) as eng:
        eng.extend(pairs, ref, qer, w)
    assert np.array_equal(results_matrix(pairs), scalar)


def test_class_mirror_reads_like_the_reference(lib):
    """BandedPairWiseSW(o_del,e_del,o_ins,e_ins,zdrop,end_bonus,mat,w_match,w_mismatch,nthreads)
    + getScores16(pairArray, seqBufRef, seqBufQer, numPairs, numThreads, w): main_banded.cpp:253-258,286."""
    pairs, ref, qer, w, params, expect, scalar = load_golden("small_151bp")
    mat = []
    for i in range(4):                                   # bwa_fill_scmat, main_banded.cpp:73-81
        mat += [1 if i == j else -4 for j in range(4)] + [-1]
    mat += [-1] * 5
    bsw = lib.BandedPairWiseSW(6, 1, 6, 1, 100, 5, mat, 1, 4, 1)
    padded = np.zeros(len(pairs) + 16, dtype=pairs.dtype)          # reference callers over-allocate pads
    padded[: len(pairs)] = pairs
    padded[len(pairs):]["score"] = -77
    bsw.getScores16(padded, ref, qer, len(pairs), 1, w)
    assert np.array_equal(results_matrix(padded[: len(pairs)]), expect)
    assert (padded[len(pairs):]["score"] == -77).all()             # pads are never written
    assert bsw.SW_cells > 0
    p8 = pairs.copy()
    bsw.getScores8(p8, ref, qer, len(p8), 1, w)
    assert np.array_equal(results_matrix(p8), expect)
    ps = pairs.copy()
    bsw.scalarBandedSWAWrapper(ps, ref, qer, len(ps), 1, w)
    assert np.array_equal(results_matrix(ps), scalar)
    bsw.close()


SWEEP = [
    ("small", {}, {}, 100, 10000),
    ("short8", {}, {}, 100, 60000),
    ("long16", {}, {}, 100, 6000),
    ("large", {}, {}, 100, 30000),
    ("sweep", {}, {"zdrop": 100}, 32, 20000),
    ("sweep", {}, {"zdrop": 32767}, 32, 10000),
    ("sweep", {}, {"zdrop": 100}, 500, 10000),
    ("sweep", {}, {"zdrop": 32767}, 500, 10000),
    ("large", {"n_rate": 0.01}, {}, 100, 10000),
    ("large", {}, {"o_del": 5, "e_del": 2, "o_ins": 7, "e_ins": 3, "zdrop": 20}, 100, 10000),
    ("large", {}, {"o_del": 7, "e_del": 3, "o_ins": 5, "e_ins": 2, "zdrop": 30}, 50, 10000),
    ("sweep", {"n_rate": 0.02}, {"match": 2, "mismatch": 3, "o_del": 4, "e_del": 2, "o_ins": 4, "e_ins": 2, "zdrop": 50}, 500, 8000),
    ("small", {}, {"end_bonus": 0, "zdrop": 10}, 16, 8000),
    ("small", {"h0_min": 1, "h0_max": 9}, {}, 1, 8000),
    ("small", {"h0_min": 1, "h0_max": 6}, {"zdrop": 3}, 0, 4000),
    ("long16", {"qlen_min": 300, "qlen_max": 860, "tail_min": 0, "tail_max": 400, "h0_max": 900}, {}, 100, 1500),
]


@pytest.mark.parametrize("warp_max_pairs", [-1, 1 << 20])
@pytest.mark.parametrize("idx", range(len(SWEEP)))
def test_cuda_matches_oracle_on_seeded_inputs(lib, oracle, idx, warp_max_pairs):
    name, over, sc, w, n = SWEEP[idx]
    sc = dict(sc)
    cfg = lib.gen_named_config(name)
    cfg.seed = 0x5EED0000 + idx
    for k, v in over.items():
        setattr(cfg, k, v)
    pairs, ref, qer = lib.gen_pairs(cfg, 0, n)
    want = pairs.copy()
    cells = oracle.batch(make_params(**sc), want, ref, qer, w)
    # warp_max_pairs = 2^20: every plain pair with a query of up to 255 bases on the warp-per-pair register kernel,
    # whatever the batch size; -1: none
    with lib.Engine(**sc, warp_max_pairs=warp_max_pairs) as eng:
        eng.extend(pairs, ref, qer, w)
        st = eng.stats()
    a, b = results_matrix(pairs), results_matrix(want)
    bad = np.nonzero((a != b).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} pairs differ; first {bad[:3]}: gpu {a[bad[:3]]} oracle {b[bad[:3]]}"
    assert st["cells_effective"] == cells                 # the SW_cells hook (bandedSWA.cpp:211) agrees too
    assert st["cells_nominal"] == int((pairs["len1"].astype(np.int64) * pairs["len2"]).sum())
    # same inputs, split between the two kernels at a mid length: results and cell count unchanged
    mid = max(1, int(np.median(pairs["len2"])))          # len2 >= mid -> warp-per-pair kernel
    again = pairs.copy()
    with lib.Engine(**sc, long_min_qlen=mid) as eng:
        eng.extend(again, ref, qer, w)
        st2 = eng.stats()
    assert np.array_equal(results_matrix(again), b)
    assert st2["cells_effective"] == cells and st2["n_long"] > 0


@pytest.mark.parametrize("case", ["small_151bp", "with_N", "long_1k", "tiny_w3", "high_h0", "asym_gaps"])
def test_latency_route_matches_reference_golden(lib, case):
    """bsw_params.tiny_batch: calls with few pairs skip bucketing / packing and run every pair on the
    warp-per-pair kernel with its rows in shared memory (what the BandedPairWiseSW class does for the
    reference driver's -b 512 batches).  Same results, on pageable and on page-locked buffers."""
    pairs, ref, qer, w, params, expect, _ = load_golden(case)
    pairs = pairs[:1200].copy(); expect = expect[:1200]
    with lib.Engine(tiny_batch=1536, warp_max_pairs=-1, **params) as eng:
        a = pairs.copy()
        eng.extend(a, ref, qer, w)
        st = eng.stats()
        assert np.array_equal(results_matrix(a), expect)
        assert st["n_long"] == len(pairs) and st["n_short"] == 0          # every pair took the 32-bit warp-per-pair kernel
        assert st["kernel_launches"] == 1                                 # the fused latency route: one kernel, 4 CUDA calls
    with lib.Engine(tiny_batch=1536, **params) as eng:
        # default: a call of plain pairs (no N, scores within 16 bits, queries <= 255) runs the warp-per-pair REGISTER
        # kernel on 2-bit words packed by the host pass; any other call the kernel above.  One kernel either way.
        a = pairs.copy()
        eng.extend(a, ref, qer, w)
        st = eng.stats()
        assert np.array_equal(results_matrix(a), expect)
        assert (st["n_long"], st["n_short"]) == ((0, len(pairs)) if case in ("small_151bp", "tiny_w3") else (len(pairs), 0))
        assert st["kernel_launches"] == 1
    with lib.Engine(tiny_batch=1536, warp_max_pairs=1, **params) as eng:
        # ... and, when the device's budget of warp-per-pair pairs is taken (here: a budget of one pair), the thread-per-pair
        # kernel on the same words
        a = pairs.copy()
        eng.extend(a, ref, qer, w)
        assert np.array_equal(results_matrix(a), expect)
        assert eng.stats()["cells_effective"] == st["cells_effective"] and eng.stats()["kernel_launches"] == 1
        with lib.Engine(**params) as plain:                               # same effective cells as the throughput route counts
            b = pairs.copy()
            plain.extend(b, ref, qer, w)
            assert plain.stats()["cells_effective"] == st["cells_effective"] > 0
        pp, pr, pq = lib.pinned_copy(pairs), lib.pinned_copy(ref), lib.pinned_copy(qer)
        eng.extend(pp, pr, pq, w)
        assert np.array_equal(results_matrix(pp), expect)
        reps = 2 + 1536 // len(pairs)                                     # above the threshold: the throughput kernels
        big = np.tile(pairs, reps)
        eng.extend(big, ref, qer, w)
        assert eng.stats()["n_short"] > 0 or case == "long_1k"
        assert np.array_equal(results_matrix(big), np.tile(expect, (reps, 1)))


def test_edge_cases(lib, oracle):
    """Empty batch, single pair, 1-base sequences, ragged lengths, N-only sequences, h0 = 1."""
    from genomicsbench_b200 import SEQPAIR_DTYPE
    with lib.Engine() as eng:
        empty = np.zeros(0, dtype=SEQPAIR_DTYPE)
        eng.extend(empty, np.zeros(1, np.uint8), np.zeros(1, np.uint8), 100)
        assert eng.stats()["pairs"] == 0
        rng = np.random.default_rng(7)
        lens = [(1, 1), (1, 50), (50, 1), (2, 3), (127, 128), (128, 127), (129, 129), (300, 17), (17, 300),
                (824, 824), (5, 824), (33, 31), (64, 64), (65, 63)]
        n = len(lens)
        pairs = np.zeros(n, dtype=SEQPAIR_DTYPE)
        ref = rng.integers(0, 4, size=sum(a for a, _ in lens) + 8, dtype=np.uint8)
        qer = rng.integers(0, 4, size=sum(b for _, b in lens) + 8, dtype=np.uint8)
        ro = qo = 0
        for k, (l1, l2) in enumerate(lens):
            pairs[k]["len1"], pairs[k]["len2"], pairs[k]["idr"], pairs[k]["idq"] = l1, l2, ro, qo
            pairs[k]["h0"] = (1 if k % 3 == 0 else 40) if k % 5 else 0          # h0 == 0 is in the domain (ksw_extend2 accepts it)
            m = min(l1, l2)                                # make them alignable
            qer[qo: qo + m] = ref[ro: ro + m]
            ro += l1; qo += l2
        qer[pairs[4]["idq"]: pairs[4]["idq"] + 10] = 4      # a run of N
        ref[pairs[5]["idr"]: pairs[5]["idr"] + 128] = 4     # N-only reference
        want = pairs.copy()
        oracle.batch(make_params(), want, ref, qer, 100)
        eng.extend(pairs, ref, qer, 100)
        assert np.array_equal(results_matrix(pairs), results_matrix(want))


def test_domain_errors(lib):
    from genomicsbench_b200 import SEQPAIR_DTYPE
    with lib.Engine() as eng:
        pairs = np.zeros(2, dtype=SEQPAIR_DTYPE)
        pairs["len1"], pairs["len2"], pairs["h0"] = 10, 10, 5
        buf = np.zeros(64, np.uint8)
        for field, val in (("len1", 0), ("len2", 0), ("h0", -1), ("len2", 40000), ("h0", 32767)):
            bad = pairs.copy()
            bad[field][1] = val
            with pytest.raises(lib.BswError) as ei:
                eng.extend(bad, buf, buf, 100)
            assert ei.value.code == -2
        # the same on page-locked buffers (direct route: the device-side scan finds it; the speculative
        # sequence copy must not act on a malformed first / last record)
        pbuf = lib.pinned_copy(np.zeros(1 << 16, np.uint8))
        for field, val, at in (("len2", 40000, 0), ("len1", 0, 1), ("h0", -3, 1), ("len1", 50000, 1)):
            bad = lib.pinned_copy(pairs)
            bad[field][at] = val
            with pytest.raises(lib.BswError) as ei:
                eng.extend(bad, pbuf, pbuf, 100)
            assert ei.value.code == -2, (field, val)
        with pytest.raises(lib.BswError) as ei:
            eng.run_staged()
        assert ei.value.code == -5
        eng.extend(pairs, buf, buf, 100)                    # engine still usable afterwards
        assert (pairs["score"] >= 5).all()


@pytest.mark.parametrize("kw", [{}, {"short_variant": 1}, {"long_min_qlen": 1}])
def test_h0_zero_pairs(lib, oracle, kw):
    """h0 == 0 next to ordinary pairs, on every kernel (packed 16-bit, 32-bit, warp-per-pair) and both routes."""
    import genomicsbench_b200 as gb
    cfg = gb.gen_named_config("large")
    pairs, ref, qer = gb.gen_pairs(cfg, 31, 20000)
    pairs["h0"][::3] = 0
    for w in (5, 100, 400):
        want = pairs.copy()
        oracle.batch(make_params(), want, ref, qer, w)
        with lib.Engine(**kw) as eng:
            got = pairs.copy()
            eng.extend(got, ref, qer, w)
            assert np.array_equal(results_matrix(got), results_matrix(want)), (kw, w)
            if not kw:
                b = lib.PackedBatch.from_pairs(pairs, ref, qer)
                out = eng.extend_packed(b, w)
                for f in lib.RESULT_FIELDS:
                    assert np.array_equal(out[f], want[f]), (f, w)
