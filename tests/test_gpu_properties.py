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


def test_concurrent_engines_from_threads(lib):
    """The reference driver's OpenMP loop shape (main_banded.cpp:253-291): one engine per caller thread, all
    on the same GPU at once, two on pinned buffers (direct route) and two on pageable ones (staged route).
    Results equal a single engine's."""
    import threading
    cfg = lib.gen_named_config("large")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 120_000)
    with lib.Engine() as e0:
        want = pairs.copy()
        e0.extend(want, ref, qer, 100)
    parts = np.array_split(np.arange(len(pairs)), 4)
    outs = [pairs[p].copy() for p in parts]
    pin = [(lib.pinned_copy(o), lib.pinned_copy(ref), lib.pinned_copy(qer)) for o in outs[:2]]
    errors = []

    def work(k):
        try:
            with lib.Engine() as e:
                for _ in range(2):
                    if k < 2:
                        e.extend(pin[k][0], pin[k][1], pin[k][2], 100)
                    else:
                        e.extend(outs[k], ref, qer, 100)
        except Exception as exc:                                     # surfaced below, in the main thread
            errors.append((k, exc))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(4):
        got = pin[k][0] if k < 2 else outs[k]
        assert np.array_equal(results_matrix(got), results_matrix(want)[parts[k]]), k


def test_direct_route_equals_staged_route(lib, oracle):
    """Pinned host buffers (bsw_host_alloc) take the engine's direct route -- records and sequences
    DMA'd as they are, results written into the records from the device; pageable buffers take the
    staged route.  Same results, and both equal the oracle."""
    cfg = lib.gen_named_config("large")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 300_000)             # > 1 chunk
    ref[::997] = 4                                               # sprinkle N: byte-kernel path on both routes
    want = pairs[:30_000].copy()
    oracle.batch(make_params(), want, ref, qer, 100)
    with lib.Engine() as eng:
        a = pairs.copy()
        eng.extend(a, ref, qer, 100)
        st_a = eng.stats()
        pp, pr, pq = lib.pinned_copy(pairs), lib.pinned_copy(ref), lib.pinned_copy(qer)
        eng.extend(pp, pr, pq, 100)
        st_p = eng.stats()
        assert np.array_equal(results_matrix(pp), results_matrix(a))
        assert np.array_equal(results_matrix(a[:30_000]), results_matrix(want))
        assert st_a["ms_pack"] > 0 and st_p["ms_pack"] == 0        # no host pass on the direct route
        assert st_a["cells_effective"] == st_p["cells_effective"]
        assert (pp["h0"] == pairs["h0"]).all() and (pp["idr"] == pairs["idr"]).all()
        # permuted records over the same pinned sequences: offsets no longer ascend
        perm = np.random.default_rng(5).permutation(len(pairs))
        pb = lib.pinned_copy(pairs[perm])
        eng.extend(pb, pr, pq, 100)
        assert np.array_equal(results_matrix(pb), results_matrix(a)[perm])


def test_pcie_bound_batch_runs_partitioned(lib, oracle):
    """A batch whose DP is shorter than its transfer (short pairs) runs with the SMs split into a
    service partition (copies, scan, bucket, pack, write-back) and a DP partition, small chunks, one
    FIFO H2D stream and speculative sequence copies.  Same results as the staged route and the
    oracle; records in reverse and in random order (the speculative range then misses and the
    scanned range is copied instead) give the same results too."""
    cfg = lib.gen_named_config("short8")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 400_000)
    want = pairs[:20_000].copy()
    oracle.batch(make_params(), want, ref, qer, 100)
    with lib.Engine() as eng:
        a = pairs.copy()
        eng.extend(a, ref, qer, 100)                                 # pageable: staged route
        assert np.array_equal(results_matrix(a[:20_000]), results_matrix(want))
        pp, pr, pq = lib.pinned_copy(pairs), lib.pinned_copy(ref), lib.pinned_copy(qer)
        eng.extend(pp, pr, pq, 100)
        st = eng.stats()
        if st["partitioned"] != 1:                                   # (driver without green contexts, or BSW_SERVICE_SMS=0:
            import warnings                                          #  the engine then runs the batch on whole-device streams)
            warnings.warn("green-context SM partitions unavailable on this box: the batch ran unpartitioned")
        assert np.array_equal(results_matrix(pp), results_matrix(a))
        assert st["h2d_bytes"] < 1.02 * (pairs.nbytes + ref.nbytes + qer.nbytes)     # nothing copied twice
        for order in (np.arange(len(pairs))[::-1].copy(), np.random.default_rng(11).permutation(len(pairs))):
            pb = lib.pinned_copy(pairs[order])
            eng.extend(pb, pr, pq, 100)
            assert np.array_equal(results_matrix(pb), results_matrix(a)[order])
        # a compute-bound batch on the same engine goes back to whole-device streams
        cfg2 = lib.gen_named_config("long16")
        p2, r2, q2 = lib.gen_pairs(cfg2, 0, 140_000)
        b = p2.copy()
        eng.extend(b, r2, q2, 100)
        pp2, pr2, pq2 = lib.pinned_copy(p2), lib.pinned_copy(r2), lib.pinned_copy(q2)
        eng.extend(pp2, pr2, pq2, 100)
        assert eng.stats()["partitioned"] == 0
        assert np.array_equal(results_matrix(pp2), results_matrix(b))


def test_pcie_bound_batch_on_a_multi_device_engine(lib):
    """One engine over two device contexts (the same GPU twice when the box has one): a PCIe-bound batch on
    pinned buffers is dealt chunk by chunk over both, each with its own FIFO H2D stream and SM partitions."""
    import torch
    devices = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    cfg = lib.gen_named_config("short8")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 600_000)
    with lib.Engine() as one:
        want = pairs.copy()
        one.extend(want, ref, qer, 100)
    with lib.Engine(devices=devices) as eng:
        pp, pr, pq = lib.pinned_copy(pairs), lib.pinned_copy(ref), lib.pinned_copy(qer)
        for _ in range(2):
            eng.extend(pp, pr, pq, 100)
            assert np.array_equal(results_matrix(pp), results_matrix(want))
        b = pairs.copy()
        eng.extend(b, ref, qer, 100)                                  # staged route over both contexts
        assert np.array_equal(results_matrix(b), results_matrix(want))


def test_sparse_sequence_buffers_are_read_in_place(lib, oracle):
    """The reference loader's layout (main_banded.cpp:55-58,244-246: one 2048-byte slot per
    reference, 256 per query): on the direct route the engine must not DMA the whole slots."""
    from genomicsbench_b200 import SEQPAIR_DTYPE
    cfg = lib.gen_named_config("small")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 5000)
    n = len(pairs)
    sref = lib.pinned_empty(n * 2048, np.uint8); sqer = lib.pinned_empty(n * 256, np.uint8)
    sref[:] = 0; sqer[:] = 0
    sp = lib.pinned_copy(pairs)
    for k in range(n):
        l1, l2 = int(pairs["len1"][k]), int(pairs["len2"][k])
        sref[k * 2048: k * 2048 + l1] = ref[pairs["idr"][k]: pairs["idr"][k] + l1]
        sqer[k * 256: k * 256 + l2] = qer[pairs["idq"][k]: pairs["idq"][k] + l2]
    sp["idr"] = np.arange(n) * 2048
    sp["idq"] = np.arange(n) * 256
    want = pairs.copy()
    oracle.batch(make_params(), want, ref, qer, 100)
    with lib.Engine() as eng:
        eng.extend(sp, sref, sqer, 100)
        st = eng.stats()
    assert np.array_equal(results_matrix(sp), results_matrix(want))
    assert st["h2d_bytes"] < n * 2304 // 2                         # far less than the slots' span


@pytest.mark.parametrize("devices", [[0, 0], [0, 1], [0, 1, 0]])
def test_engine_deals_chunks_over_its_devices(lib, devices):
    """One engine driving several devices (bsw_params.devices): the call is cut into one contiguous cost-balanced
    range per device, every range runs on its own host thread through its device's chunk pipeline, results land
    in input order (bsw_stage still deals chunks round-robin).  [0, 0] runs the multi-device plumbing -- two device
    contexts, two host threads -- on a single GPU; [0, 1] needs two."""
    import torch
    if max(devices) >= torch.cuda.device_count():
        pytest.skip("needs more GPUs")
    cfg = lib.gen_named_config("large")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 700_000)                 # several chunks per device
    with lib.Engine() as one:
        a = pairs.copy()
        one.extend(a, ref, qer, 100)
        cells = one.stats()["cells_effective"]
    with lib.Engine(devices=devices) as eng:
        b = pairs.copy()
        eng.extend(b, ref, qer, 100)                                 # staged route
        assert np.array_equal(results_matrix(b), results_matrix(a))
        assert eng.stats()["cells_effective"] == cells
        assert eng.stats()["shards"] == len(devices)
        pp, pr, pq = lib.pinned_copy(pairs), lib.pinned_copy(ref), lib.pinned_copy(qer)
        eng.extend(pp, pr, pq, 100)                                  # direct route
        assert np.array_equal(results_matrix(pp), results_matrix(a))
        c = pairs.copy()
        eng.stage(c, ref, qer, 100); eng.run_staged(); eng.fetch(c)
        assert np.array_equal(results_matrix(c), results_matrix(a))


@pytest.mark.parametrize("devices", [[0, 0], [0, 1], [0, 1, 2, 3]])
def test_partitioner_on_the_hot_path(lib, devices):
    """The in-call multi-GPU partitioner (replaces main_banded.cpp:279-291) on a batch whose cost is NOT uniform in
    input order -- short pairs first, long pairs last: the cut balances sum len1*min(len2, 2w+1), not the pair count;
    all three routes (pageable, page-locked, packed) return the single-device results in input order, and a batch
    too small to split runs on one device."""
    import torch
    if max(devices) >= torch.cuda.device_count():
        pytest.skip("needs more GPUs")
    ca, cb = lib.gen_named_config("short8"), lib.gen_named_config("long16")
    pa, ra, qa = lib.gen_pairs(ca, 0, 600_000)
    pb, rb, qb = lib.gen_pairs(cb, 0, 150_000)
    pb["idr"] += len(ra); pb["idq"] += len(qa)
    pairs = np.zeros(len(pa) + len(pb), dtype=lib.SEQPAIR_DTYPE)     # (np.concatenate would drop the record's padding)
    pairs[:len(pa)] = pa; pairs[len(pa):] = pb
    ref = np.concatenate([ra, rb]); qer = np.concatenate([qa, qb])
    pairs["id"] = np.arange(len(pairs))
    g = len(devices)
    cut = lib.split_by_cost(pairs, 100, g)
    assert cut[0] == 0 and cut[-1] == len(pairs) and np.all(np.diff(cut) > 0)
    cost = pairs["len1"].astype(np.int64) * np.minimum(pairs["len2"], 201) + 64
    shares = np.add.reduceat(cost, cut[:-1]) / cost.sum()
    assert np.all(np.abs(shares - 1.0 / g) < 0.02), shares
    assert cut[1] > len(pairs) // g                                  # more (cheap) pairs in the first range than an equal count
    with lib.Engine() as one:
        a = pairs.copy()
        one.extend(a, ref, qer, 100)
        cells = one.stats()["cells_effective"]
    with lib.Engine(devices=devices) as eng:
        b = pairs.copy()
        eng.extend(b, ref, qer, 100)
        st = eng.stats()
        assert np.array_equal(results_matrix(b), results_matrix(a))
        assert st["cells_effective"] == cells and st["shards"] == g
        pp, pr, pq = lib.pinned_copy(pairs), lib.pinned_copy(ref), lib.pinned_copy(qer)
        eng.extend(pp, pr, pq, 100)
        assert np.array_equal(results_matrix(pp), results_matrix(a)) and eng.stats()["shards"] == g
        batch = lib.PackedBatch.from_pairs(pairs, ref, qer, pinned=True)
        out = eng.extend_packed(batch, 100)
        assert eng.stats()["shards"] == g and eng.stats()["cells_effective"] == cells
        for f in lib.RESULT_FIELDS:
            assert np.array_equal(out[f], a[f]), f
        small = pairs[:3000].copy()
        eng.extend(small, ref, qer, 100)
        assert eng.stats()["shards"] == 1 and np.array_equal(results_matrix(small), results_matrix(a[:3000]))
        bad = pairs.copy()
        bad["len2"][700_000] = 0                                     # a domain error inside one device's range fails the call
        with pytest.raises(lib.BswError) as ei:
            eng.extend(bad, ref, qer, 100)
        assert ei.value.code == -2
        eng.extend(b, ref, qer, 100)                                 # and the engine keeps working
        assert np.array_equal(results_matrix(b), results_matrix(a))


def test_pinned_buffers_with_far_apart_sequences(lib, oracle):
    """Page-locked buffers whose pairs, in one chunk, address sequences more than 2^30 bytes apart (shuffled pair order
    over a large reference buffer): beyond the direct route's 32-bit chunk-relative offsets, so the call runs on the
    staged route instead -- same results, no error (the pageable route always accepted such input)."""
    cfg = lib.gen_named_config("small")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 4000)
    want = pairs.copy()
    oracle.batch(make_params(), want, ref, qer, 100)
    far = (1 << 30) + (1 << 27)
    big = lib.pinned_empty(far + len(ref) + 64, np.uint8)
    big[:len(ref)] = ref
    big[far:far + len(ref)] = ref
    pp, pq = lib.pinned_copy(pairs), lib.pinned_copy(qer)
    pp["idr"][1::2] += far                                           # every other pair reads the far copy
    with lib.Engine() as eng:
        eng.extend(pp, big, pq, 100)
        assert np.array_equal(results_matrix(pp), results_matrix(want))
        eng.stage(pp, big, pq, 100); eng.run_staged()
        c = pairs.copy(); eng.fetch(c)
        assert np.array_equal(results_matrix(c), results_matrix(want))
