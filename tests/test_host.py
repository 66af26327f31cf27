"""Host-side logic that needs no GPU: generator, bucketing, partitioner, text format."""
import numpy as np
import pytest


def test_generator_deterministic_and_shardable(lib):
    cfg = lib.gen_named_config("large")
    a, ra, qa = lib.gen_pairs(cfg, 0, 300)
    b, rb, qb = lib.gen_pairs(cfg, 0, 300)
    assert np.array_equal(a, b) and np.array_equal(ra, rb) and np.array_equal(qa, qb)
    c, rc, qc = lib.gen_pairs(cfg, 100, 50)              # a shard of the same stream
    for k in range(50):
        assert c["len1"][k] == a["len1"][100 + k] and c["len2"][k] == a["len2"][100 + k]
        assert c["h0"][k] == a["h0"][100 + k]
        assert np.array_equal(rc[c["idr"][k]: c["idr"][k] + c["len1"][k]],
                              ra[a["idr"][100 + k]: a["idr"][100 + k] + a["len1"][100 + k]])
        assert np.array_equal(qc[c["idq"][k]: c["idq"][k] + c["len2"][k]],
                              qa[a["idq"][100 + k]: a["idq"][100 + k] + a["len2"][100 + k]])


def test_named_config_envelopes(lib):
    p, r, q = lib.gen_pairs(lib.gen_named_config("short8"), 0, 5000)
    assert p["len2"].min() >= 16 and p["len2"].max() <= 96 and p["len1"].max() <= 127 and p["len1"].min() >= 1
    assert (p["h0"] + p["len2"]).max() <= 127 and p["h0"].min() >= 1          # int8 envelope of getScores8
    p, r, q = lib.gen_pairs(lib.gen_named_config("small"), 0, 2000)
    assert (p["len2"] == 151).all() and 240 <= p["len1"].min() and p["len1"].max() <= 262
    assert r.max() <= 3 and q.max() <= 3
    p, r, q = lib.gen_pairs(lib.gen_named_config("long16"), 0, 500)
    assert (p["len2"] == 250).all() and p["len1"].min() >= 530 and p["len1"].max() <= 770
    assert p["h0"].min() >= 100 and p["h0"].max() <= 250
    for name, n in (("small", 10_000), ("short8", 1_000_000), ("long16", 1_000_000), ("large", 50_000_000),
                    ("sweep", 8_000_000)):
        assert lib.gen_named_config(name).n_pairs == n


def test_error_rate_is_respected(lib, oracle):
    from oracle.pyoracle import make_params
    cfg = lib.gen_named_config("small")
    p, r, q = lib.gen_pairs(cfg, 0, 400)
    oracle.batch(make_params(), p, r, q, 100)
    # 2% error on 151 bp: most reads extend to the end of the query with a high score
    assert np.median(p["score"]) > 120 and (p["qle"] == 151).mean() > 0.5


def test_bucket_order_is_sorted_permutation(lib):
    p, _, _ = lib.gen_pairs(lib.gen_named_config("large"), 0, 20000)
    order = lib.bucket_order(p)
    assert np.array_equal(np.sort(order), np.arange(len(p)))
    key = (p["len2"][order].astype(np.int64) << 30) | (p["h0"][order].astype(np.int64) << 15) | p["len1"][order]
    assert (np.diff(key) >= 0).all()
    assert len(lib.bucket_order(p[:0])) == 0


@pytest.mark.parametrize("shards", [1, 2, 3, 8])
def test_split_by_cost_balanced_and_contiguous(lib, shards):
    """The multi-GPU partitioner's cut (replaces the OpenMP batch loop, main_banded.cpp:279-291): contiguous ranges of
    the input order with equal sum len1 * min(len2, 2w + 1) -- on a batch whose cost is not uniform in input order."""
    a, _, _ = lib.gen_pairs(lib.gen_named_config("short8"), 0, 40000)
    b, _, _ = lib.gen_pairs(lib.gen_named_config("long16"), 0, 20000)
    p = np.zeros(len(a) + len(b), dtype=lib.SEQPAIR_DTYPE)
    p[:len(a)] = a; p[len(a):] = b
    w = 100
    begin = lib.split_by_cost(p, w, shards)
    assert begin[0] == 0 and begin[-1] == len(p) and (np.diff(begin) > 0).all()
    cost = p["len1"].astype(np.int64) * np.minimum(p["len2"], 2 * w + 1) + 64
    per = np.add.reduceat(cost, begin[:-1]).astype(np.float64)
    assert per.max() / per.mean() < 1.01
    if shards > 1:
        assert begin[1] > len(p) // shards          # more of the cheap pairs in the first range than an equal count
    with pytest.raises(lib.BswError):
        bad = p.copy(); bad["len1"][5] = 0
        lib.split_by_cost(bad, w, shards)


def test_text_format_roundtrip(lib, tmp_path):
    cfg = lib.gen_named_config("small")
    cfg.n_rate = 0.01
    p, r, q = lib.gen_pairs(cfg, 0, 64)
    path = str(tmp_path / "pairs.txt")
    lib.write_pairs_file(path, p, r, q)
    lines = open(path).read().split("\n")
    assert len(lines) == 3 * 64 + 1 and lines[0] == str(p["h0"][0])          # main_banded.cpp:131-141
    assert set("".join(lines[1:3])) <= set("01234")
    p2, r2, q2 = lib.read_pairs_file(path)
    assert len(p2) == 64
    for k in range(64):
        assert (p2["len1"][k], p2["len2"][k], p2["h0"][k]) == (p["len1"][k], p["len2"][k], p["h0"][k])
        assert np.array_equal(r2[p2["idr"][k]: p2["idr"][k] + p2["len1"][k]], r[p["idr"][k]: p["idr"][k] + p["len1"][k]])
        assert np.array_equal(q2[p2["idq"][k]: p2["idq"][k] + p2["len2"][k]], q[p["idq"][k]: p["idq"][k] + p["len2"][k]])
    with pytest.raises(lib.BswError):
        lib.read_pairs_file(str(tmp_path / "missing.txt"))
