"""Builders of the packed host format (include/bsw.h: bsw_packed_batch) -- host logic only, no GPU:
pack -> unpack round trips over every golden input, RAW routing of N pairs and long queries, the generator's
packed form equals packing its byte form, the text-format loader, libbsw_host.so carries no CUDA."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden

ROOT = Path(__file__).resolve().parent.parent


def unpack_words(words, first, length):
    k = np.arange(length)
    return ((words[first + (k >> 4)] >> ((k & 15) * 2)) & 3).astype(np.uint8)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_pack_unpack_round_trip(lib, case):
    pairs, ref, qer, *_ = load_golden(case)
    b = lib.PackedBatch.from_pairs(pairs, ref, qer)
    d = b.desc
    assert b.n_pairs == len(pairs) and b.c.ordered == 1
    assert np.array_equal(d["len1"], pairs["len1"]) and np.array_equal(d["len2"], pairs["len2"]) and np.array_equal(d["h0"], pairs["h0"])
    # independent decode of a few pairs (numpy bit arithmetic, not the library's unpacker)
    for i in np.linspace(0, len(pairs) - 1, 25).astype(int):
        q = qer[pairs["idq"][i]: pairs["idq"][i] + pairs["len2"][i]]
        r = ref[pairs["idr"][i]: pairs["idr"][i] + pairs["len1"][i]]
        if d["flags"][i] & lib.BSW_PAIR_RAW:
            assert (q.max() > 3 or r.max() > 3 or len(q) > lib.BSW_PACKED_MAX_QLEN)
            assert np.array_equal(b.raw_q[d["q_off"][i]: d["q_off"][i] + len(q)], q)
            assert np.array_equal(b.raw_r[d["r_off"][i]: d["r_off"][i] + len(r)], r)
        else:
            assert q.max() <= 3 and r.max() <= 3
            assert np.array_equal(unpack_words(b.q2, int(d["q_off"][i]), len(q)), q)
            assert np.array_equal(unpack_words(b.r2, int(d["r_off"][i]), len(r)), r)
    # RAW exactly where needed
    has_n = np.array([qer[a:a + l].max() > 3 or ref[c:c + m].max() > 3
                      for a, l, c, m in zip(pairs["idq"], pairs["len2"], pairs["idr"], pairs["len1"])])
    assert np.array_equal((d["flags"] & lib.BSW_PAIR_RAW) != 0, has_n | (pairs["len2"] > lib.BSW_PACKED_MAX_QLEN))
    # offsets ascend (the `ordered` promise)
    pk = d[(d["flags"] & 1) == 0]
    assert np.all(np.diff(pk["q_off"].astype(np.int64)) >= 0) and np.all(np.diff(pk["r_off"].astype(np.int64)) >= 0)
    # library round trip
    p2, r2, q2 = b.to_pairs()
    for i in range(len(pairs)):
        assert np.array_equal(q2[p2["idq"][i]: p2["idq"][i] + p2["len2"][i]], qer[pairs["idq"][i]: pairs["idq"][i] + pairs["len2"][i]])
        assert np.array_equal(r2[p2["idr"][i]: p2["idr"][i] + p2["len1"][i]], ref[pairs["idr"][i]: pairs["idr"][i] + pairs["len1"][i]])


def test_raw_min_qlen_and_sizes(lib):
    cfg = lib.gen_named_config("large")
    pairs, ref, qer = lib.gen_pairs(cfg, 1000, 5000)
    b = lib.PackedBatch.from_pairs(pairs, ref, qer, raw_min_qlen=200)
    d = b.desc
    assert np.array_equal((d["flags"] & 1) != 0, pairs["len2"] >= 200)
    assert b.c.raw_q_bytes == int(pairs["len2"][pairs["len2"] >= 200].sum())
    small = pairs["len2"] < 200
    assert b.c.q2_words == int(((pairs["len2"][small] + 15) // 16).sum())
    assert b.c.r2_words == int(((pairs["len1"][small] + 15) // 16).sum())
    assert b.nbytes() == 16 * len(pairs) + 4 * (b.c.q2_words + b.c.r2_words) + b.c.raw_q_bytes + b.c.raw_r_bytes


def test_generator_packed_equals_packing_its_bytes(lib):
    cfg = lib.gen_named_config("short8")
    pairs, ref, qer = lib.gen_pairs(cfg, 777, 300_000)           # spans more than one generator slice
    a = lib.PackedBatch.from_pairs(pairs, ref, qer)
    g = lib.PackedBatch.gen(cfg, 777, 300_000)
    assert np.array_equal(a.desc, g.desc) and np.array_equal(a.q2, g.q2) and np.array_equal(a.r2, g.r2)
    assert a.nbytes() / a.n_pairs < 75                           # the PCIe budget the format exists for (194 B unpacked)
    cfg.n_rate = 0.003
    pairs, ref, qer = lib.gen_pairs(cfg, 5, 30_000)
    a = lib.PackedBatch.from_pairs(pairs, ref, qer)
    g = lib.PackedBatch.gen(cfg, 5, 30_000)
    assert (a.desc["flags"] & 1).sum() > 100
    assert np.array_equal(a.desc, g.desc) and np.array_equal(a.raw_q, g.raw_q) and np.array_equal(a.raw_r, g.raw_r)


def test_text_format_loader_packs(lib, tmp_path):
    pairs, ref, qer, *_ = load_golden("with_N")
    path = str(tmp_path / "pairs.txt")
    lib.write_pairs_file(path, pairs, ref, qer)
    f = lib.PackedBatch.from_file(path)
    a = lib.PackedBatch.from_pairs(pairs, ref, qer)
    assert np.array_equal(f.desc, a.desc) and np.array_equal(f.q2, a.q2) and np.array_equal(f.raw_r, a.raw_r)
    assert lib.PackedBatch.from_file(path, max_pairs=10).n_pairs == 10
    with pytest.raises(lib.BswError):
        lib.PackedBatch.from_file(str(tmp_path / "missing.txt"))


def test_builder_rejects_bad_input(lib):
    pairs, ref, qer, *_ = load_golden("short8")
    bad = pairs[:100].copy()
    bad["len2"][3] = 0
    with pytest.raises(lib.BswError) as ei:
        lib.PackedBatch.from_pairs(bad, ref, qer)
    assert ei.value.code == -2
    q5 = qer.copy()
    q5[pairs["idq"][7]] = 5                                      # not a base code
    with pytest.raises(lib.BswError) as ei:
        lib.PackedBatch.from_pairs(pairs[:100].copy(), ref, q5)
    assert ei.value.code == -2
    empty = lib.PackedBatch.from_pairs(pairs[:0].copy(), ref, qer)
    assert empty.n_pairs == 0


def test_host_library_maps_no_cuda():
    """bench.py's reference arm builds its inputs through libbsw_host.so: the process must not map libbsw_b200.so
    (or any CUDA runtime) on that path."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import genomicsbench_b200 as gb\n"
        "cfg = gb.gen_named_config('short8', host_only=True)\n"
        "pairs, ref, qer = gb.gen_pairs(cfg, 0, 2000, host_only=True)\n"
        "b = gb.PackedBatch.from_pairs(pairs, ref, qer, host_only=True)\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libbsw_host.so' in maps\n"
        "assert 'libbsw_b200' not in maps and 'libcuda' not in maps and 'libcudart' not in maps, 'CUDA mapped'\n"
        "print(b.n_pairs)\n" % str(ROOT))
    out = subprocess.run([sys.executable, "-c", code], check=True, capture_output=True, text=True).stdout
    assert out.strip() == "2000"
