"""Banded global alignment + CIGAR (tools/bwa/ksw.c:489-606, ksw_global2).

CPU: oracle/global_oracle.c against the golden scores / CIGARs the reference's own ksw_global2 produced
(tests/golden/make_golden_global.py) and, where the reference tree is mounted, against live calls.
GPU: bsw_global against the same goldens."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle.pyoracle import KswReference, make_params

GLOBAL_CASES = sorted(p.stem for p in (GOLDEN_DIR / "global").glob("*.npz"))


def load_global_case(name):
    z = np.load(GOLDEN_DIR / "global" / f"{name}.npz")
    o_del, e_del, o_ins, e_ins, match, mismatch, ambig = (int(v) for v in z["params"])
    qoff = np.concatenate([[0], np.cumsum(z["len2"])]).astype(np.int64)
    toff = np.concatenate([[0], np.cumsum(z["len1"])]).astype(np.int64)
    coff = np.concatenate([[0], np.cumsum(z["n_cigar"])]).astype(np.int64)
    return dict(z=z, qoff=qoff, toff=toff, coff=coff,
                P=dict(o_del=o_del, e_del=e_del, o_ins=o_ins, e_ins=e_ins, match=match, mismatch=mismatch, ambig=ambig))


@pytest.mark.parametrize("case", GLOBAL_CASES)
def test_global_oracle_matches_reference_golden(oracle, case):
    c = load_global_case(case); z = c["z"]
    P = make_params(**c["P"])
    for k in range(len(z["w"])):
        q = z["query"][c["qoff"][k]: c["qoff"][k + 1]]; t = z["target"][c["toff"][k]: c["toff"][k + 1]]
        score, cig = oracle.global_align(P, q, t, int(z["w"][k]))
        assert score == int(z["score"][k]), f"pair {k}"
        assert np.array_equal(cig, z["cigar"][c["coff"][k]: c["coff"][k + 1]]), f"pair {k}"
        # a CIGAR consumes exactly the two sequences
        ln, op = cig >> 4, cig & 0xf
        assert ln[(op == 0) | (op == 1)].sum() == len(q) and ln[(op == 0) | (op == 2)].sum() == len(t)


def test_global_oracle_matches_live_reference(oracle):
    if not KswReference.available():
        pytest.skip("oracle/_ref/libkswref.so not built (needs /root/reference)")
    K = KswReference()
    rng = np.random.default_rng(0xB5B20405)
    for params in (dict(), dict(o_del=3, e_del=2, o_ins=7, e_ins=3, match=3, mismatch=5)):
        P = make_params(**params)
        for _ in range(150):
            q = rng.integers(0, 5, int(rng.integers(1, 120))).astype(np.uint8)
            t = rng.integers(0, 5, int(rng.integers(max(1, len(q) - 15), len(q) + 16))).astype(np.uint8)
            w = abs(len(q) - len(t)) + int(rng.integers(0, 12))
            a, b = oracle.global_align(P, q, t, w), K.global_align(P, q, t, w)
            assert a[0] == b[0] and np.array_equal(a[1], b[1])


def _pairs_of(lib, c):
    z = c["z"]
    n = len(z["w"])
    pairs = np.zeros(n, dtype=lib.SEQPAIR_DTYPE)
    pairs["len1"], pairs["len2"] = z["len1"], z["len2"]
    pairs["idr"], pairs["idq"] = c["toff"][:-1], c["qoff"][:-1]
    pairs["h0"] = 1
    return pairs, np.ascontiguousarray(z["target"]), np.ascontiguousarray(z["query"])


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["", "2w", "1"])
@pytest.mark.parametrize("case", GLOBAL_CASES)
def test_bsw_global_matches_reference_golden(lib, case, kernel, monkeypatch):
    """kernel: "" = the default (second kernel, csrc/bsw_global2.cuh, 16-bit slots), "2w" = its 64-bit slots, "1" = the
    first kernel (csrc/bsw_global.cuh), which stays the route for what the second one's domain excludes."""
    if kernel:
        monkeypatch.setenv("BSW_GLOBAL_KERNEL", kernel)
    else:
        monkeypatch.delenv("BSW_GLOBAL_KERNEL", raising=False)
    c = load_global_case(case); z = c["z"]
    pairs, ref, qer = _pairs_of(lib, c)
    with lib.Engine(**c["P"]) as eng:
        score, cigar, off = eng.global_align(pairs, ref, qer, z["w"])
        st = eng.stats()
    assert np.array_equal(score, z["score"])
    assert np.array_equal(np.diff(off), z["n_cigar"])
    assert np.array_equal(cigar, z["cigar"])
    assert st["kernel_launches"] >= 2 and st["cells_effective"] > 0
    # which form ran: the second gathers every query 8-aligned and every target 4-aligned, the first packs them tight
    l1, l2 = z["len1"].astype(np.int64), z["len2"].astype(np.int64)
    tight = 40 * len(l1) + int(l1.sum() + l2.sum())            # 40 = sizeof(GlobalDesc)
    padded = 40 * len(l1) + int(((l1 + 3) // 4 * 4).sum() + ((l2 + 7) // 8 * 8).sum())
    assert st["h2d_bytes"] == (tight if kernel == "1" else padded)


@pytest.mark.gpu
def test_bsw_global_outside_the_second_kernels_domain(lib, oracle, monkeypatch):
    """Rows beyond shared memory (a band as wide as long sequences) run the first kernel with its rows in HBM; values
    beyond 16 bits run the second kernel with 64-bit slots."""
    monkeypatch.delenv("BSW_GLOBAL_KERNEL", raising=False)
    rng = np.random.default_rng(0xB5B20409)
    for params, n, qlen, w in ((dict(), 3, 2000, 1500), (dict(match=9, mismatch=9), 3, 4000, 12)):
        P = make_params(**params)
        qs, ts = [], []
        for _ in range(n):
            q = rng.integers(0, 4, qlen).astype(np.uint8)
            t = q.copy(); t[::17] = (t[::17] + 1) % 4; t = np.delete(t, [5, 40])
            qs.append(q); ts.append(t)
        pairs = np.zeros(n, dtype=lib.SEQPAIR_DTYPE)
        pairs["len1"], pairs["len2"] = [len(t) for t in ts], [len(q) for q in qs]
        pairs["idr"] = np.concatenate([[0], np.cumsum(pairs["len1"])[:-1]]); pairs["idq"] = np.concatenate([[0], np.cumsum(pairs["len2"])[:-1]])
        pairs["h0"] = 1
        with lib.Engine(**params) as eng:
            score, cigar, off = eng.global_align(pairs, np.concatenate(ts), np.concatenate(qs), w)
            st = eng.stats()
        l1, l2 = pairs["len1"].astype(np.int64), pairs["len2"].astype(np.int64)
        first_kernel = st["h2d_bytes"] == 40 * n + int(l1.sum() + l2.sum())
        assert first_kernel == (w == 1500)
        for k in range(n):
            sc, cg = oracle.global_align(P, qs[k], ts[k], w)
            assert sc == score[k] and np.array_equal(cg, cigar[off[k]: off[k + 1]])


@pytest.mark.gpu
def test_bsw_global_edges_and_errors(lib, oracle):
    """Empty batch, one band for all pairs, order invariance, several chunks, domain errors."""
    c = load_global_case("global_default"); z = c["z"]
    pairs, ref, qer = _pairs_of(lib, c)
    with lib.Engine(**c["P"]) as eng:
        s0, c0, o0 = eng.global_align(pairs[:0], ref, qer, 10)
        assert len(s0) == 0 and len(c0) == 0 and list(o0) == [0]
        wide = int(np.abs(pairs["len1"] - pairs["len2"]).max()) + 5
        score, cigar, off = eng.global_align(pairs, ref, qer, wide)            # scalar band
        P = make_params(**c["P"])
        for k in range(0, len(pairs), 37):
            q = qer[pairs["idq"][k]: pairs["idq"][k] + pairs["len2"][k]]; t = ref[pairs["idr"][k]: pairs["idr"][k] + pairs["len1"][k]]
            sc, cg = oracle.global_align(P, q, t, wide)
            assert sc == score[k] and np.array_equal(cg, cigar[off[k]: off[k + 1]])
        # a band too wide for the shared-memory rows (W = 2 w + 2 columns per thread): the kernel's
        # global-scratch form runs instead
        s300, c300, o300 = eng.global_align(pairs[:200], ref, qer, 300)
        for k in range(0, 200, 7):
            q = qer[pairs["idq"][k]: pairs["idq"][k] + pairs["len2"][k]]; t = ref[pairs["idr"][k]: pairs["idr"][k] + pairs["len1"][k]]
            sc, cg = oracle.global_align(P, q, t, 300)
            assert sc == s300[k] and np.array_equal(cg, c300[o300[k]: o300[k + 1]])
        # w = 0: the diagonal only (equal lengths)
        eq = np.flatnonzero(pairs["len1"] == pairs["len2"])[:50]
        s0w, c0w, o0w = eng.global_align(pairs[eq], ref, qer, 0)
        for j, k in enumerate(eq):
            q = qer[pairs["idq"][k]: pairs["idq"][k] + pairs["len2"][k]]; t = ref[pairs["idr"][k]: pairs["idr"][k] + pairs["len1"][k]]
            sc, cg = oracle.global_align(P, q, t, 0)
            assert sc == s0w[j] and np.array_equal(cg, c0w[o0w[j]: o0w[j + 1]])
        perm = np.random.default_rng(3).permutation(len(pairs))
        s2, c2, o2 = eng.global_align(pairs[perm], ref, qer, z["w"][perm])
        assert np.array_equal(s2, z["score"][perm])
        for j in (0, 5, len(perm) - 1):
            k = perm[j]
            assert np.array_equal(c2[o2[j]: o2[j + 1]], z["cigar"][c["coff"][k]: c["coff"][k + 1]])
        big = np.tile(pairs, 300)                                                # 210 000 alignments: several chunks
        s3, c3, o3 = eng.global_align(big, ref, qer, np.tile(z["w"], 300))
        assert np.array_equal(s3, np.tile(z["score"], 300)) and np.array_equal(c3, np.tile(z["cigar"], 300))
        # caller-owned result arrays, reused across calls; a cigar buffer one entry short is refused
        out = (np.zeros(len(big), np.int32), np.zeros(len(big), np.int32), np.zeros(len(c3), np.uint32), np.zeros(len(big) + 1, np.int64))
        for _ in range(2):
            s4, c4, o4 = eng.global_align(big, ref, qer, np.tile(z["w"], 300), out=out)
            assert np.array_equal(s4, s3) and np.array_equal(c4, c3) and np.array_equal(o4, o3)
        with pytest.raises(lib.BswError) as ei:
            eng.global_align(big, ref, qer, np.tile(z["w"], 300), out=(out[0], out[1], out[2][:-1], out[3]))
        assert ei.value.code == -1
        bad = pairs[:4].copy(); bad["len1"][2] = bad["len2"][2] + 50
        with pytest.raises(lib.BswError) as ei:
            eng.global_align(bad, ref, qer, 10)
        assert ei.value.code == -2
