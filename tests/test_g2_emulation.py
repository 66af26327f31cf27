"""bsw_global's second kernel (csrc/bsw_global2.cuh: rows in band coordinates, 16-bit slots, 4-bit directions) and its
chunk planner (csrc/bsw_global_plan.h) executed on the CPU (tests/emu/g2_emu.cu) and compared with the goldens the
reference's own ksw_global2 produced (tools/bwa/ksw.c:502-606) and with the oracle on seeded sweeps -- scores and
CIGARs bit for bit, in both slot widths, across chunk cuts and launch classes."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from oracle.pyoracle import make_params
from conftest import GOLDEN_DIR
from test_global import GLOBAL_CASES, load_global_case

# goldens of the reference's ksw_global2 that only this file runs (tests/golden/make_golden_global.py, CASES_EMU)
EMU_CASES = sorted(p.stem for p in (GOLDEN_DIR / "global_emu").glob("*.npz"))


def load_emu_case(name):
    z = np.load(GOLDEN_DIR / "global_emu" / f"{name}.npz")
    o_del, e_del, o_ins, e_ins, match, mismatch, ambig = (int(v) for v in z["params"])
    return z, dict(o_del=o_del, e_del=e_del, o_ins=o_ins, e_ins=e_ins, match=match, mismatch=mismatch, ambig=ambig)

SEQPAIR = np.dtype([("idr", "<i8"), ("idq", "<i8"), ("id", "<i8"), ("len1", "<i4"), ("len2", "<i4"), ("h0", "<i4"),
                    ("seqid", "<i4"), ("regid", "<i4"), ("score", "<i4"), ("tle", "<i4"), ("gtle", "<i4"), ("qle", "<i4"),
                    ("gscore", "<i4"), ("max_off", "<i4"), ("pad", "<i4")])


@pytest.fixture(scope="module")
def emu():
    from emu.build import build
    lib = C.CDLL(str(build(name="g2")))
    lib.g2_emu_global.restype = C.c_int
    lib.g2_emu_global.argtypes = [C.c_void_p] * 4 + [C.c_longlong, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong] + \
        [C.c_void_p] * 3 + [C.c_longlong, C.c_void_p, C.c_void_p]
    return lib


def test_seqpair_layout_matches_header():
    import genomicsbench_b200._lib as L
    assert SEQPAIR.itemsize == 72 == L.SEQPAIR_DTYPE.itemsize
    for f in ("idr", "idq", "len1", "len2"):
        assert SEQPAIR.fields[f][1] == L.SEQPAIR_DTYPE.fields[f][1]


def run_emu(lib, P: dict, len1, len2, target, query, w, rows=0, caps_m=0, caps_z=0):
    n = len(len1)
    pairs = np.zeros(n, dtype=SEQPAIR)
    pairs["len1"], pairs["len2"] = len1, len2
    pairs["idr"] = np.concatenate([[0], np.cumsum(len1)[:-1]]) if n else []
    pairs["idq"] = np.concatenate([[0], np.cumsum(len2)[:-1]]) if n else []
    prm = np.array([P["o_del"], P["e_del"], P["o_ins"], P["e_ins"], P["match"], P["mismatch"], P["ambig"]], dtype=np.int32)
    w = np.ascontiguousarray(np.broadcast_to(np.asarray(w, dtype=np.int32), (n,)))
    score = np.zeros(n, np.int32); ncig = np.zeros(n, np.int32); off = np.zeros(n + 1, np.int64)
    cap = int((pairs["len1"].astype(np.int64) + pairs["len2"]).sum()) + 1
    cigar = np.zeros(cap, np.uint32); info = np.zeros(5, np.int64)
    target = np.ascontiguousarray(target, dtype=np.uint8); query = np.ascontiguousarray(query, dtype=np.uint8)
    rc = lib.g2_emu_global(prm.ctypes.data, pairs.ctypes.data, target.ctypes.data, query.ctypes.data, n, w.ctypes.data, rows,
                           caps_m, caps_z, score.ctypes.data, ncig.ctypes.data, cigar.ctypes.data, cap, off.ctypes.data,
                           info.ctypes.data)
    return rc, score, ncig, cigar[: off[n]], off, info


@pytest.mark.parametrize("rows", [16, 32])
@pytest.mark.parametrize("case", GLOBAL_CASES)
def test_emulated_kernel_matches_reference_golden(emu, case, rows):
    c = load_global_case(case); z = c["z"]
    rc, score, ncig, cigar, off, info = run_emu(emu, c["P"], z["len1"], z["len2"], z["target"], z["query"], z["w"], rows=rows)
    assert rc == 0
    assert np.array_equal(score, z["score"])
    assert np.array_equal(ncig, z["n_cigar"]) and np.array_equal(np.diff(off), z["n_cigar"])
    assert np.array_equal(cigar, z["cigar"])
    assert info[0] == 1 and info[1] >= 1 and info[2] == (rows == 16) and info[4] == 0


def test_emulated_default_choice_and_chunk_cuts(emu):
    """The engine's own choice of slot width; chunks cut by the alignment cap and by the direction-matrix cap (whole
    ranges, and alignment by alignment when a range alone is too much) all give the same answer."""
    c = load_global_case("global_default"); z = c["z"]
    tile = 7                                   # 4 900 alignments: three ranges of the planner
    args = (c["P"], np.tile(z["len1"], tile), np.tile(z["len2"], tile), np.tile(z["target"], tile), np.tile(z["query"], tile),
            np.tile(z["w"], tile))
    want = (np.tile(z["score"], tile), np.tile(z["cigar"], tile))
    for kw, chunks in ((dict(), 1), (dict(caps_m=2048), 3), (dict(caps_m=3000), 2), (dict(caps_m=1), 4900), (dict(caps_z=1 << 21), None), (dict(caps_z=40000), None)):
        rc, score, ncig, cigar, off, info = run_emu(emu, *args, **kw)
        assert rc == 0 and info[2] == 1
        assert np.array_equal(score, want[0]) and np.array_equal(cigar, want[1]), kw
        if chunks:
            assert info[0] == chunks
        else:
            assert info[0] > 3
    # effective cells = the sum over rows of the window width
    rc, *_, info = run_emu(emu, *args)
    w = z["w"].astype(np.int64); q = z["len2"].astype(np.int64); t = z["len1"].astype(np.int64)
    cells = sum(int(sum(min(i + wv + 1, ql) - max(i - wv, 0) for i in range(tl))) for ql, tl, wv in zip(q, t, w))
    assert info[3] == cells * tile


@pytest.mark.parametrize("rows", [16, 32])
def test_emulated_kernel_matches_oracle_on_seeded_sweeps(emu, oracle, rows):
    """Random pairs over three parameter sets: wide bands against short queries (row_slots' second branch), w = 0,
    N on either side, length-1 sequences, bands wider than any sequence, free gap opens."""
    rng = np.random.default_rng(0xB5B20406 + rows)
    for params in (dict(), dict(o_del=3, e_del=2, o_ins=7, e_ins=3, match=3, mismatch=5, ambig=-2),
                   dict(o_del=0, e_del=1, o_ins=0, e_ins=1, match=2, mismatch=1, ambig=0)):
        P = make_params(**params)
        Pd = dict(o_del=P.o_del, e_del=P.e_del, o_ins=P.o_ins, e_ins=P.e_ins, match=P.match, mismatch=P.mismatch, ambig=P.ambig)
        qs, ts, ws = [], [], []
        for k in range(500):
            ql = int(rng.integers(1, 140)) if k % 5 else int(rng.integers(1, 9))
            q = rng.integers(0, 5 if k % 3 == 0 else 4, ql).astype(np.uint8)
            if k % 2:
                t = q.copy()
                for _ in range(int(rng.integers(0, 6))):
                    pos = int(rng.integers(0, len(t) + 1)); u = rng.random()
                    if u < 0.4 and len(t) > 1: t = np.delete(t, min(pos, len(t) - 1))
                    elif u < 0.8: t = np.insert(t, pos, rng.integers(0, 4))
                    else: t[min(pos, len(t) - 1)] = rng.integers(0, 5)
            else:
                t = rng.integers(0, 4, int(rng.integers(max(1, ql - 12), ql + 13))).astype(np.uint8)
            d = abs(len(q) - len(t))
            w = d + (0 if k % 7 == 0 else int(rng.integers(0, 12)) if k % 11 else int(rng.integers(100, 300)))
            qs.append(q); ts.append(t.astype(np.uint8)); ws.append(w)
        rc, score, ncig, cigar, off, info = run_emu(emu, Pd, [len(t) for t in ts], [len(q) for q in qs], np.concatenate(ts),
                                                    np.concatenate(qs), ws, rows=rows)
        assert rc == 0 and info[1] > 1 and info[4] == 0  # several launch classes; nothing left 16 bits
        for k in range(len(qs)):
            sc, cg = oracle.global_align(P, qs[k], ts[k], ws[k])
            assert sc == score[k], (params, k)
            assert np.array_equal(cg, cigar[off[k]: off[k + 1]]), (params, k)


def test_emulated_long_alignments_need_wide_slots(emu, oracle):
    """Values beyond 16 bits: the 16-bit slots are refused, the default choice runs 64-bit slots and is exact."""
    rng = np.random.default_rng(0xB5B20407)
    P = make_params(match=9, mismatch=9)
    Pd = dict(o_del=P.o_del, e_del=P.e_del, o_ins=P.o_ins, e_ins=P.e_ins, match=9, mismatch=9, ambig=P.ambig)
    q = rng.integers(0, 4, 4000).astype(np.uint8)
    t = q.copy(); t[::97] = (t[::97] + 1) % 4; t = np.delete(t, [500, 501, 2000])
    assert run_emu(emu, Pd, [len(t)], [len(q)], t, q, 10, rows=16)[0] == -1
    rc, score, ncig, cigar, off, info = run_emu(emu, Pd, [len(t)], [len(q)], t, q, 10)
    assert rc == 0 and info[2] == 0
    sc, cg = oracle.global_align(P, q, t, 10)
    assert sc == score[0] and sc > 32767 and np.array_equal(cg, cigar)


def test_emulated_16_bit_slots_at_the_edge_of_their_domain(emu, oracle):
    """rows16_ok's bound is what decides: the worst alignments it still admits (nothing matches, heavy penalties, the
    longest lengths) stay inside 16 bits and exact; one base longer and the default choice leaves the 16-bit slots."""
    rng = np.random.default_rng(0xB5B20408)
    P = make_params(o_del=40, e_del=9, o_ins=35, e_ins=11, match=2, mismatch=30, ambig=-25)
    Pd = dict(o_del=40, e_del=9, o_ins=35, e_ins=11, match=2, mismatch=30, ambig=-25)
    # lo = 30 * (len + 1) + 3 * 49 + 11 * (w + 2) <= 32000
    w = 20
    n = (32000 - 3 * 49 - 11 * (w + 2)) // 30 - 1
    for kind in range(3):
        q = np.zeros(n, np.uint8) if kind < 2 else rng.integers(0, 4, n).astype(np.uint8)
        t = np.ones(n - (w if kind == 1 else 0), np.uint8) if kind < 2 else (q[: n - 7] + 1) % 4
        rc, score, ncig, cigar, off, info = run_emu(emu, Pd, [len(t)], [len(q)], t, q, w)
        assert rc == 0 and info[2] == 1 and info[4] == 0
        sc, cg = oracle.global_align(P, q, t, w)
        assert sc == score[0] and np.array_equal(cg, cigar)
    q = np.zeros(n + 2, np.uint8); t = np.ones(n + 2, np.uint8)
    rc, score, ncig, cigar, off, info = run_emu(emu, Pd, [len(t)], [len(q)], t, q, w)
    assert rc == 0 and info[2] == 0
    sc, cg = oracle.global_align(P, q, t, w)
    assert sc == score[0] and np.array_equal(cg, cigar)


def test_emulated_domain_checks(emu):
    """Scores beyond a signed byte and rows beyond shared memory are refused (the engine runs the first kernel then)."""
    P = dict(o_del=6, e_del=1, o_ins=6, e_ins=1, match=1, mismatch=4, ambig=-1)
    q = np.zeros(40, np.uint8)
    assert run_emu(emu, dict(P, match=200), [40], [40], q, q, 3)[0] == -1
    assert run_emu(emu, P, [40], [40], q, q, 3000)[0] == 0          # a band wider than the sequences costs no more slots than the sequences
    q = np.zeros(3000, np.uint8)
    assert run_emu(emu, P, [3000], [3000], q, q, 2900)[0] == -1


@pytest.mark.parametrize("n,slices", [(1, 4), (5000, 1), (70000, 3), (300000, 16), (262144, 7)])
def test_planner_radix_sort_is_a_stable_sort(emu, n, slices):
    """The work order's parallel LSD radix sort (slices counted and scattered side by side) = std::stable_sort on the key bits."""
    emu.g2_emu_sort_check.restype = C.c_int
    emu.g2_emu_sort_check.argtypes = [C.c_longlong, C.c_int, C.c_ulonglong]
    assert emu.g2_emu_sort_check(n, slices, 0x9E3779B97F4A7C15 + n) == 0


def test_emulated_large_chunk_goes_through_the_sliced_sort(emu):
    """70 000 alignments in one chunk: the planner's sort runs with several slices, ranges of every pass are many."""
    c = load_global_case("global_default"); z = c["z"]
    tile = 100
    rc, score, ncig, cigar, off, info = run_emu(emu, c["P"], np.tile(z["len1"], tile), np.tile(z["len2"], tile), np.tile(z["target"], tile),
                                                np.tile(z["query"], tile), np.tile(z["w"], tile))
    assert rc == 0 and info[0] == 1 and info[4] == 0
    assert np.array_equal(score, np.tile(z["score"], tile)) and np.array_equal(cigar, np.tile(z["cigar"], tile))


def test_emulated_16_bit_domain_randomised(emu, oracle):
    """Random scoring schemes with the longest sequences rows16_ok still admits for them (up to 2 500 bases), built to
    push values down (nothing matches, N everywhere, one long gap) or up (everything matches): the 16-bit slots run,
    nothing leaves 16 bits, results are the oracle's."""
    rng = np.random.default_rng(0xB5B2040A)
    for trial in range(60):
        match = int(rng.integers(1, 25)); mismatch = int(rng.integers(1, 60)); ambig = -int(rng.integers(0, 50))
        o_del, o_ins = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        e_del, e_ins = int(rng.integers(1, 16)), int(rng.integers(1, 16))
        w = int(rng.integers(0, 40))
        worst = max(mismatch, -ambig); gap = max(o_del + e_del, o_ins + e_ins); ext = max(e_del, e_ins)
        n = min((32000 - 3 * gap - ext * (w + 2)) // worst, (32000 - gap) // match) - 1
        n = int(min(n, 2500))
        assert n > w + 2
        P = make_params(o_del=o_del, e_del=e_del, o_ins=o_ins, e_ins=e_ins, match=match, mismatch=mismatch, ambig=ambig)
        Pd = dict(o_del=o_del, e_del=e_del, o_ins=o_ins, e_ins=e_ins, match=match, mismatch=mismatch, ambig=ambig)
        kind = trial % 5
        if kind == 0:   q, t = np.zeros(n, np.uint8), np.ones(n, np.uint8)                    # nothing matches
        elif kind == 1: q, t = np.full(n, 4, np.uint8), np.full(n - w, 4, np.uint8)           # N everywhere, the longest gap the band allows
        elif kind == 2: q = rng.integers(0, 4, n).astype(np.uint8); t = q.copy()              # everything matches
        elif kind == 3: q = np.zeros(n - w, np.uint8); t = np.ones(n, np.uint8)               # nothing matches + gap on the other side
        else:           q = rng.integers(0, 5, n).astype(np.uint8); t = rng.integers(0, 5, n - int(rng.integers(0, w + 1))).astype(np.uint8)
        rc, score, ncig, cigar, off, info = run_emu(emu, Pd, [len(t)], [len(q)], t, q, w)
        assert rc == 0 and info[2] == 1 and info[4] == 0, (trial, Pd, n, w)
        sc, cg = oracle.global_align(P, q, t, w)
        assert sc == score[0] and np.array_equal(cg, cigar), (trial, Pd, n, w)


@pytest.mark.parametrize("rows", [16, 32])
@pytest.mark.parametrize("case", EMU_CASES)
def test_emulated_kernel_matches_more_reference_goldens(emu, oracle, case, rows):
    """Bands far wider than needed (up to 250 slots, many launch classes), heavy scores with a positive score against N,
    free gap opens: scores and CIGARs of the reference's own ksw_global2; the oracle is held to the same vectors."""
    z, P = load_emu_case(case)
    rc, score, ncig, cigar, off, info = run_emu(emu, P, z["len1"], z["len2"], z["target"], z["query"], z["w"], rows=rows)
    assert rc == 0 and info[4] == 0
    assert np.array_equal(score, z["score"]) and np.array_equal(ncig, z["n_cigar"]) and np.array_equal(cigar, z["cigar"])
    if rows == 16:
        Po = make_params(**P)
        qoff = np.concatenate([[0], np.cumsum(z["len2"])]); toff = np.concatenate([[0], np.cumsum(z["len1"])])
        coff = np.concatenate([[0], np.cumsum(z["n_cigar"])])
        for k in range(0, len(z["w"]), 3):
            sc, cg = oracle.global_align(Po, z["query"][qoff[k]: qoff[k + 1]], z["target"][toff[k]: toff[k + 1]], int(z["w"][k]))
            assert sc == int(z["score"][k]) and np.array_equal(cg, z["cigar"][coff[k]: coff[k + 1]])
