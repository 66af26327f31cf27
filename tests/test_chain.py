"""Seed -> pair construction and chain extension (tools/bwa/bwamem.c:632-822).

CPU: the restatement oracle/chain_oracle.c against the golden mem_alnreg_t lists that the reference's
own mem_chain2aln produced (tests/golden/make_golden_chain.py).  GPU: bsw_extend_chains against the
same goldens and against the oracle."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle.pyoracle import CHAIN_REG_FIELDS, CHAIN_SEED_DTYPE, make_params

CHAIN_CASES = sorted(p.stem for p in (GOLDEN_DIR / "chain").glob("*.npz"))


def load_chain_case(name):
    z = np.load(GOLDEN_DIR / "chain" / f"{name}.npz")
    G = z["genome"]
    D = np.concatenate([G, (3 - G[::-1]).astype(np.uint8)])        # both strands, bntseq.c:bns_get_seq
    a, b, o_del, e_del, o_ins, e_ins, zdrop, clip5, clip3, w = (int(v) for v in z["params"])
    seeds = np.zeros(len(z["seeds"]), dtype=CHAIN_SEED_DTYPE)
    for k, f in enumerate(("rbeg", "qbeg", "len", "score")):
        seeds[f] = z["seeds"][:, k]
    return dict(D=D, l_pac=len(G), query=z["query"], query_off=z["query_off"], l_query=z["l_query"], seeds=seeds,
                chain_first=z["chain_first"], chain_n=z["chain_n"], regs=z["regs"], reg_n=z["reg_n"],
                chain_read=z["chain_read"],
                P=dict(match=a, mismatch=b, o_del=o_del, e_del=e_del, o_ins=o_ins, e_ins=e_ins, zdrop=zdrop),
                clip5=clip5, clip3=clip3, w=w)


@pytest.mark.parametrize("case", CHAIN_CASES)
def test_chain_oracle_matches_reference_golden(oracle, case):
    c = load_chain_case(case)
    # ksw_extend2's z-drop rule is the scalar one (ksw.c:462-470): zdrop_mode = 1
    P = make_params(end_bonus=c["clip5"], ambig=-1, zdrop_mode=1, **c["P"])
    pos = 0
    prior = None
    for k in range(len(c["chain_n"])):
        sd = c["seeds"][c["chain_first"][k]: c["chain_first"][k] + c["chain_n"][k]]
        lq = int(c["l_query"][k])
        q = c["query"][c["query_off"][k]: c["query_off"][k] + lq]
        r0, r1 = oracle.chain_window(P, c["w"], c["l_pac"], sd, lq)
        if k == 0 or c["chain_read"][k] != c["chain_read"][k - 1]:
            prior = None                                            # a new read: a new mem_alnreg_v
        got = oracle.chain(P, c["w"], c["clip5"], c["clip3"], 2, q, sd, r0, r1, c["D"][r0:r1], prior)
        prior = got if prior is None else np.concatenate([prior, got])
        want = c["regs"][pos: pos + c["reg_n"][k]]
        pos += int(c["reg_n"][k])
        assert len(got) == len(want), f"chain {k}: {len(got)} regions, reference made {len(want)}"
        gm = np.stack([got[f] for f in CHAIN_REG_FIELDS], axis=1).astype(np.int64)
        assert np.array_equal(gm, want), f"chain {k}"
    assert pos == len(c["regs"])


def _build_batch(lib, eng, c):
    """Chains of a golden case in the C ABI's layout: windows from bsw_chain_window, window bytes
    concatenated into one reference buffer."""
    n = len(c["chain_n"])
    chains = np.zeros(n, dtype=lib.CHAIN_DTYPE)
    seeds = np.zeros(len(c["seeds"]), dtype=lib.SEED_DTYPE)
    for f in ("rbeg", "qbeg", "len", "score"):
        seeds[f] = c["seeds"][f]
    parts, off = [], 0
    for k in range(n):
        first, ns, lq = int(c["chain_first"][k]), int(c["chain_n"][k]), int(c["l_query"][k])
        r0, r1 = eng.chain_window(c["w"], c["l_pac"], seeds[first: first + ns], lq)
        same = int(k > 0 and c["chain_read"][k] == c["chain_read"][k - 1])
        chains[k] = (first, ns, lq, int(c["query_off"][k]), r0, r1, off, same, 0)
        parts.append(c["D"][r0:r1]); off += r1 - r0
    return chains, seeds, np.ascontiguousarray(c["query"]), np.ascontiguousarray(np.concatenate(parts))


def test_chain_window_matches_oracle(lib, oracle):
    """bsw_chain_window is host arithmetic (no device): same windows as the restatement of
    bwamem.c:643-659, including the strand-boundary rule."""
    from genomicsbench_b200 import default_params
    import ctypes as C
    L = lib.load_library()
    for case in CHAIN_CASES:
        c = load_chain_case(case)
        P = make_params(end_bonus=c["clip5"], zdrop_mode=1, **c["P"])
        bp = default_params()
        for k, v in c["P"].items():
            setattr(bp, k, v)
        seeds = np.zeros(len(c["seeds"]), dtype=lib.SEED_DTYPE)
        for f in ("rbeg", "qbeg", "len", "score"):
            seeds[f] = c["seeds"][f]
        crossed = 0
        for k in range(len(c["chain_n"])):
            first, ns, lq = int(c["chain_first"][k]), int(c["chain_n"][k]), int(c["l_query"][k])
            want = oracle.chain_window(P, c["w"], c["l_pac"], c["seeds"][first: first + ns], lq)
            r0, r1 = C.c_int64(0), C.c_int64(0)
            sd = np.ascontiguousarray(seeds[first: first + ns])
            assert L.bsw_chain_window(C.byref(bp), c["w"], c["l_pac"], sd.ctypes.data, ns, lq, C.byref(r0), C.byref(r1)) == 0
            assert (r0.value, r1.value) == want
            crossed += r0.value == c["l_pac"] or r1.value == c["l_pac"]
        assert crossed > 0, "no chain exercised the strand-boundary rule"


@pytest.mark.gpu
@pytest.mark.parametrize("case", CHAIN_CASES)
def test_extend_chains_matches_reference_golden(lib, case):
    """bsw_extend_chains (GPU extensions, batched over chains) == the mem_alnreg_t lists of the
    reference's own mem_chain2aln, bit for bit and in the reference's order."""
    c = load_chain_case(case)
    with lib.Engine(end_bonus=c["clip5"], zdrop_mode=lib.BSW_ZDROP_SCALAR, **c["P"]) as eng:
        chains, seeds, query, ref = _build_batch(lib, eng, c)
        regs, count = eng.extend_chains(chains, seeds, query, ref, c["w"], c["clip5"], c["clip3"], 2)
        st = eng.stats()
    assert np.array_equal(count, c["reg_n"])
    got = np.concatenate([regs[int(ch["seed_first"]): int(ch["seed_first"]) + int(n)] for ch, n in zip(chains, count)])
    gm = np.stack([got[f] for f in lib.ALNREG_FIELDS], axis=1).astype(np.int64)
    bad = np.flatnonzero((gm != c["regs"]).any(axis=1))
    assert len(bad) == 0, f"{len(bad)} regions differ, first {bad[:5]}: got {gm[bad[:2]]} want {c['regs'][bad[:2]]}"
    assert st["kernel_launches"] > 0 and st["cells_effective"] > 0


@pytest.mark.gpu
def test_extend_chains_vector_zdrop_equals_scalar_for_unit_extension(lib):
    """With e_del = e_ins = 1 and zdrop > 0 the vector z-drop rule (the engine's default mode) and
    ksw_extend2's coincide: the default engine gives the golden regions too."""
    c = load_chain_case("chain_default")
    with lib.Engine(end_bonus=c["clip5"], **c["P"]) as eng:
        chains, seeds, query, ref = _build_batch(lib, eng, c)
        regs, count = eng.extend_chains(chains, seeds, query, ref, c["w"], c["clip5"], c["clip3"], 2)
        got = np.concatenate([regs[int(ch["seed_first"]): int(ch["seed_first"]) + int(n)] for ch, n in zip(chains, count)])
        gm = np.stack([got[f] for f in lib.ALNREG_FIELDS], axis=1).astype(np.int64)
        assert np.array_equal(gm, c["regs"])
        # error behaviour: clipping penalties must match the engine's end_bonus
        with pytest.raises(lib.BswError) as ei:
            eng.extend_chains(chains, seeds, query, ref, c["w"], c["clip5"] + 1, c["clip3"], 2)
        assert ei.value.code == -1
        # empty batch
        r, n = eng.extend_chains(chains[:0], seeds[:0], query, ref, c["w"], c["clip5"], c["clip3"], 2)
        assert len(r) == 0 and len(n) == 0


@pytest.mark.gpu
def test_extend_chains_result_independent_of_batch_composition(lib):
    """Every multi-chain read of chain_multi submitted ALONE gives the golden regions: a chain that waits for
    its read's earlier chains must still run when those only skip contained seeds and no other read of the
    batch keeps the round loop alive (mem_chain2aln always processes every chain, bwamem.c:1105-1112)."""
    c = load_chain_case("chain_multi")
    read = c["chain_read"]
    reg_first = np.concatenate([[0], np.cumsum(c["reg_n"])])
    with lib.Engine(end_bonus=c["clip5"], zdrop_mode=lib.BSW_ZDROP_SCALAR, **c["P"]) as eng:
        chains, seeds, query, ref = _build_batch(lib, eng, c)
        starts = [k for k in range(len(read)) if k == 0 or read[k] != read[k - 1]]
        n_multi = 0
        for gi, k0 in enumerate(starts):
            k1 = starts[gi + 1] if gi + 1 < len(starts) else len(read)
            if k1 - k0 < 2:
                continue
            n_multi += 1
            regs, count = eng.extend_chains(chains[k0:k1], seeds, query, ref, c["w"], c["clip5"], c["clip3"], 2)
            assert np.array_equal(count, c["reg_n"][k0:k1]), f"read {read[k0]}: counts {count} want {c['reg_n'][k0:k1]}"
            for j, k in enumerate(range(k0, k1)):
                first = int(chains[k]["seed_first"])
                got = regs[first: first + int(count[j])]
                gm = np.stack([got[f] for f in lib.ALNREG_FIELDS], axis=1).astype(np.int64).reshape(-1, len(lib.ALNREG_FIELDS))
                assert np.array_equal(gm, c["regs"][reg_first[k]: reg_first[k + 1]]), f"read {read[k0]} chain {k}"
        assert n_multi > 100
