"""C-ABI surface: the library loads, exports every symbol include/bsw.h declares, the record
layouts match the reference, and the product path fails loudly without a GPU (no CPU fallback).
No DP compute happens here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "bsw.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bsw_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 18
    cdll = lib.load_library()
    for s in syms:
        assert hasattr(cdll, s), f"{s} declared in include/bsw.h but not exported"
    # and the ctypes table covers exactly the header
    from genomicsbench_b200._lib import ABI
    assert sorted(n for n, _, _ in ABI) == syms


def test_layouts(lib):
    from genomicsbench_b200._lib import BswParams, BswStats, BswGenConfig
    assert lib.SEQPAIR_DTYPE.itemsize == 72                       # bandedSWA.h:91-100
    assert lib.SEQPAIR_DTYPE.fields["len1"][1] == 24 and lib.SEQPAIR_DTYPE.fields["max_off"][1] == 64
    assert C.sizeof(BswParams) == 4 * (11 + 16 + 1 + 8)
    assert C.sizeof(BswGenConfig) == 8 + 8 + 4 * 8 + 8 + 8 + 4 * 8
    assert C.sizeof(BswStats) == 8 * 3 + 8 * 7 + 8 * 2 + 4 * 8


def test_default_params_are_bwa_defaults(lib):
    p = lib.default_params()
    # main_banded.cpp:49-53,250
    assert (p.match, p.mismatch, p.o_del, p.e_del, p.o_ins, p.e_ins) == (1, 4, 6, 1, 6, 1)
    assert (p.zdrop, p.end_bonus, p.ambig) == (100, 5, -1)
    assert lib.load_library().bsw_version().decode().endswith("sm_100a")


@pytest.mark.parametrize("bad", [dict(zdrop=0), dict(zdrop=-5), dict(e_del=0), dict(match=0), dict(zdrop=40000),
                                 dict(n_devices=99)])
def test_param_validation_rejects(lib, bad):
    with pytest.raises(lib.BswError) as ei:
        lib.Engine(**bad)
    assert ei.value.code == -1


def test_no_cpu_fallback(lib):
    """Without a device the engine must refuse to exist rather than compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.BswError) as ei:
        lib.Engine()
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)


def test_product_does_not_reference_oracle():
    """The oracle is test infrastructure: nothing in the package may import / link it."""
    pkg = ROOT / "genomicsbench_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")) + \
            list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.inl")):
        t = p.read_text()
        for word in ("pyoracle", "ksw_oracle", "chain_oracle", "global_oracle", "libbsw_oracle", "libbswref", "libbwamemref",
                     "libkswref"):
            assert word not in t, (p, word)


def test_header_compiles_as_c_and_layouts_match_the_mirrors(lib, tmp_path):
    """include/bsw.h is a C header (plain pointers and sizes): compile it with gcc -std=c99 and compare the
    sizes / offsets the C compiler sees with the numpy and ctypes mirrors the tests and the Python class use."""
    import shutil
    import subprocess
    from genomicsbench_b200._lib import (ALNREG_DTYPE, CHAIN_DTYPE, OUTSCORE_DTYPE, PAIR_DESC_DTYPE, SCORE16_DTYPE, SEED_DTYPE,
                                         BswChainOpt, BswGenConfig, BswPackedBatch, BswParams, BswStats)
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    src = tmp_path / "layout.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "bsw.h"
#define S(T) printf(#T " %zu\\n", sizeof(T))
#define O(T, F) printf(#T "." #F " %zu\\n", offsetof(T, F))
int main(void) {
    S(SeqPair); O(SeqPair, len1); O(SeqPair, h0); O(SeqPair, score); O(SeqPair, max_off);
    S(bsw_params); S(bsw_stats); O(bsw_stats, kernel_launches); O(bsw_stats, partitioned); S(bsw_gen_config);
    S(bsw_seed); O(bsw_seed, qbeg); O(bsw_seed, score);
    S(bsw_chain); O(bsw_chain, l_query); O(bsw_chain, rmax0); O(bsw_chain, ref_off); O(bsw_chain, same_read);
    S(bsw_alnreg); O(bsw_alnreg, qb); O(bsw_alnreg, truesc); O(bsw_alnreg, seedlen0);
    S(bsw_chain_opt);
    S(bsw_pair_desc); O(bsw_pair_desc, len2); O(bsw_pair_desc, h0); O(bsw_pair_desc, flags);
    S(bsw_packed_batch); O(bsw_packed_batch, q2_words); O(bsw_packed_batch, raw_r_bytes); O(bsw_packed_batch, ordered);
    S(OutScore); O(OutScore, qle); S(bsw_score16); O(bsw_score16, max_off);
    return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    got = {k: int(v) for k, v in got.items()}
    sp = lib.SEQPAIR_DTYPE
    assert got["SeqPair"] == sp.itemsize == 72
    for f in ("len1", "h0", "score", "max_off"):
        assert got[f"SeqPair.{f}"] == sp.fields[f][1]
    assert got["bsw_params"] == C.sizeof(BswParams) and got["bsw_stats"] == C.sizeof(BswStats)
    assert got["bsw_stats.kernel_launches"] == BswStats.kernel_launches.offset
    assert got["bsw_stats.partitioned"] == BswStats.partitioned.offset
    assert got["bsw_gen_config"] == C.sizeof(BswGenConfig) and got["bsw_chain_opt"] == C.sizeof(BswChainOpt)
    assert got["bsw_packed_batch"] == C.sizeof(BswPackedBatch)
    for f in ("q2_words", "raw_r_bytes", "ordered"):
        assert got[f"bsw_packed_batch.{f}"] == getattr(BswPackedBatch, f).offset
    for name, dt, fields in (("bsw_pair_desc", PAIR_DESC_DTYPE, ("len2", "h0", "flags")), ("OutScore", OUTSCORE_DTYPE, ("qle",)),
                             ("bsw_score16", SCORE16_DTYPE, ("max_off",)),
                             ("bsw_seed", SEED_DTYPE, ("qbeg", "score")),
                             ("bsw_chain", CHAIN_DTYPE, ("l_query", "rmax0", "ref_off", "same_read")),
                             ("bsw_alnreg", ALNREG_DTYPE, ("qb", "truesc", "seedlen0"))):
        assert got[name] == dt.itemsize, name
        for f in fields:
            assert got[f"{name}.{f}"] == dt.fields[f][1], (name, f)
