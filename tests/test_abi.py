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
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cuh")):
        t = p.read_text()
        assert "pyoracle" not in t and "ksw_oracle" not in t and "libbsw_oracle" not in t and "libbswref" not in t, p
