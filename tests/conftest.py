"""Shared fixtures.  Tests marked `gpu` need a B200; everything else runs on CPU."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN_DIR = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; without a device they are skipped, not failed
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The product library, built in-tree (nvcc cross-compiles without a GPU)."""
    from genomicsbench_b200 import build as _b
    _b.build()
    import genomicsbench_b200 as gb
    gb.load_library()
    return gb


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.pyoracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libbswref.so not built (needs /root/reference)")
    return Reference()


GOLDEN_CASES = sorted(p.stem for p in GOLDEN_DIR.glob("*.npz"))


def load_golden(name: str):
    """-> (pairs SeqPair array, seq_ref, seq_qer, w, params dict, expect[n,6], scalar[n,6])"""
    from genomicsbench_b200 import SEQPAIR_DTYPE
    z = np.load(GOLDEN_DIR / f"{name}.npz")
    n = len(z["len1"])
    pairs = np.zeros(n, dtype=SEQPAIR_DTYPE)
    for f in ("len1", "len2", "h0", "idr", "idq"):
        pairs[f] = z[f]
    pairs["id"] = np.arange(n)
    for f in ("score", "qle", "tle", "gtle", "gscore", "max_off", "seqid", "regid"):
        pairs[f] = -1
    keys = ("o_del", "e_del", "o_ins", "e_ins", "zdrop", "end_bonus", "match", "mismatch", "ambig")
    params = {k: int(v) for k, v in zip(keys, z["params"])}
    return (pairs, np.ascontiguousarray(z["seq_ref"]), np.ascontiguousarray(z["seq_qer"]), int(z["w"]),
            params, z["expect"].astype(np.int32), z["scalar"].astype(np.int32))


def results_matrix(pairs) -> np.ndarray:
    from genomicsbench_b200 import RESULT_FIELDS
    return np.stack([pairs[f] for f in RESULT_FIELDS], axis=1).astype(np.int32)
