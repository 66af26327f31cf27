"""The reference's own driver (benchmarks/bsw/main_banded.cpp, unmodified) running on the B200
engine through the header-compatible BandedPairWiseSW class.  The
binary is built by oracle/Makefile where /root/reference exists and travels prebuilt."""
import re
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
BIN = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "bsw_main_b200"


@pytest.mark.skipif(not BIN.exists(), reason="oracle/_ref/bsw_main_b200 not built (no reference tree at build time)")
@pytest.mark.parametrize("batch,threads", [(None, 1), (512, 1), (512, 4)])
def test_reference_driver_runs_on_the_engine(lib, tmp_path, batch, threads):
    """-t 4: the driver's OpenMP loop (main_banded.cpp:279-291) -- one BandedPairWiseSW, i.e. one engine, per
    thread, all calling getScores16 on the same GPU at once."""
    cfg = lib.gen_named_config("small")                       # 151 bp / ~251 bp: fits the stock loader's slots
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 6000)
    path = tmp_path / "pairs.txt"
    lib.write_pairs_file(str(path), pairs, ref, qer)
    cmd = [str(BIN), "-pairs", str(path), "-t", str(threads)] + (["-b", str(batch)] if batch else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 1, res.stderr                    # main_banded.cpp:352 returns 1 on success
    assert re.search(r"Overall SW cycles", res.stdout), res.stdout
    assert "bsw_b200:" not in res.stderr                      # the shim's failure prefix


# ---------------------------------------------------------------------------------------------------------------
# results of the C++ drop-in (not only its exit code)
# ---------------------------------------------------------------------------------------------------------------
import struct
import numpy as np
from conftest import load_golden

CHECK = Path(__file__).resolve().parent / "cxx" / "dropin_check"


def _build_check():
    if CHECK.exists():
        return True
    import shutil
    root = Path(__file__).resolve().parent.parent
    if not shutil.which("g++"):
        return False
    cmd = ["g++", "-O2", "-fopenmp", "-I", str(root / "include"), str(CHECK) + ".cpp", "-L", str(root / "genomicsbench_b200" / "lib"),
           "-lbsw_b200", "-Wl,-rpath," + str(root / "genomicsbench_b200" / "lib"), "-o", str(CHECK)]
    return subprocess.run(cmd).returncode == 0


@pytest.mark.parametrize("coalesce", [0, 8192])
@pytest.mark.parametrize("case", ["small_151bp", "short8", "with_N", "gap_e2_e3_z20", "large_mix"])
def test_cxx_dropin_class_results(lib, tmp_path, case, coalesce):
    """BandedPairWiseSW of include/bandedSWA.h called from C++ (tests/cxx/dropin_check.cpp): getScores16, getScores8
    and the -t 4 -b 512 shape (instances created lazily inside the OpenMP region) give the golden getScores16
    fields; scalarBandedSWAWrapper and scalarBandedSWA (scalar z-drop rule, `mat` consulted for the ambiguity score)
    give the golden scalarBandedSWA fields.  coalesce = BSW_SHIM_COALESCE: small calls of all instances through the
    shared coalescing queue (bsw_extend_async) or every instance on its own engine (the default)."""
    import os
    if not _build_check():
        pytest.skip("no C++ compiler for tests/cxx/dropin_check.cpp")
    pairs, ref, qer, w, P, expect, scalar = load_golden(case)
    n = len(pairs)
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<11i", n, w, P["o_del"], P["e_del"], P["o_ins"], P["e_ins"], P["zdrop"], P["end_bonus"],
                            P["match"], P["mismatch"], P["ambig"]))
        f.write(struct.pack("<2q", len(ref), len(qer)))
        f.write(pairs.tobytes()); f.write(ref.tobytes()); f.write(qer.tobytes())
    res = subprocess.run([str(CHECK), str(inp), str(outp)], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, BSW_SHIM_COALESCE=str(coalesce)))
    assert res.returncode == 0, res.stderr
    out = np.fromfile(outp, dtype=np.int32).reshape(5, n, 6)
    in8 = (pairs["len1"] < 128) & (pairs["len2"] < 128) & (pairs["h0"] + pairs["len2"] * P["match"] <= 127)
    assert np.array_equal(out[0], expect), "getScores16"
    assert np.array_equal(out[1], expect), "getScores8 (same engine; equals the 8-bit kernel inside its envelope, SURVEY Q5)"
    assert in8.any() or case != "short8"
    assert np.array_equal(out[4], expect), "getScores16, 4 threads x 512-pair calls"
    assert np.array_equal(out[2], scalar), "scalarBandedSWAWrapper"
    k = min(n, 48)
    assert np.array_equal(out[3][:k], scalar[:k]), "scalarBandedSWA"


@pytest.mark.skipif(not BIN.exists(), reason="oracle/_ref/bsw_main_b200 not built (no reference tree at build time)")
@pytest.mark.parametrize("batch,threads", [(512, 4), (None, 2)])
def test_reference_driver_results(lib, oracle, tmp_path, batch, threads):
    """The UNMODIFIED main_banded.cpp on the engine, results checked pair by pair: the shim appends every call's
    result fields to BSW_SHIM_DUMP (the driver itself prints timings only); records are ordered by the address of
    the batch they belong to and compared with the oracle on the same file (driver defaults: w = 100, zdrop = 100,
    main_banded.cpp:250)."""
    import os
    from oracle.pyoracle import make_params
    cfg = lib.gen_named_config("small")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 5000)
    path = tmp_path / "pairs.txt"
    lib.write_pairs_file(str(path), pairs, ref, qer)
    dump = tmp_path / "dump.bin"
    cmd = [str(BIN), "-pairs", str(path), "-t", str(threads)] + (["-b", str(batch)] if batch else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, BSW_SHIM_DUMP=str(dump)))
    assert res.returncode == 1, res.stderr
    raw = dump.read_bytes()
    recs, pos = [], 0
    while pos < len(raw):
        addr, cnt = struct.unpack_from("<Qq", raw, pos)
        pos += 16
        recs.append((addr, np.frombuffer(raw, dtype=np.int32, count=cnt * 6, offset=pos).reshape(cnt, 6)))
        pos += cnt * 24
    recs.sort(key=lambda r: r[0])
    got = np.concatenate([r[1] for r in recs])
    assert len(got) == len(pairs)
    want = pairs.copy()
    oracle.batch(make_params(), want, ref, qer, 100)
    from conftest import results_matrix
    assert np.array_equal(got, results_matrix(want))
