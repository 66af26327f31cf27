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
