"""Builds tests/emu/libk16_emu.so: the packed kernel's row sweep compiled as HOST code (nvcc)."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
SO = HERE / "libk16_emu.so"
SRC = HERE / "k16_emu.cu"
DEPS = [SRC, ROOT / "genomicsbench_b200/csrc/bsw_kernel16.cuh", ROOT / "genomicsbench_b200/csrc/bsw_kernels.cuh"]


def build(force: bool = False) -> Path:
    if not force and SO.exists() and all(SO.stat().st_mtime >= d.stat().st_mtime for d in DEPS):
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
           "-Xcompiler", "-fPIC,-fopenmp", "-I", str(ROOT / "include"), "-o", str(SO), str(SRC), "-lgomp"]
    env = dict(os.environ)
    env["PATH"] = "/usr/bin:" + env.get("PATH", "")
    subprocess.run(cmd, check=True, env=env)
    return SO


if __name__ == "__main__":
    print(build(force=True))
