"""Builds libbsw_b200.so (CUDA kernels + C ABI + host pipeline) in-tree for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the
resulting .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libbsw_b200.so"
STAMP = LIB_DIR / ".build_stamp"

SOURCES = [
    "bsw_gen.cpp",
    "bsw_host.cpp",
    "bsw_pack.cpp",
    "bsw_engine.cu",
    "bsw_shim.cpp",
]
# libbsw_host.so: the host-only part of the ABI (generator, text format, packed-batch builders, bucketing /
# partition utilities) without any CUDA code, for tools that must not map the CUDA library -- bench.py's
# reference arm generates its inputs through it
HOST_LIB_PATH = LIB_DIR / "libbsw_host.so"
HOST_SOURCES = ["bsw_gen.cpp", "bsw_host.cpp", "bsw_pack.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function,-pthread",
    "-Xptxas", "-v",
    "-cudart", "static",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the bsw_b200 library cannot be built")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + [PKG_DIR.parent / "include" / "bsw.h",
                                              PKG_DIR.parent / "include" / "bandedSWA.h"]):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if sources changed; returns the path of the .so."""
    LIB_DIR.mkdir(exist_ok=True)
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and HOST_LIB_PATH.exists() and STAMP.exists() and STAMP.read_text() == fp:
        return LIB_PATH
    # never try to rebuild on a box without nvcc sources context (e.g. GPU box has nvcc too,
    # but a prebuilt library with a stale stamp is still better than none)
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(PKG_DIR.parent / "include"), "-o", str(LIB_PATH)]
    cmd += [str(CSRC / s) for s in SOURCES]
    cmd += ["-lpthread", "-ldl"]
    env = dict(os.environ)
    env["PATH"] = "/usr/bin:" + env.get("PATH", "")      # host compiler with a complete toolchain
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    log = res.stdout + res.stderr
    (LIB_DIR / "build.log").write_text(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libbsw_b200.so (see genomicsbench_b200/lib/build.log)")
    if verbose:
        print(log)
    hcmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-Wno-unused-function",
            "-I", str(PKG_DIR.parent / "include"), "-o", str(HOST_LIB_PATH)] + [str(CSRC / s) for s in HOST_SOURCES]
    res = subprocess.run(hcmd, capture_output=True, text=True, env=env)
    with open(LIB_DIR / "build.log", "a") as f:
        f.write(" ".join(hcmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building libbsw_host.so (see genomicsbench_b200/lib/build.log)")
    STAMP.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("built", p)
