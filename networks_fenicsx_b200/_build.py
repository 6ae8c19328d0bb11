"""Build libnxfx_b200.so in-tree with nvcc for sm_100a (called by ``__graft_entry__.build``)."""

from __future__ import annotations

import os
import pathlib
import shutil
import subprocess

CSRC = pathlib.Path(__file__).parent / "csrc"
LIB = CSRC / "libnxfx_b200.so"
SOURCES = ["nxfx_b200.cu"]
HEADERS = ["ctx.cuh", "assemble.cuh", "spmv.cuh", "precond.cuh", "generic.cuh", "condense.cuh", "peer.cuh", "../../include/nxfx_b200.h"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libnxfx_b200.so cannot be built")


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any((CSRC / f).resolve().stat().st_mtime > t for f in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> pathlib.Path:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> csrc/libnxfx_b200.so"""
    if not force and not is_stale():
        return LIB
    cmd = [
        nvcc_path(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared",
        "-o", str(LIB),
    ] + [str(CSRC / s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return LIB
