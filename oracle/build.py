"""Build the C++ restatement of the oracle (oracle/cpu_frame.cpp -> oracle/_build/libfers_cpu.so) with g++ -fopenmp.
TEST INFRASTRUCTURE: called by __graft_entry__.build() and, lazily, by oracle.cpu_frame.load()."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_frame.cpp")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libfers_cpu.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found: cannot build the CPU oracle library")
    os.makedirs(OUT_DIR, exist_ok=True)
    # AVX2 + FMA only (the build container and the GPU box's host differ; the dense kernels come from OpenBLAS anyway)
    subprocess.check_call([gxx, "-O3", "-std=c++17", "-mavx2", "-mfma", "-fopenmp", "-fPIC", "-shared", SRC, "-o", LIB, "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
