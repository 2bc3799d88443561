"""In-tree build of the CUDA library (sm_100a only).  Used by ``__graft_entry__.build()``."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libsnowtri.so")
SOURCES = [os.path.join(PKG, "csrc", "snowtri_capi.cu")]
HEADERS = [os.path.join(PKG, "csrc", "snowtri_kernels.cuh"), os.path.join(PKG, "csrc", "snowtri_math.cuh"),
           os.path.join(ROOT, "include", "snowtri.h")]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile snowmocap_b200/libsnowtri.so for sm_100a (cross-compiles without a GPU)."""
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"),
           "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    # $CC/$CXX in this image point at a wrapper nvcc does not need; use the system g++
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    subprocess.run(cmd, check=True, env=env)
    return LIB
