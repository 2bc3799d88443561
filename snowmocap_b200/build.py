"""In-tree build of the CUDA library (sm_100a only).  Used by ``__graft_entry__.build()``."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libsnowtri.so")
OBJ_DIR = os.path.join(PKG, "build")
SOURCES = [os.path.join(CSRC, "snowtri_capi.cu"), os.path.join(CSRC, "snowtri_p1.cu"),
           os.path.join(CSRC, "snowtri_smooth.cu"), os.path.join(CSRC, "snowtri_general.cu"),
           os.path.join(CSRC, "snowtri_jit.cu"), os.path.join(CSRC, "snowtri_dlt.cu"),
           os.path.join(CSRC, "snowtri_blender.cu"), os.path.join(CSRC, "snowtri_comm.cu")]


def _headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(ROOT, "include", "snowtri.h")]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in SOURCES + _headers())


def _flags(verbose):
    fl = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
          "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-Xcompiler", "-fPIC"]
    # $CC/$CXX in this image point at a wrapper nvcc does not need; use the system g++
    if os.path.exists("/usr/bin/g++"):
        fl = ["-ccbin", "/usr/bin/g++"] + fl
    if verbose:
        fl.append("-Xptxas=-v")
    fl += os.environ.get("SNOWTRI_NVCC_FLAGS", "").split()   # experiments only (e.g. -DP1_MINB=3)
    return fl


def build(force=False, verbose=False, only=None):
    """Compile snowmocap_b200/libsnowtri.so for sm_100a (cross-compiles without a GPU).

    One object per translation unit, compiled in parallel, then one shared-library link."""
    if not force and only is None and up_to_date():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc, hdr_t = nvcc_path(), max(os.path.getmtime(p) for p in _headers())

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t)
        if only is not None:
            stale = os.path.basename(src) in only or not os.path.exists(obj)
        if stale:
            r = subprocess.run([nvcc] + _flags(verbose) + ["-c", src, "-o", obj], capture_output=True, text=True)
            if verbose or r.returncode:
                print(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.run([nvcc] + _flags(False) + ["-shared", "-o", LIB] + objs + ["-ldl"], check=True)
    return LIB
