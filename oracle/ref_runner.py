"""TEST / BASELINE INFRASTRUCTURE ONLY -- the REAL reference (``snowvision``) as a checker and as the CPU arm.

The reference is pure Python (``/root/reference/snowvision``: camera.py, triangulation.py, blender.py, util.py).
``/root/reference`` exists in the build container only, so ``make_ref()`` -- called by ``__graft_entry__.build()`` --
vendors the package byte for byte into the git-ignored ``oracle/_ref/snowvision`` (listed in ``.gitignore``, not in
``.gpurunignore``: it travels to the GPU box with the snapshot, it never enters the history).  Nothing here is
imported by the product package; only ``tests/``, ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and the
golden-vector scripts use it.

``run_frames`` is the reference's own ``main.py:55-71`` body per frame:
    add_human_2D_points (camera.py:234-253) for every (camera, person)
    Human_Triangulation (triangulation.py:50-93)
    Human_Triangulation_Condense (triangulation.py:95-162)
    clear_2D_points (camera.py:255-261)
"""
from __future__ import annotations

import os
import shutil
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
SRC = os.environ.get("SNOW_REFERENCE", "/root/reference")
_MOD = None


def make_ref():
    """Copy the reference package to oracle/_ref/snowvision when the reference tree is present.  Returns the path,
    or None when neither the source nor an earlier copy exists."""
    src, dst = os.path.join(SRC, "snowvision"), os.path.join(REF_DIR, "snowvision")
    if os.path.isdir(src):
        os.makedirs(REF_DIR, exist_ok=True)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__"))
        for root, _, files in os.walk(dst):      # the source mount is read-only; make the copy removable
            os.chmod(root, 0o755)
            for f in files:
                os.chmod(os.path.join(root, f), 0o644)
    return dst if os.path.isdir(dst) else None


def available():
    return os.path.isdir(os.path.join(REF_DIR, "snowvision"))


def load():
    """Import the vendored reference (needs numpy, cv2 and scipy, which snowvision/__init__.py star-imports)."""
    global _MOD
    if _MOD is None:
        if not available():
            raise ImportError("oracle/_ref/snowvision is missing: run __graft_entry__.build() where /root/reference exists")
        sys.dont_write_bytecode = True
        sys.path.insert(0, REF_DIR)
        try:
            import snowvision as mod
        finally:
            sys.path.remove(REF_DIR)
        _MOD = mod
    return _MOD


def group_of(K, R, t):
    """A reference CameraGroup carrying the given calibration (what CameraGroup(camera_group_info_path=...) loads)."""
    ref = load()
    C = len(K)
    group = ref.CameraGroup(cap_ids=list(range(C)), resolutions=[(1280, 720)] * C)
    for c in range(C):
        group.cameras[c].K = np.array(K[c], np.float64)
        group.cameras[c].R = np.array(R[c], np.float64)
        group.cameras[c].t = np.array(t[c], np.float64).reshape(3, 1)
    return group


def run_frames(K, R, t, kpts, scores, counts, prm, keypoint_num=None, keep_tri=False):
    """main.py:55-71 for every frame.  Returns (list of (tri or None, condensed) dicts, seconds)."""
    ref = load()
    group = group_of(K, R, t)
    F, C, P, J = scores.shape
    kp64, sc64 = kpts.astype(np.float64), scores.astype(np.float64)
    out = []
    t0 = time.perf_counter()
    for f in range(F):
        for c in range(C):
            n = P if counts is None else int(counts[f, c])
            for p in range(n):
                group.add_human_2D_points(kp64[f, c, p], sc64[f, c, p], c)
        with np.errstate(all="ignore"):
            tri = ref.Human_Triangulation(group, keypoint_score_threshold=prm["kst"],
                                          average_score_threshold=prm["ast"], distance_threshold=prm["dthr"])
            con = ref.Human_Triangulation_Condense(
                tri, condense_distance_tol=prm["cond_tol"], condense_person_num_tol=prm["num_tol"],
                condense_score_tol=prm["score_tol"], center_point_index=prm["center"],
                keypoint_num=J if keypoint_num is None else keypoint_num)
        group.clear_2D_points()
        out.append((tri if keep_tri else None, con))
    return out, time.perf_counter() - t0


def dense(results, Pout, J):
    """Condensed dicts -> dense arrays like the fused path's: points (F,Pout,J,3), kscores, pscores, nout."""
    F = len(results)
    pts = np.zeros((F, Pout, J, 3))
    ks = np.zeros((F, Pout, J))
    ps = np.zeros((F, Pout))
    nout = np.zeros(F, np.int32)
    for f, (_, con) in enumerate(results):
        n = len(con["hrnet_triangulate_points"])
        nout[f] = n
        for k in range(min(n, Pout)):
            pts[f, k] = con["hrnet_triangulate_points"][k]
            ks[f, k] = con["hrnet_triangulate_keypoint_scores"][k]
            ps[f, k] = con["hrnet_triangulate_person_scores"][k]
    return {"points": pts, "kscores": ks, "pscores": ps, "nout": nout}
