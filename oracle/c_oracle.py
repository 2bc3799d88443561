"""ctypes front-end of the C oracle (``snow_oracle.c``) -- TEST INFRASTRUCTURE ONLY.

Same restatement as ``loop_oracle.py`` (which cites the reference lines), compiled for
speed and threaded over frames.  Only tests, ``smoke()`` and bench.py's CPU-baseline legs
may import this module.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build():
    """Compile libsnow_oracle.so next to its source (idempotent; ``make`` checks timestamps)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return os.path.join(_HERE, "libsnow_oracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libsnow_oracle.so")
        try:
            build()            # make: rebuilds only when snow_oracle.c is newer than the library
        except Exception:
            if not os.path.exists(path):
                raise
        L = ct.CDLL(path)
        L.snow_oracle_fused.restype = ct.c_int
        L.snow_oracle_fused.argtypes = [ct.c_int] * 4 + [_f32, _f32, ct.c_void_p, _f64, _f64, _f64,
                                                          ct.c_double, ct.c_double, ct.c_double, ct.c_double,
                                                          ct.c_int, ct.c_double, ct.c_int, ct.c_int, ct.c_int,
                                                          _f64, _f64, _f64, _i32, _i32, ct.c_int]
        L.snow_oracle_candidates.restype = ct.c_int
        L.snow_oracle_candidates.argtypes = [ct.c_int] * 3 + [_f32, _f32, ct.c_void_p, _f64, _f64, _f64,
                                                               ct.c_double, ct.c_double, ct.c_double, ct.c_int,
                                                               _f64, _f64, _f64, _i32]
        L.snow_oracle_condense.restype = ct.c_int
        L.snow_oracle_condense.argtypes = [ct.c_int, ct.c_int, _f64, _f64, ct.c_double, ct.c_int, ct.c_double,
                                           ct.c_int, ct.c_int, ct.c_int, _f64, _f64, _f64]
        L.snow_oracle_skew_ray.restype = None
        L.snow_oracle_skew_ray.argtypes = [ct.c_int, _f64, _f64, _f64, _f64, _f64, _f64]
        L.snow_oracle_smooth.restype = ct.c_int
        L.snow_oracle_smooth.argtypes = [ct.c_int, ct.c_int, ct.c_int, _f64, _i32, _i32, ct.c_double, ct.c_double,
                                         ct.c_double, ct.c_double, _f64]
        L.snow_oracle_max_threads.restype = ct.c_int
        _LIB = L
    return _LIB


def max_threads():
    return int(lib().snow_oracle_max_threads())


def _counts_ptr(counts):
    if counts is None:
        return None, None
    c = np.ascontiguousarray(counts, np.int32)
    return c, c.ctypes.data_as(ct.c_void_p)


def fused(kpts, scores, counts, K, R, t, params, Pout, keypoint_num=None, nthreads=0):
    """Frames (F,C,P,J,2)/(F,C,P,J) -> dict(points (F,Pout,Jout,3), kscores, pscores, nout, ncand)."""
    kpts = np.ascontiguousarray(kpts, np.float32)
    scores = np.ascontiguousarray(scores, np.float32)
    F, C, P, J = scores.shape
    Jout = J if keypoint_num is None else int(keypoint_num)
    keep, cptr = _counts_ptr(counts)
    pts = np.zeros((F, Pout, Jout, 3))
    ks = np.zeros((F, Pout, Jout))
    ps = np.zeros((F, Pout))
    nout = np.zeros(F, np.int32)
    ncand = np.zeros(F, np.int32)
    rc = lib().snow_oracle_fused(F, C, P, J, kpts, scores, cptr,
                                 np.ascontiguousarray(K, np.float64), np.ascontiguousarray(R, np.float64),
                                 np.ascontiguousarray(t, np.float64).reshape(C, 3),
                                 params["kst"], params["ast"], params["dthr"], params["cond_tol"],
                                 int(params["num_tol"]), params["score_tol"], int(params["center"]),
                                 Jout, Pout, pts, ks, ps, nout, ncand, nthreads)
    if rc != 0:
        raise RuntimeError("snow_oracle_fused failed (bad arguments or out of memory)")
    del keep
    return {"points": pts, "kscores": ks, "pscores": ps, "nout": nout, "ncand": ncand}


def candidates(kpts, scores, counts, K, R, t, kst, ast, dthr):
    """One frame (C,P,J,2)/(C,P,J) -> dict(points (N,J,3), kscores (N,J), pscores (N), index (N,4))."""
    kpts = np.ascontiguousarray(kpts, np.float32)
    scores = np.ascontiguousarray(scores, np.float32)
    C, P, J = scores.shape
    nmax = max(1, C * (C - 1) // 2 * P * P)
    keep, cptr = _counts_ptr(counts)
    pts = np.zeros((nmax, J, 3))
    ks = np.zeros((nmax, J))
    ps = np.zeros(nmax)
    idx = np.zeros((nmax, 4), np.int32)
    n = lib().snow_oracle_candidates(C, P, J, kpts, scores, cptr,
                                     np.ascontiguousarray(K, np.float64), np.ascontiguousarray(R, np.float64),
                                     np.ascontiguousarray(t, np.float64).reshape(C, 3),
                                     kst, ast, dthr, nmax, pts, ks, ps, idx)
    if n < 0:
        raise RuntimeError("snow_oracle_candidates failed")
    del keep
    return {"points": pts[:n], "kscores": ks[:n], "pscores": ps[:n], "index": idx[:n]}


def condense(points, kscores, cond_tol, num_tol, score_tol, center, keypoint_num, Pout=None):
    """Candidates (N,J,3)/(N,J) -> dict(points (M,Jout,3), kscores (M,Jout), pscores (M))."""
    points = np.ascontiguousarray(points, np.float64)
    kscores = np.ascontiguousarray(kscores, np.float64)
    N = points.shape[0]
    J = points.shape[1] if N else int(keypoint_num)
    Pout = max(1, N) if Pout is None else Pout
    o_p = np.zeros((Pout, keypoint_num, 3))
    o_k = np.zeros((Pout, keypoint_num))
    o_s = np.zeros(Pout)
    if N == 0:
        points = np.zeros((1, J, 3))
        kscores = np.zeros((1, J))
    m = lib().snow_oracle_condense(N, J, points, kscores, cond_tol, int(num_tol), score_tol, int(center),
                                   int(keypoint_num), Pout, o_p, o_k, o_s)
    if m < 0:
        raise RuntimeError("snow_oracle_condense failed (center / keypoint_num out of range)")
    m = min(m, Pout)
    return {"points": o_p[:m], "kscores": o_k[:m], "pscores": o_s[:m]}


def skew_ray(hm, hs, tm, ts):
    hm, hs, tm, ts = (np.ascontiguousarray(a, np.float64).reshape(-1, 3) for a in (hm, hs, tm, ts))
    n = hm.shape[0]
    dist = np.zeros(n)
    W = np.zeros((n, 3))
    lib().snow_oracle_skew_ray(n, hm, hs, tm, ts, dist, W)
    return dist, W


def smooth(points, nout, f=2.0, z=0.75, r=0.0, delta_time=1 / 30, state=None):
    """Human_Triangulation_Smooth over consecutive frames, dense layout: points (F,P,J,3) f64, nout (F,).
    Returns (smoothed points, nsm (F,), state) -- pass `state` back in to continue the same clip."""
    pts = np.array(points, np.float64, order="C", copy=True)
    F, P, J, _ = pts.shape
    nout = np.ascontiguousarray(nout, np.int32)
    nsm = np.zeros(F, np.int32)
    if state is None:
        state = np.zeros(2 + 3 * P * J * 3, np.float64)
    lib().snow_oracle_smooth(F, P, J, pts, nout, nsm, float(f), float(z), float(r), float(delta_time), state)
    return pts, nsm, state
