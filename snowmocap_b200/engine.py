"""Batch API of the triangulation engine: torch tensors as device memory, ctypes into the C ABI.

``TriangulationEngine.run`` is the fused hot path (rays -> candidates -> clustering -> fuse,
i.e. reference ``main.py:55-71`` for every frame of a batch) and what the benchmark times.
PyTorch is used for allocation, streams and host<->device copies only.
"""
from __future__ import annotations

import ctypes as ct

import numpy as np
import torch

from . import _lib

PARAM_KEYS = ("kst", "ast", "dthr", "cond_tol", "num_tol", "score_tol", "center")
# defaults of the reference signatures (snowvision/triangulation.py:50, 95-100)
REFERENCE_DEFAULTS = dict(kst=0.5, ast=0.0, dthr=0.05, cond_tol=0.1, num_tol=0, score_tol=0.0, center=18)
# reference config keys (configs/snowmocap_default_config.json:10-16) -> engine parameter names
CONFIG_KEYS = {"keypoint_score_threshold": "kst", "average_score_threshold": "ast", "distance_threshold": "dthr",
               "condense_distance_tol": "cond_tol", "condense_person_num_tol": "num_tol",
               "condense_score_tol": "score_tol", "center_point_index": "center"}


def _ptr(t):
    return ct.c_void_p(t.data_ptr()) if t is not None else None


def _np_ptr(a):
    return ct.c_void_p(a.ctypes.data) if a is not None else None


def _stream():
    return ct.c_void_p(torch.cuda.current_stream().cuda_stream)


class TriangulationEngine:
    """One handle per device holding the camera parameters (K, R camera->world, t camera centre)."""

    def __init__(self, K, R, t, device=0, precision="f64", **params):
        self._h = ct.c_void_p()
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("snowmocap_b200 needs a CUDA device; there is no CPU fallback")
        K = np.ascontiguousarray(K, np.float64).reshape(-1, 3, 3)
        R = np.ascontiguousarray(R, np.float64).reshape(-1, 3, 3)
        t = np.ascontiguousarray(t, np.float64).reshape(-1, 3)
        if not (K.shape[0] == R.shape[0] == t.shape[0]):
            raise ValueError("K, R, t must describe the same number of cameras")
        self.C = int(K.shape[0])
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.snowtri_create(ct.byref(self._h), self.device.index, self.C,
                                                _np_ptr(K), _np_ptr(R), _np_ptr(t)))
        self.params = dict(REFERENCE_DEFAULTS)
        self.set_params(**params)
        self.set_precision(precision)

    # -- configuration ---------------------------------------------------------------------
    def set_params(self, **params):
        for k, v in params.items():
            k = CONFIG_KEYS.get(k, k)
            if k not in PARAM_KEYS:
                raise TypeError(f"unknown triangulation parameter {k!r}")
            self.params[k] = v
        p = self.params
        _lib.check(self._lib.snowtri_set_params(self._h, float(p["kst"]), float(p["ast"]), float(p["dthr"]),
                                                float(p["cond_tol"]), int(p["num_tol"]), float(p["score_tol"]),
                                                int(p["center"])), self._h)

    def set_precision(self, precision):
        code = {"f64": _lib.PREC_F64, "f32": _lib.PREC_F32, "mixed": _lib.PREC_MIXED,
                "f32x": _lib.PREC_F32_EXPERIMENTAL}[precision]
        _lib.check(self._lib.snowtri_set_precision(self._h, code), self._h)
        self.precision = precision

    def set_tuning(self, frames_per_group=0, max_ctas=0, threads=0):
        _lib.check(self._lib.snowtri_set_tuning(self._h, int(frames_per_group), int(max_ctas), int(threads)), self._h)

    def set_general_kernels(self, generation=2):
        """Several persons per camera, float modes: 2 = second-generation kernels (3 launches per chunk), 1 = the first."""
        _lib.check(self._lib.snowtri_set_general_kernels(self._h, int(generation)), self._h)

    def set_jit(self, mode="auto"):
        """Rig-specialised single-person kernel compiled at run time with NVRTC: "off", "auto" (long batches) or "always"."""
        _lib.check(self._lib.snowtri_set_jit(self._h, {"off": 0, "auto": 1, "always": 2}[mode]), self._h)

    @property
    def jit_status(self):
        return (self._lib.snowtri_jit_status(self._h) or b"").decode()

    def set_pipeline(self, frames_per_chunk=0):
        """Frames per chunk of run_host's copy/compute pipeline (0 = automatic)."""
        _lib.check(self._lib.snowtri_set_pipeline(self._h, int(frames_per_chunk)), self._h)

    @property
    def launch_count(self):
        return int(self._lib.snowtri_launch_count(self._h))

    def last_launch_info(self):
        g, b, s, fg = ct.c_int(), ct.c_int(), ct.c_int(), ct.c_int()
        self._lib.snowtri_last_launch_info(self._h, ct.byref(g), ct.byref(b), ct.byref(s), ct.byref(fg))
        return {"grid": g.value, "block": b.value, "smem_bytes": s.value, "frames_per_group": fg.value,
                "kernel": (self._lib.snowtri_last_kernel(self._h) or b"").decode()}

    def close(self):
        if self._h:
            self._lib.snowtri_destroy(self._h)
            self._h = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------
    def _check_inputs(self, kpts, scores, counts):
        if kpts.dim() != 5 or kpts.shape[-1] != 2 or scores.shape != kpts.shape[:-1]:
            raise ValueError("expected kpts (F,C,P,J,2) and scores (F,C,P,J)")
        F, C, P, J = scores.shape
        if C != self.C:
            raise ValueError(f"batch has {C} cameras, engine was built for {self.C}")
        for name, x, dt in (("kpts", kpts, torch.float32), ("scores", scores, torch.float32)):
            if x.dtype != dt or not x.is_cuda or not x.is_contiguous() or x.device != self.device:
                raise ValueError(f"{name} must be a contiguous {dt} tensor on {self.device}")
        if counts is not None:
            if counts.shape != (F, C) or counts.dtype != torch.int32 or not counts.is_cuda or not counts.is_contiguous():
                raise ValueError("counts must be a contiguous int32 cuda tensor of shape (F,C)")
        return F, C, P, J

    # -- fused hot path --------------------------------------------------------------------
    def run(self, kpts, scores, counts=None, Pout=None, keypoint_num=None, out=None):
        """Fused path on device tensors.  Returns dict(out (F,Pout,Jout,4) f32 [x,y,z,score],
        pscores (F,Pout) f32, nout (F,) i32).  Asynchronous on the current stream."""
        F, C, P, J = self._check_inputs(kpts, scores, counts)
        Jout = J if keypoint_num is None else int(keypoint_num)
        Pout = P if Pout is None else int(Pout)
        if out is None:
            out = {"out": torch.empty((F, Pout, Jout, 4), dtype=torch.float32, device=self.device),
                   "pscores": torch.empty((F, Pout), dtype=torch.float32, device=self.device),
                   "nout": torch.empty((F,), dtype=torch.int32, device=self.device)}
        with torch.cuda.device(self.device):
            _lib.check(self._lib.snowtri_run(self._h, _ptr(kpts), _ptr(scores), _ptr(counts), F, P, J, Jout, Pout,
                                             _ptr(out["out"]), _ptr(out["pscores"]), _ptr(out["nout"]), _stream()),
                       self._h)
        return out

    # -- optional DLT mode (not the reference's estimator) ---------------------------------------
    def dlt(self, kpts, scores, accumulate_f64=False):
        """Homogeneous linear triangulation of every (frame, joint) from all cameras with score >= kst; one person
        per camera: kpts (F,C,1,J,2), scores (F,C,1,J).  Returns (F,J,4) float32: x, y, z, views used."""
        F, C, P, J = self._check_inputs(kpts, scores, None)
        if P != 1:
            raise ValueError("the DLT mode takes one (already matched) person per camera")
        out = torch.empty((F, J, 4), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.snowtri_dlt_run(self._h, _ptr(kpts), _ptr(scores), F, J, _ptr(out),
                                                 1 if accumulate_f64 else 0, _stream()), self._h)
        return out

    # -- ragged ingestion -----------------------------------------------------------------
    def pack_detections(self, detections, P=None):
        """Detector outputs of a clip -> dense device batch.  ``detections[f][c]`` is the pair
        ``(keypoints (N,J,2), scores (N,J))`` the reference's ``main.py:53`` gets for frame f, camera c
        (N may be 0).  One concatenate per array on the host, one copy, one kernel; no per-keypoint loop.
        Returns (kpts (F,C,P,J,2), scores (F,C,P,J), counts (F,C)) cuda tensors; P defaults to the largest N."""
        F, C = len(detections), self.C
        ks, ss, ns = [], [], []
        for frame in detections:
            if len(frame) != C:
                raise ValueError(f"every frame needs {C} cameras")
            for k, s in frame:
                k = np.asarray(k, np.float32)
                s = np.asarray(s, np.float32)
                ns.append(k.shape[0])
                if k.shape[0]:
                    ks.append(k.reshape(k.shape[0], -1, 2))
                    ss.append(s.reshape(s.shape[0], -1))
        if not ks:
            raise ValueError("no detections in the clip")
        det_k, det_s = np.concatenate(ks), np.concatenate(ss)
        J = det_k.shape[1]
        if det_s.shape != det_k.shape[:2]:
            raise ValueError("scores and keypoints disagree")
        offsets = np.zeros(F * C + 1, np.int64)
        np.cumsum(ns, out=offsets[1:])
        P = int(max(ns)) if P is None else int(P)
        dev = self.device
        dk, ds = torch.from_numpy(det_k).to(dev), torch.from_numpy(det_s).to(dev)
        do = torch.from_numpy(offsets).to(dev)
        kpts = torch.empty((F, C, P, J, 2), dtype=torch.float32, device=dev)
        scores = torch.empty((F, C, P, J), dtype=torch.float32, device=dev)
        counts = torch.empty((F, C), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self._lib.snowtri_pack_ragged(self._h, _ptr(dk), _ptr(ds), _ptr(do), F, P, J, _ptr(kpts),
                                                     _ptr(scores), _ptr(counts), _stream()), self._h)
        return kpts, scores, counts

    def run_host(self, kpts, scores, counts=None, Pout=None, keypoint_num=None, out=None):
        """Fused path through HOST numpy buffers (H2D + kernel + D2H inside the C call, synchronous)."""
        kpts = np.ascontiguousarray(kpts, np.float32)
        scores = np.ascontiguousarray(scores, np.float32)
        F, C, P, J = scores.shape
        if C != self.C or kpts.shape != (F, C, P, J, 2):
            raise ValueError("expected kpts (F,C,P,J,2) and scores (F,C,P,J) for this engine's cameras")
        if counts is not None:
            counts = np.ascontiguousarray(counts, np.int32).reshape(F, C)
        Jout = J if keypoint_num is None else int(keypoint_num)
        Pout = P if Pout is None else int(Pout)
        if out is None:
            out = {"out": np.empty((F, Pout, Jout, 4), np.float32), "pscores": np.empty((F, Pout), np.float32),
                   "nout": np.empty((F,), np.int32)}
        with torch.cuda.device(self.device):
            _lib.check(self._lib.snowtri_run_host(self._h, _np_ptr(kpts), _np_ptr(scores), _np_ptr(counts), F, P, J,
                                                  Jout, Pout, _np_ptr(out["out"]), _np_ptr(out["pscores"]),
                                                  _np_ptr(out["nout"]), _stream()), self._h)
        return out

    # -- the two reference functions separately (float64) ----------------------------------
    def candidates(self, kpts, scores, counts=None):
        """Human_Triangulation for every frame: dict(cand (F,Nc,J,4) f64, avg (F,Nc) f64, keep (F,Nc) i32),
        dense candidate index n = ((pair*P + pm)*P + ps)."""
        F, C, P, J = self._check_inputs(kpts, scores, counts)
        nc = C * (C - 1) // 2 * P * P
        res = {"cand": torch.zeros((F, nc, J, 4), dtype=torch.float64, device=self.device),
               "avg": torch.zeros((F, nc), dtype=torch.float64, device=self.device),
               "keep": torch.zeros((F, nc), dtype=torch.int32, device=self.device)}
        if nc:
            with torch.cuda.device(self.device):
                _lib.check(self._lib.snowtri_candidates(self._h, _ptr(kpts), _ptr(scores), _ptr(counts), F, P, J,
                                                        _ptr(res["cand"]), _ptr(res["avg"]), _ptr(res["keep"]),
                                                        _stream()), self._h)
        return res

    def condense(self, cand, ncand, keypoint_num=None, Pout=None):
        """Human_Triangulation_Condense on materialised candidates cand (F,N,J,4) f64, ncand (F,) i32."""
        if cand.dim() != 4 or cand.shape[-1] != 4 or cand.dtype != torch.float64 or not cand.is_contiguous():
            raise ValueError("cand must be a contiguous float64 tensor of shape (F,N,J,4)")
        F, N, J, _ = cand.shape
        Jout = J if keypoint_num is None else int(keypoint_num)
        Pout = max(1, N) if Pout is None else int(Pout)
        res = {"out": torch.empty((F, Pout, Jout, 4), dtype=torch.float64, device=self.device),
               "pscores": torch.empty((F, Pout), dtype=torch.float64, device=self.device),
               "nout": torch.empty((F,), dtype=torch.int32, device=self.device)}
        with torch.cuda.device(self.device):
            _lib.check(self._lib.snowtri_condense(self._h, _ptr(cand), _ptr(ncand), F, N, J, Jout, Pout,
                                                  _ptr(res["out"]), _ptr(res["pscores"]), _ptr(res["nout"]),
                                                  _stream()), self._h)
        return res

    def skew_ray(self, hm, hs, tm, ts):
        """Batched Skew_Ray_Solver on (n,3) float64 cuda tensors -> (dist (n,), mid (n,3))."""
        n = hm.shape[0]
        dist = torch.empty((n,), dtype=torch.float64, device=self.device)
        mid = torch.empty((n, 3), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.snowtri_skew_ray(self._h, n, _ptr(hm), _ptr(hs), _ptr(tm), _ptr(ts), _ptr(dist),
                                                  _ptr(mid), _stream()), self._h)
        return dist, mid


class SmoothState:
    """Device-resident followers of ``Human_Triangulation_Smooth`` (reference triangulation.py:4-22, 164-186)
    for one clip: create once, then ``run`` batches of consecutive frames in order.

    ``max_persons`` must cover the person count of the clip's FIRST frame (``max_persons >= Pout`` is always enough):
    the reference keeps one follower per first-frame person; a first frame with more persons than ``max_persons`` gets
    followers for the first ``max_persons`` only and the others stay unsmoothed on later frames, silently."""

    def __init__(self, engine, max_persons, J, f=2.0, z=0.75, r=0.0):
        self._eng, self._lib, self._s = engine, engine._lib, ct.c_void_p()
        self.max_persons, self.J = int(max_persons), int(J)
        with torch.cuda.device(engine.device):
            _lib.check(self._lib.snowtri_smooth_create(engine._h, ct.byref(self._s), self.max_persons, self.J,
                                                       float(f), float(z), float(r)), engine._h)

    def set_chunked(self, enabled=True):
        """Long batches run as parallel chunks (default: one pass with a warm-up per chunk when the filter forgets its
        state within 256 frames, else the chunk scan); ``"scan"`` forces the chunk scan, False the sequential kernel."""
        mode = 2 if enabled == "scan" else (1 if enabled else 0)
        _lib.check(self._lib.snowtri_smooth_set_chunked(self._s, mode), self._eng._h)

    def reset(self):
        """The next frame starts a new clip (passes through and seeds the followers)."""
        _lib.check(self._lib.snowtri_smooth_reset(self._eng._h, self._s, _stream()), self._eng._h)

    def run(self, out, nout, delta_time=1 / 30):
        """Smooth ``out`` (F,Pout,J,4) float32/float64 cuda tensor in place (x, y, z; score untouched), frames
        in order.  ``nout`` (F,) int32 persons per frame.  Returns nsmooth (F,) int32: persons in the smoothed
        list of every frame (first frame of the clip: all of them; later: min(nout, first frame's count))."""
        if out.dim() != 4 or out.shape[-1] != 4 or not out.is_cuda or not out.is_contiguous():
            raise ValueError("out must be a contiguous cuda tensor of shape (F,Pout,J,4)")
        if out.dtype not in (torch.float32, torch.float64):
            raise ValueError("out must be float32 or float64")
        F, Pout, J, _ = out.shape
        if nout.shape != (F,) or nout.dtype != torch.int32 or not nout.is_cuda:
            raise ValueError("nout must be an int32 cuda tensor of shape (F,)")
        nsm = torch.empty((F,), dtype=torch.int32, device=out.device)
        fn = self._lib.snowtri_smooth_run if out.dtype == torch.float32 else self._lib.snowtri_smooth_run_f64
        with torch.cuda.device(self._eng.device):
            _lib.check(fn(self._eng._h, self._s, _ptr(out), _ptr(nout), _ptr(nsm), F, Pout, J, float(delta_time),
                          _stream()), self._eng._h)
        return nsm

    def close(self):
        if self._s:
            self._lib.snowtri_smooth_destroy(self._s)
            self._s = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def triangulate_batch(kpts, scores, counts, K, R, t, Pout=None, keypoint_num=None, precision="f64", **params):
    """One-shot convenience wrapper: build an engine, run the fused path, return its result dict."""
    eng = TriangulationEngine(K, R, t, device=kpts.device.index or 0, precision=precision, **params)
    try:
        res = eng.run(kpts, scores, counts, Pout=Pout, keypoint_num=keypoint_num)
        torch.cuda.synchronize(kpts.device)
    finally:
        eng.close()
    return res
