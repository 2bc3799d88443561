"""Drop-in mirror of the hot-path functions of ``snowvision.triangulation``.

Same names, signatures, defaults, dict keys and list-of-ndarray results as the reference
(triangulation.py:24-31, 50-93, 95-162); the arithmetic runs in hand-written sm_100a kernels
through the C ABI of ``include/snowtri.h``.  There is no CPU fallback.

Errors: the reference raises ``IndexError`` when ``keypoint_num`` exceeds the number of joints
or ``center_point_index`` is out of range; so do these functions.  Exactly parallel rays make
the reference raise ``LinAlgError``; here they produce inf/NaN like near-parallel rays do.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import SmoothState, TriangulationEngine

POINTS = "hrnet_triangulate_points"
KSCORES = "hrnet_triangulate_keypoint_scores"
PSCORES = "hrnet_triangulate_person_scores"

_util_engines = {}


def _util_engine(device=0):
    """Camera-less handle for the entry points that do not need camera parameters."""
    if device not in _util_engines:
        _util_engines[device] = TriangulationEngine(np.eye(3)[None], np.eye(3)[None], np.zeros((1, 3)), device=device)
    return _util_engines[device]


def Skew_Ray_Solver(hm, hs, tm, ts):
    """Closest points of two rays -> (distance, midpoint (3,)); reference triangulation.py:24-31."""
    eng = _util_engine()
    dev = eng.device
    args = [torch.as_tensor(np.asarray(a, np.float64).reshape(1, 3), device=dev) for a in (hm, hs, tm, ts)]
    dist, mid = eng.skew_ray(*args)
    return float(dist.cpu()[0]), mid.cpu().numpy().reshape(3)


def Human_Triangulation(camera_group, keypoint_score_threshold=0.5, average_score_threshold=0.0,
                        distance_threshold=0.05):
    """All (camera pair x person pair) candidates of the current frame; reference triangulation.py:50-93."""
    kpts, scores, counts = camera_group.pack_frame()
    result = {POINTS: [], KSCORES: [], PSCORES: []}
    if kpts.shape[2] == 0 or kpts.shape[1] < 2:
        return result
    eng = camera_group.engine()
    saved = dict(eng.params)
    eng.set_params(kst=keypoint_score_threshold, ast=average_score_threshold, dthr=distance_threshold)
    try:
        dev = eng.device
        res = eng.candidates(torch.from_numpy(kpts).to(dev), torch.from_numpy(scores).to(dev),
                             torch.from_numpy(counts).to(dev))
        keep = res["keep"][0].cpu().numpy().astype(bool)
        cand = res["cand"][0].cpu().numpy()[keep]
        avg = res["avg"][0].cpu().numpy()[keep]
    finally:
        eng.set_params(**saved)
    for i in range(cand.shape[0]):
        result[POINTS].append(np.ascontiguousarray(cand[i, :, :3]))
        result[KSCORES].append(np.ascontiguousarray(cand[i, :, 3]))
        result[PSCORES].append(np.float64(avg[i]))
    return result


def Human_Triangulation_Condense(result, condense_distance_tol=0.1, condense_person_num_tol=0,
                                 condense_score_tol=0.0, center_point_index=18, keypoint_num=30):
    """Greedy centre-joint clustering + score-weighted fuse; reference triangulation.py:95-162."""
    pts, ks = result[POINTS], result[KSCORES]
    out = {POINTS: [], KSCORES: [], PSCORES: []}
    N = len(pts)
    if N <= 1:          # range(person_num - 1) is empty: nothing is ever emitted
        return out
    J = int(np.asarray(pts[0]).shape[0])
    if keypoint_num > J or not (-J <= center_point_index < J):
        raise IndexError(f"index out of bounds for {J} keypoints (keypoint_num={keypoint_num}, "
                         f"center_point_index={center_point_index})")
    center = center_point_index % J
    if keypoint_num < 1:
        # the reference emits zero-length persons with a NaN mean; nothing useful to compute
        keypoint_num = 0
    cand = np.empty((1, N, J, 4), np.float64)
    cand[0, :, :, :3] = np.asarray(pts, np.float64).reshape(N, J, 3)
    cand[0, :, :, 3] = np.asarray(ks, np.float64).reshape(N, J)
    eng = _util_engine()
    saved = dict(eng.params)
    eng.set_params(cond_tol=condense_distance_tol, num_tol=condense_person_num_tol,
                   score_tol=condense_score_tol, center=center)
    try:
        if keypoint_num == 0:
            raise _lib.SnowtriError(_lib.E_ARG, "keypoint_num must be >= 1")
        dev = eng.device
        res = eng.condense(torch.from_numpy(cand).to(dev), torch.tensor([N], dtype=torch.int32, device=dev),
                           keypoint_num=keypoint_num, Pout=N)
        n = int(res["nout"].cpu()[0])
        o = res["out"][0, :n].cpu().numpy()
        ps = res["pscores"][0, :n].cpu().numpy()
    finally:
        eng.set_params(**saved)
    for i in range(n):
        out[POINTS].append(np.ascontiguousarray(o[i, :, :3]))
        out[KSCORES].append(np.ascontiguousarray(o[i, :, 3]))
        out[PSCORES].append(np.float64(ps[i]))
    return out


SODS = "second_order_dynamics"


class _Followers:
    """What this package stores under ``'second_order_dynamics'``: the device-resident followers of one clip
    (the reference stores a list of lists of ``SecondOrderDynamic`` objects there, triangulation.py:176,184)."""

    def __init__(self, n0, J, f, z, r):
        self.n0, self.J = n0, J
        self.state = SmoothState(_util_engine(), max(n0, 1), J, f, z, r) if n0 > 0 and J > 0 else None

    def __len__(self):          # len(previous_result['second_order_dynamics']) == persons of the first frame
        return self.n0


def Human_Triangulation_Smooth(result, previous_result=None, f=2, z=0.75, r=0, delta_time=1 / 30):
    """Second-order-dynamics smoothing of the 3D joints; reference triangulation.py:164-186 (main.py:72-78).

    First call of a clip (``previous_result`` not a dict): points pass through, followers are created from them.
    Later calls: persons are zipped by list position with the followers of the FIRST frame (later persons are
    dropped, absent persons' followers do not advance); scores pass through unaligned, like the reference."""
    pts = result[POINTS]
    if not isinstance(previous_result, dict):
        n0 = len(pts)
        J = int(np.asarray(pts[0]).shape[0]) if n0 else 0
        fol = _Followers(n0, J, f, z, r)
        if fol.state is not None:
            _smooth_step(fol, pts, delta_time)      # seeds the followers; the frame itself is unchanged
        return {POINTS: result[POINTS], KSCORES: result[KSCORES], PSCORES: result[PSCORES], SODS: fol}
    fol = previous_result[SODS]
    damped = _smooth_step(fol, pts, delta_time) if fol.state is not None else []
    return {POINTS: damped, KSCORES: result[KSCORES], PSCORES: result[PSCORES], SODS: fol}


def _smooth_step(fol, pts, delta_time):
    n = len(pts)
    if n == 0:
        # no person in this frame: nothing to update (the reference's zip is empty)
        return []
    J = fol.J
    dev = fol.state._eng.device
    buf = np.zeros((1, n, J, 4), np.float64)
    buf[0, :, :, :3] = np.asarray(pts, np.float64).reshape(n, -1, 3)[:, :J]
    out = torch.from_numpy(buf).to(dev)
    nsm = fol.state.run(out, torch.tensor([n], dtype=torch.int32, device=dev), delta_time)
    m = int(nsm.cpu()[0])
    o = out[0, :m, :, :3].cpu().numpy()
    return [np.ascontiguousarray(o[i]) for i in range(m)]
