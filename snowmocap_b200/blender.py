"""Drop-in mirror of the per-frame functions of ``snowvision.blender`` (reference blender.py:93-187) and the batch
API beneath them.

``Human_Triangulation_Blender``, ``Human_Triangulation_Blender_Smooth``, ``Human_Triangulation_To_Blender_Result``
and ``save_blender_result`` keep the reference's names, signatures, dict keys and JSON schema (``main.py:80-87,104``);
the arithmetic runs in ``snowtri_blender_run`` / ``snowtri_blender_smooth_run`` (``csrc/snowtri_blender.cu``) through
the C ABI.  There is no CPU fallback.

Errors, like the reference: a person with fewer than 130 joints raises ``IndexError`` (``person[129]``,
blender.py:103); an armature-profile name that is not one of the 24 control points raises ``NameError`` (the
reference ``eval``s the name, blender.py:133); a NaN root rotation raises ``numpy.linalg.LinAlgError`` (SciPy's SVD
inside ``Rotation.from_matrix``, util.py:26-28).  The batch API (``BlenderControl``) reports that last case through
the control point's valid bit instead of raising.
"""
from __future__ import annotations

import copy
import ctypes as ct
import json

import numpy as np
import torch

from . import _lib
from .engine import _ptr, _stream

CONTROL_POINTS = ("root_position", "root_rotation", "clavicle_r_ik", "clavicle_l_ik", "arm_r_ik", "arm_r_pole",
                  "arm_l_ik", "arm_l_pole", "leg_r_ik", "leg_r_pole", "leg_l_ik", "leg_l_pole", "hand_r_ik",
                  "hand_r_pole", "hand_l_ik", "hand_l_pole", "foot_r_ik", "foot_r_pole", "foot_l_ik", "foot_l_pole",
                  "chest_ik", "chest_pole", "head_ik", "head_pole")
_INDEX = {n: i for i, n in enumerate(CONTROL_POINTS)}
_LEN = {n: 4 if n == "root_rotation" else 3 for n in CONTROL_POINTS}

CPOINTS = "blender_armature_control_points"
CSCORES = "blender_armature_control_points_scores"
SODS = "second_order_dynamics"


class BlenderControl:
    """Batch API: control points of every person row of a ``TriangulationEngine.run`` / ``condense`` result."""

    def __init__(self, engine):
        self._eng, self._lib = engine, engine._lib

    def run(self, out, nout=None):
        """``out`` (F,Pout,J,4) float32/float64 cuda tensor (x, y, z, score), ``nout`` (F,) int32 or None ->
        ``ctrl`` (F,Pout,24,4) of the same dtype and ``valid`` (F,Pout) int32 bit masks (bit k = control point k)."""
        if out.dim() != 4 or out.shape[-1] != 4 or not out.is_cuda or not out.is_contiguous():
            raise ValueError("out must be a contiguous cuda tensor of shape (F,Pout,J,4)")
        if out.dtype not in (torch.float32, torch.float64):
            raise ValueError("out must be float32 or float64")
        F, Pout, J, _ = out.shape
        if J < 130:
            raise IndexError(f"index 129 is out of bounds for axis 0 with size {J}")
        if nout is not None and (nout.shape != (F,) or nout.dtype != torch.int32 or not nout.is_cuda):
            raise ValueError("nout must be an int32 cuda tensor of shape (F,)")
        ctrl = torch.empty((F, Pout, 24, 4), dtype=out.dtype, device=out.device)
        valid = torch.empty((F, Pout), dtype=torch.int32, device=out.device)
        fn = self._lib.snowtri_blender_run if out.dtype == torch.float32 else self._lib.snowtri_blender_run_f64
        with torch.cuda.device(self._eng.device):
            _lib.check(fn(self._eng._h, _ptr(out), _ptr(nout) if nout is not None else None, F, Pout, J, _ptr(ctrl),
                          _ptr(valid), _stream()), self._eng._h)
        return ctrl, valid


class BlenderSmoothState:
    """Device-resident followers of ``Human_Triangulation_Blender_Smooth`` for one clip (blender.py:145-178).
    ``fzr`` (24,3): f, z, r per control point in ``CONTROL_POINTS`` order, or the smooth-profile dict."""

    def __init__(self, engine, max_persons, fzr):
        if isinstance(fzr, dict):
            fzr = [fzr[n] for n in CONTROL_POINTS]
        self._fzr = np.ascontiguousarray(np.asarray(fzr, np.float64).reshape(24, 3))
        self._eng, self._lib, self._s = engine, engine._lib, ct.c_void_p()
        self.max_persons = int(max_persons)
        with torch.cuda.device(engine.device):
            _lib.check(self._lib.snowtri_blender_smooth_create(engine._h, ct.byref(self._s), self.max_persons,
                                                               self._fzr.ctypes.data), engine._h)

    def set_chunked(self, enabled=True):
        """Long batches run as parallel chunks (default: one pass with a warm-up per chunk when every follower forgets
        its state within 256 frames, else the chunk scan); ``"scan"`` forces the chunk scan, False the sequential
        kernel."""
        mode = 2 if enabled == "scan" else (1 if enabled else 0)
        _lib.check(self._lib.snowtri_blender_smooth_set_chunked(self._s, mode), self._eng._h)

    def reset(self):
        _lib.check(self._lib.snowtri_blender_smooth_reset(self._eng._h, self._s, _stream()), self._eng._h)

    def run(self, ctrl, valid, nout, delta_time=1 / 30):
        """Smooth ``ctrl`` (F,Pout,24,4) in place, frames in order; returns nsmooth (F,) int32."""
        if ctrl.dim() != 4 or tuple(ctrl.shape[2:]) != (24, 4) or not ctrl.is_cuda or not ctrl.is_contiguous():
            raise ValueError("ctrl must be a contiguous cuda tensor of shape (F,Pout,24,4)")
        if ctrl.dtype not in (torch.float32, torch.float64):
            raise ValueError("ctrl must be float32 or float64")
        F, Pout = ctrl.shape[:2]
        if valid.shape != (F, Pout) or valid.dtype != torch.int32 or not valid.is_cuda or not valid.is_contiguous():
            raise ValueError("valid must be a contiguous int32 cuda tensor of shape (F,Pout)")
        if nout.shape != (F,) or nout.dtype != torch.int32 or not nout.is_cuda:
            raise ValueError("nout must be an int32 cuda tensor of shape (F,)")
        nsm = torch.empty((F,), dtype=torch.int32, device=ctrl.device)
        fn = self._lib.snowtri_blender_smooth_run if ctrl.dtype == torch.float32 \
            else self._lib.snowtri_blender_smooth_run_f64
        with torch.cuda.device(self._eng.device):
            _lib.check(fn(self._eng._h, self._s, _ptr(ctrl), _ptr(valid), _ptr(nout), _ptr(nsm), F, Pout,
                          float(delta_time), _stream()), self._eng._h)
        return nsm

    def close(self):
        if self._s:
            self._lib.snowtri_blender_smooth_destroy(self._s)
            self._s = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_clip(engine, kpts, scores, counts=None, smooth_state=None, blender_smooth_state=None, Pout=None,
             delta_time=1 / 30):
    """``snowtri_clip_run``: the per-frame body of the reference's main.py (:55-87) for a whole clip in one C call --
    triangulate + condense, ``Human_Triangulation_Smooth`` (``smooth_state``: an ``engine.SmoothState`` or None),
    Blender control points, ``Human_Triangulation_Blender_Smooth`` (``blender_smooth_state`` or None).  Device
    tensors in, device tensors out: dict(out, pscores, nout, nsmooth, ctrl, valid, nfinal)."""
    F, C, P, J = engine._check_inputs(kpts, scores, counts)
    if J < 130:
        raise IndexError(f"index 129 is out of bounds for axis 0 with size {J}")
    Pout = P if Pout is None else int(Pout)
    dev = engine.device
    res = {"out": torch.empty((F, Pout, J, 4), dtype=torch.float32, device=dev),
           "pscores": torch.empty((F, Pout), dtype=torch.float32, device=dev),
           "nout": torch.empty((F,), dtype=torch.int32, device=dev),
           "nsmooth": torch.empty((F,), dtype=torch.int32, device=dev) if smooth_state is not None else None,
           "ctrl": torch.empty((F, Pout, 24, 4), dtype=torch.float32, device=dev),
           "valid": torch.empty((F, Pout), dtype=torch.int32, device=dev),
           "nfinal": torch.empty((F,), dtype=torch.int32, device=dev) if blender_smooth_state is not None else None}
    with torch.cuda.device(dev):
        _lib.check(engine._lib.snowtri_clip_run(
            engine._h, smooth_state._s if smooth_state is not None else None,
            blender_smooth_state._s if blender_smooth_state is not None else None,
            _ptr(kpts), _ptr(scores), _ptr(counts), F, P, J, Pout, _ptr(res["out"]), _ptr(res["pscores"]),
            _ptr(res["nout"]), _ptr(res["nsmooth"]), _ptr(res["ctrl"]), _ptr(res["valid"]), _ptr(res["nfinal"]),
            float(delta_time), _stream()), engine._h)
    return res


def _engine():
    from .triangulation import _util_engine
    return _util_engine()


def _check_profile(profile):
    for name in profile.keys():
        if name not in _INDEX:          # the reference eval()s the name (blender.py:133)
            raise NameError(f"name '{name}' is not defined")


def save_blender_result(blender_result, file_path):
    """Reference blender.py:7-9."""
    with open(file_path, "w") as outfile:
        outfile.write(json.dumps(blender_result, indent=4))


def Human_Triangulation_Blender(result, blender_armature_profile):
    """24 armature control points per person; reference blender.py:93-143."""
    out = {CPOINTS: [], CSCORES: []}
    persons = list(zip(result["hrnet_triangulate_points"], result["hrnet_triangulate_keypoint_scores"]))
    if not persons:
        return out
    n = len(persons)
    J = min(int(np.asarray(p).shape[0]) for p, _ in persons)
    if J < 130:
        raise IndexError(f"index 129 is out of bounds for axis 0 with size {J}")
    _check_profile(blender_armature_profile)
    buf = np.zeros((1, n, J, 4), np.float64)
    for i, (p, _) in enumerate(persons):
        buf[0, i, :, :3] = np.asarray(p, np.float64).reshape(-1, 3)[:J]
    eng = _engine()
    ctrl, valid = BlenderControl(eng).run(torch.from_numpy(buf).to(eng.device))
    ctrl, valid = ctrl[0].cpu().numpy(), valid[0].cpu().numpy()
    if not np.all((valid >> _INDEX["root_rotation"]) & 1):
        raise np.linalg.LinAlgError("SVD did not converge")   # SciPy's from_matrix on a NaN frame (util.py:26-28)
    for i in range(n):
        cp = copy.deepcopy(blender_armature_profile)
        sc = copy.deepcopy(blender_armature_profile)
        for name in blender_armature_profile.keys():
            k = _INDEX[name]
            cp[name] = ctrl[i, k, :_LEN[name]].tolist()
            sc[name] = int((valid[i] >> k) & 1)
        out[CPOINTS].append(cp)
        out[CSCORES].append(sc)
    return out


class _ControlFollowers:
    """What this package stores under ``'second_order_dynamics'`` (the reference keeps a list of dicts of
    ``SecondOrderDynamic`` objects there, blender.py:165-175)."""

    def __init__(self, n0, fzr):
        self.n0 = n0
        self.state = BlenderSmoothState(_engine(), max(n0, 1), fzr) if n0 > 0 else None

    def __len__(self):
        return self.n0


def _pack_control(cp_list, sc_list, names):
    n = len(cp_list)
    ctrl = np.zeros((1, n, 24, 4), np.float64)
    valid = np.zeros((1, n), np.int32)
    for i, (cp, sc) in enumerate(zip(cp_list, sc_list)):
        for name in names:
            k = _INDEX[name]
            ctrl[0, i, k, :_LEN[name]] = np.asarray(cp[name], np.float64)
            if sc[name]:
                valid[0, i] |= 1 << k
    return ctrl, valid


def Human_Triangulation_Blender_Smooth(current_blender_result, blender_armature_profile, blender_smooth_profile,
                                       previous_blender_result=None, delta_time=1 / 30):
    """Second-order-dynamics smoothing of the control points; reference blender.py:145-178 (main.py:81-86)."""
    _check_profile(blender_armature_profile)
    names = list(blender_armature_profile.keys())
    cps, scs = current_blender_result[CPOINTS], current_blender_result[CSCORES]
    if not isinstance(previous_blender_result, dict):
        n0 = min(len(cps), len(scs))
        # control points outside the armature profile have no follower in the reference; give them a neutral one
        fzr = [blender_smooth_profile[n] if n in blender_armature_profile else [1.0, 1.0, 0.0] for n in CONTROL_POINTS]
        fol = _ControlFollowers(n0, fzr)
        if fol.state is not None:
            _control_step(fol, cps[:n0], scs[:n0], names, delta_time)   # seeds the followers
        return {CPOINTS: cps, CSCORES: scs, SODS: fol}
    fol = previous_blender_result[SODS]
    m = min(len(cps), len(scs), len(fol))
    smoothed = _control_step(fol, cps[:m], scs[:m], names, delta_time) if m > 0 else []
    out = []
    for i in range(m):
        cp = copy.deepcopy(blender_armature_profile)
        for name in names:
            cp[name] = smoothed[i, _INDEX[name], :_LEN[name]].tolist()
        out.append(cp)
    return {CPOINTS: out, CSCORES: scs, SODS: fol}


def _control_step(fol, cps, scs, names, delta_time):
    ctrl, valid = _pack_control(cps, scs, names)
    dev = fol.state._eng.device
    c = torch.from_numpy(ctrl).to(dev)
    nsm = fol.state.run(c, torch.from_numpy(valid).to(dev),
                        torch.tensor([ctrl.shape[1]], dtype=torch.int32, device=dev), delta_time)
    return c[0, :int(nsm.cpu()[0])].cpu().numpy()


def Human_Triangulation_To_Blender_Result(result):
    """Reference blender.py:180-187: ``{'armature': [...], 'score': [...]}`` (zip truncates to the shorter list)."""
    blender_result = {"armature": [], "score": []}
    for control_points, control_points_scores in zip(result[CPOINTS], result[CSCORES]):
        blender_result["armature"].append(control_points)
        blender_result["score"].append(control_points_scores)
    return blender_result


def clip_to_blender_result_list(ctrl, valid, nsmooth, blender_armature_profile):
    """Batch counterpart of ``main.py:87``: the per-frame ``Human_Triangulation_To_Blender_Result`` dicts of a whole
    clip from the arrays of ``BlenderControl.run`` / ``BlenderSmoothState.run`` -- ``ctrl`` (F,Pout,24,4), ``valid``
    (F,Pout) bit masks, ``nsmooth`` (F,) persons per frame (tensors or arrays; copied to the host once).  The list can
    be handed to ``save_blender_result`` and is what the reference appends frame after frame (blender.py:180-187)."""
    _check_profile(blender_armature_profile)
    ctrl = ctrl.detach().cpu().numpy() if isinstance(ctrl, torch.Tensor) else np.asarray(ctrl)
    valid = valid.detach().cpu().numpy() if isinstance(valid, torch.Tensor) else np.asarray(valid)
    nsmooth = nsmooth.detach().cpu().numpy() if isinstance(nsmooth, torch.Tensor) else np.asarray(nsmooth)
    names = list(blender_armature_profile.keys())
    frames = []
    for f in range(ctrl.shape[0]):
        armature, score = [], []
        for k in range(min(int(nsmooth[f]), ctrl.shape[1])):
            armature.append({n: ctrl[f, k, _INDEX[n], :_LEN[n]].astype(np.float64).tolist() for n in names})
            score.append({n: int((int(valid[f, k]) >> _INDEX[n]) & 1) for n in names})
        frames.append({"armature": armature, "score": score})
    return frames
