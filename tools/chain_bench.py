#!/usr/bin/env python
"""Experiment: run the stages after the triangulation (snowtri_smooth_run -> snowtri_blender_run ->
snowtri_blender_smooth_run) over the cfg2 batch in frame chunks small enough for the 126 MB L2, so that each stage
finds its predecessor's output (and the smoothing's second pass its own input) in L2 instead of DRAM.
Prints one JSON line per chunk size; `chunk == F` is the unchunked chain."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snowmocap_b200 import synth  # noqa: E402
from snowmocap_b200.blender import BlenderControl, BlenderSmoothState  # noqa: E402
from snowmocap_b200.engine import SmoothState, TriangulationEngine  # noqa: E402

F, J = 131072, 133
dev = torch.device("cuda", 0)
z = np.load(os.path.join(ROOT, "tests", "golden", "floor_rig.npz"))
rig = synth.Rig(z["K"], z["R"], z["t"])
eng = TriangulationEngine(rig.K, rig.R, rig.t, device=0, precision="f32", **synth.DEFAULT_PARAMS)
kpts, scores = synth.make_frames_torch(rig, F, 1, J, seed=1234, device=dev)
res = eng.run(kpts, scores, None, Pout=1)
base, nout = res["out"].clone(), res["nout"]
pts = base.clone()
sm = SmoothState(eng, 1, J, 2.5, 0.75, 0.0)
bc = BlenderControl(eng)
bs = BlenderSmoothState(eng, 1, [[2.5, 0.75, 0.0]] * 24)
ctrl = torch.empty((F, 1, 24, 4), dtype=torch.float32, device=dev)
valid = torch.empty((F, 1), dtype=torch.int32, device=dev)
lib, h = eng._lib, eng._h
st = torch.cuda.current_stream().cuda_stream


def chain(chunk, with_tri):
    sm.reset()
    bs.reset()
    for s in range(0, F, chunk):
        e = min(F, s + chunk)
        if with_tri:
            eng.run(kpts[s:e], scores[s:e], None, Pout=1, out={"out": pts[s:e], "pscores": res["pscores"][s:e], "nout": nout[s:e]})
        nsm = sm.run(pts[s:e], nout[s:e], 1 / 30)
        lib.snowtri_blender_run(h, pts[s:e].data_ptr(), nsm.data_ptr(), e - s, 1, J, ctrl[s:e].data_ptr(), valid[s:e].data_ptr(), st)
        bs.run(ctrl[s:e], valid[s:e], nsm, 1 / 30)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ref_ctrl = None
for with_tri in (False, True):
    if with_tri:
        eng.set_jit("always")
    for chunk in (F, 65536, 32768, 16384, 8192, 4096):
        if not with_tri:
            pts.copy_(base)
        ms = timed(lambda: chain(chunk, with_tri))
        # correctness of the chunked chain against the unchunked one (fresh input)
        pts.copy_(base)
        chain(chunk, with_tri)
        torch.cuda.synchronize()
        if ref_ctrl is None:
            ref_ctrl = ctrl.clone()
        err = float((ctrl - ref_ctrl).abs().max())
        print(json.dumps({"with_triangulation": with_tri, "chunk_frames": chunk, "chunk_MB_of_joints": chunk * J * 16 / 1e6,
                          "ms_per_131072_frames": ms, "max_abs_diff_vs_unchunked": err, "jit": eng.jit_status}), flush=True)
