#!/usr/bin/env python
"""Time the optional DLT mode on the cfg2 shape (4 cameras, 1 person, 133 joints) and report the HBM roofline."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snowmocap_b200 import synth  # noqa: E402
from snowmocap_b200.engine import TriangulationEngine  # noqa: E402

F, C, J = 131072, 4, 133
z = np.load(os.path.join(ROOT, "tests", "golden", "floor_rig.npz"))
rig = synth.Rig(z["K"], z["R"], z["t"]).subset(C)
eng = TriangulationEngine(rig.K, rig.R, rig.t, device=0, **synth.DEFAULT_PARAMS)
kpts, scores = synth.make_frames_torch(rig, F, 1, J, seed=1234, device=torch.device("cuda", 0))
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
res = {}
for name, acc in (("f32_accumulate", False), ("f64_accumulate", True)):
    for _ in range(5):
        out = eng.dlt(kpts, scores, accumulate_f64=acc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = eng.dlt(kpts, scores, accumulate_f64=acc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gbs = (12 * C + 16) * F * J / (ms * 1e-3) / 1e9
    res[name] = {"ms": ms, "keypoints_per_s": F * J / (ms * 1e-3), "algorithmic_GBs": gbs, "frac_of_measured_hbm": gbs / peak}
print(json.dumps({"op": "snowtri_dlt_run", "F": F, "C": C, "J": J, "peak_GBs": peak, **res}))
