#!/usr/bin/env python
"""Time the single-person kernel on a ring rig for a given camera count / tile size (experiments).
Usage: python tools/p1_variant_bench.py C [frames_per_tile] [precision]   (env SNOWTRI_JIT_MINB / SNOWTRI_JIT_DEFINES apply)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snowmocap_b200 import synth  # noqa: E402
from snowmocap_b200.engine import TriangulationEngine  # noqa: E402

C = int(sys.argv[1]); G = int(sys.argv[2]) if len(sys.argv) > 2 else 0; prec = sys.argv[3] if len(sys.argv) > 3 else "f32"
F, J = 1 << 16, 133
rig = synth.ring_rig(C)
eng = TriangulationEngine(rig.K, rig.R, rig.t, device=0, precision=prec, **synth.DEFAULT_PARAMS)
eng.set_tuning(G, 0, 0)
kpts, scores = synth.make_frames_torch(rig, F, 1, J, seed=1, device=torch.device("cuda", 0))
for _ in range(3):
    eng.run(kpts, scores, None, Pout=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.run(kpts, scores, None, Pout=1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"C": C, "tile": G, "precision": prec, "ms": ms, "keypoints_per_s": F * J / (ms * 1e-3),
                  "frac_hbm": (12 * C + 16) * F * J / (ms * 1e-3) / 1e9 / 6541.5, "launch": eng.last_launch_info(), "jit": eng.jit_status,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("SNOWTRI_")}}))
