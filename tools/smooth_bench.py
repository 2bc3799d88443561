#!/usr/bin/env python
"""Time snowtri_smooth_run on the cfg2 output shape (F frames, 1 person, 133 joints, float32 layout)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snowmocap_b200 as sv  # noqa: E402
import snowmocap_b200.triangulation  # noqa: E402,F401
from snowmocap_b200.engine import SmoothState  # noqa: E402

F, P, J = int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 1, 133
eng = sv.triangulation._util_engine()
out = torch.randn((F, P, J, 4), dtype=torch.float32, device="cuda")
nout = torch.ones((F,), dtype=torch.int32, device="cuda")
st = SmoothState(eng, P, J, 2.5, 0.75, 0.0)
for _ in range(2):
    st.run(out, nout)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    st.run(out, nout)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"op": "snowtri_smooth_run", "F": F, "P": P, "J": J, "ms": ms, "frames_per_s": F / (ms * 1e-3),
                  "keypoints_per_s": F * P * J / (ms * 1e-3), "ns_per_frame_step": ms * 1e6 / F}))
