#!/usr/bin/env python
"""Experiment: does the device's L2 fetch granularity limit (cudaLimitMaxL2FetchGranularity: 32/64/128 B) change
the DRAM over-fetch of blender_kernel's scattered 16-byte joint reads, and what does it do to the streaming
single-person kernel?  Usage: python tools/l2gran_probe.py   (prints one JSON line per setting)"""
import ctypes as ct
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from snowmocap_b200 import synth  # noqa: E402
from snowmocap_b200.blender import BlenderControl  # noqa: E402
from snowmocap_b200.engine import TriangulationEngine  # noqa: E402

torch.cuda.init()
torch.zeros(1, device="cuda")
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ct.CDLL(name)
        break
    except OSError:
        pass
LIMIT = 0x05  # cudaLimitMaxL2FetchGranularity


def get_limit():
    v = ct.c_size_t(0)
    rc = rt.cudaDeviceGetLimit(ct.byref(v), LIMIT)
    return int(v.value), rc


def timed(fn, steps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


dev = torch.device("cuda", 0)
z = np.load(os.path.join(ROOT, "tests", "golden", "floor_rig.npz"))
rig = synth.Rig(z["K"], z["R"], z["t"])
eng = TriangulationEngine(rig.K, rig.R, rig.t, device=0, precision="f32", **synth.DEFAULT_PARAMS)
F, J = 131072, 133
kpts, scores = synth.make_frames_torch(rig, F, 1, J, seed=1234, device=dev)
g = torch.Generator(device=dev).manual_seed(7)
out8 = torch.rand((F, 8, J, 4), generator=g, device=dev) * 2.0
bc = BlenderControl(eng)
print(json.dumps({"default_limit": get_limit()}))
for gran in (None, 32, 64, 128, None):
    rc = None
    if gran is not None:
        rc = rt.cudaDeviceSetLimit(LIMIT, ct.c_size_t(gran))
    ms_b = timed(lambda: bc.run(out8, None))
    ms_p = timed(lambda: eng.run(kpts, scores, None, Pout=1))
    print(json.dumps({"set": gran, "rc": rc, "limit_now": get_limit(), "blender_p8_ms": ms_b,
                      "blender_frac": 724 * F * 8 / (ms_b * 1e-3) / 1e9 / 6541.5, "p1_ms": ms_p,
                      "p1_frac": 64 * F * J / (ms_p * 1e-3) / 1e9 / 6541.5, "launch": eng.last_launch_info()}))
