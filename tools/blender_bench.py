#!/usr/bin/env python
"""Time the Blender control-point stage (snowtri_blender_run, snowtri_blender_smooth_run) on the cfg2 output shape
(131 072 frames x 1 person x 133 joints, float32 x y z score) and report the HBM roofline of blender_kernel.

Algorithmic bytes per person row: 28 joints x 12 B (x, y, z) read + 24 control points x 16 B + 4 B valid mask written
= 724 B.  With the (x, y, z, score) float4 layout and 32-byte sectors the kernel cannot fetch fewer than about
608 B per row (a 320-byte body block + eight scattered hand joints), i.e. ~996 B of DRAM traffic per row."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpu_leg(seconds=10.0):
    """The CPU restatement of Human_Triangulation_Blender (oracle/blender_oracle.py, one core, like the reference's
    single-threaded per-frame call) on a bounded sample of the same synthetic rows."""
    from oracle import blender_oracle as bo
    rng = np.random.default_rng(7)
    rows = (rng.random((512, 133, 3)) * 2.0).astype(np.float32).astype(np.float64)
    n, t0 = 0, time.perf_counter()
    with np.errstate(all="ignore"):
        while time.perf_counter() - t0 < seconds:
            bo.control_points(rows[n % 512])
            n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "persons/s", "cores": 1, "kind": "port",
            "sample": f"{n} person rows in {dt:.1f} s, NumPy restatement of blender.py:93-143 (oracle/blender_oracle.py)"}


if "--cpu-only" in sys.argv:
    print(json.dumps({"op": "Human_Triangulation_Blender (CPU)", "cpu_baseline": cpu_leg(), "host_cores": os.cpu_count()}))
    sys.exit(0)

import torch  # noqa: E402
from snowmocap_b200.blender import BlenderControl, BlenderSmoothState  # noqa: E402
from snowmocap_b200.triangulation import _util_engine  # noqa: E402

F, P, J = int(os.environ.get("BLENDER_BENCH_F", 131072)), int(os.environ.get("BLENDER_BENCH_P", 1)), 133
STEPS = 20
eng = _util_engine()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
out = torch.rand((F, P, J, 4), generator=g, device=dev) * 2.0
nout = torch.full((F,), P, dtype=torch.int32, device=dev)
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
bc = BlenderControl(eng)


def timed(fn, steps=STEPS, warm=5):
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


ms = timed(lambda: bc.run(out, nout))
rows = F * P
gbs = 724 * rows / (ms * 1e-3) / 1e9
ctrl, valid = bc.run(out, nout)
fzr = [[2.5, 0.75, 0.0]] * 24
st = BlenderSmoothState(eng, P, fzr)


def smooth():
    st.reset()
    st.run(ctrl, valid, nout, 1 / 30)


res_s = {}
if not os.environ.get("BLENDER_BENCH_NO_SMOOTH"):
    for name, chunked in (("chunked", True), ("scan", "scan"), ("sequential", False)):
        st.set_chunked(chunked)
        ms_s = timed(smooth, steps=5, warm=2)
        res_s[name] = {"ms": ms_s, "frames_per_s": F / (ms_s * 1e-3), "ns_per_frame_step": ms_s * 1e6 / F}
print(json.dumps({"op": "snowtri_blender_run", "F": F, "Pout": P, "J": J, "dtype": "f32 layout, f64 arithmetic",
                  "ms": ms, "persons_per_s": rows / (ms * 1e-3), "algorithmic_bytes_per_row": 724,
                  "algorithmic_GBs": gbs, "peak_GBs": peak, "frac_of_measured_hbm": gbs / peak,
                  "input_bytes": out.numel() * 4, "timing": "input larger than L2" if out.numel() * 4 > 126e6 else "input fits L2",
                  "nvcc_flags": os.environ.get("SNOWTRI_NVCC_FLAGS", ""),
                  "cpu_baseline": None if os.environ.get("BLENDER_BENCH_NO_CPU") or os.environ.get("BLENDER_BENCH_NO_SMOOTH") else cpu_leg(),
                  "smooth": {"op": "snowtri_blender_smooth_run (reset + one batch of F frames)", **res_s}}))
