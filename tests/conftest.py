import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not p.endswith("floor_rig.npz") and not os.path.basename(p).startswith(("smooth_", "blender_", "pipeline_", "big_")))


def smooth_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "smooth_*.npz")))


class Golden:
    """One fixture written by tests/golden/make_golden.py (outputs of the real reference)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.K, self.R, self.t = z["K"], z["R"], z["t"]
        self.kpts, self.scores, self.counts = z["kpts"], z["scores"], z["counts"]
        self.params = json.loads(str(z["params"]))
        self.F = self.kpts.shape[0]
        self.tri = [(z[f"tri_pts_{f}"], z[f"tri_ks_{f}"], z[f"tri_ps_{f}"]) for f in range(self.F)]
        self.con = [(z[f"con_pts_{f}"], z[f"con_ks_{f}"], z[f"con_ps_{f}"]) for f in range(self.F)]


class BigGolden:
    """One fixture written by tests/golden/make_golden_big.py: the real reference's condensed output at a full
    multi-person size (inputs, candidate count and candidate person scores, condensed persons)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.K, self.R, self.t = z["K"], z["R"], z["t"]
        self.kpts, self.scores, self.counts = z["kpts"], z["scores"], z["counts"]
        self.params = json.loads(str(z["params"]))
        self.F = self.kpts.shape[0]
        self.tri_ps = [z[f"tri_ps_{f}"] for f in range(self.F)]
        self.con = [(z[f"con_pts_{f}"], z[f"con_ks_{f}"], z[f"con_ps_{f}"]) for f in range(self.F)]
        self.seconds_per_frame = float(z["seconds_per_frame"])


def big_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "big_*.npz")))


@pytest.fixture(params=golden_names())
def golden(request):
    return Golden(request.param)


def floor_rig():
    from snowmocap_b200 import synth
    z = np.load(os.path.join(GOLDEN, "floor_rig.npz"))
    return synth.Rig(z["K"], z["R"], z["t"])


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))
