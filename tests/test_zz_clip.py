"""``snowtri_clip_run``: main.py:55-87 for a whole clip in one C call.  (Named to run last: the entry point was added
after the round's GPU budget was spent, so its first GPU run is the driver's.)"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, floor_rig, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    return torch


def _setup(torch, precision):
    from snowmocap_b200.blender import BlenderSmoothState
    from snowmocap_b200.engine import SmoothState, TriangulationEngine
    g = np.load(os.path.join(GOLDEN, "pipeline_main.npz"))
    p = json.loads(str(g["params"]))
    fzr = np.load(os.path.join(GOLDEN, "blender_profiles.npz"))["fzr"]
    rig = floor_rig()
    eng = TriangulationEngine(rig.K, rig.R, rig.t, device=0, precision=precision, kst=p["kst"], ast=p["ast"],
                              dthr=p["dthr"], cond_tol=p["cond_tol"], num_tol=p["num_tol"], score_tol=p["score_tol"],
                              center=p["center"])
    sm = SmoothState(eng, 1, 133, p["smooth_f"], p["smooth_z"], p["smooth_r"])
    bs = BlenderSmoothState(eng, 1, fzr)
    kp, sc = torch.from_numpy(g["kpts"]).cuda(), torch.from_numpy(g["scores"]).cuda()
    return g, p, eng, sm, bs, kp, sc


@pytest.mark.parametrize("precision,tol", [("f64", 1e-6), ("f32", 1e-4)])
def test_clip_run_matches_reference_and_the_separate_calls(torch_cuda, precision, tol):
    torch = torch_cuda
    from snowmocap_b200.blender import BlenderControl, run_clip
    g, p, eng, sm, bs, kp, sc = _setup(torch, precision)
    res = run_clip(eng, kp, sc, None, sm, bs, Pout=1, delta_time=p["smooth_delta_time"])
    torch.cuda.synchronize()
    F = kp.shape[0]
    want = np.stack([g[f"ctrl_{f}"][0] for f in range(F)])
    joints = np.stack([g[f"joints_{f}"][0] for f in range(F)])
    assert (res["nfinal"].cpu().numpy() == 1).all() and (res["nsmooth"].cpu().numpy() == 1).all()
    assert rel_l2(res["out"].cpu().numpy()[:, 0, :, :3], joints) < tol
    assert rel_l2(res["ctrl"].cpu().numpy()[:, 0], want) < tol
    # the same stages called one by one give bit-identical buffers
    sm.reset()
    bs.reset()
    one = eng.run(kp, sc, None, Pout=1)
    nsm = sm.run(one["out"], one["nout"], p["smooth_delta_time"])
    ctrl, valid = BlenderControl(eng).run(one["out"], nsm)
    nfin = bs.run(ctrl, valid, nsm, p["smooth_delta_time"])
    torch.cuda.synchronize()
    assert torch.equal(one["out"], res["out"]) and torch.equal(ctrl, res["ctrl"]) and torch.equal(valid, res["valid"])
    assert torch.equal(nsm, res["nsmooth"]) and torch.equal(nfin, res["nfinal"])
    sm.close(); bs.close(); eng.close()


def test_clip_run_without_the_smoothing_stages(torch_cuda):
    torch = torch_cuda
    from snowmocap_b200.blender import BlenderControl, run_clip
    g, p, eng, sm, bs, kp, sc = _setup(torch, "f64")
    res = run_clip(eng, kp, sc, None, None, None, Pout=1)
    one = eng.run(kp, sc, None, Pout=1)
    ctrl, valid = BlenderControl(eng).run(one["out"], one["nout"])
    torch.cuda.synchronize()
    assert res["nsmooth"] is None and res["nfinal"] is None
    assert torch.equal(one["out"], res["out"]) and torch.equal(ctrl, res["ctrl"]) and torch.equal(valid, res["valid"])
    with pytest.raises(IndexError):
        run_clip(eng, kp[:, :, :, :17].contiguous(), sc[:, :, :, :17].contiguous())
    sm.close(); bs.close(); eng.close()
