"""Blender control-point stage (reference snowvision/blender.py:93-187).

CPU tests pin ``oracle/blender_oracle.py`` to golden vectors produced by the REAL reference
(``tests/golden/make_golden_blender.py``).  GPU tests compare ``snowtri_blender_run`` / ``snowtri_blender_smooth_run``
(through the C ABI) and the drop-in functions with those golden vectors and with the oracle on seeded inputs.

Tolerances: float64 layout 1e-12 relative (same arithmetic, different operation order); float32 layout 1e-6
(float64 arithmetic, float32 store); north_star bound 1e-4.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2
from oracle import blender_oracle as bo

TOL_F64 = 1e-12
TOL_F32 = 1e-6


def _golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def _profiles():
    z = _golden("blender_profiles")
    names = [str(n) for n in z["names"]]
    return names, z["fzr"]


def _persons(rng, n):
    c = np.concatenate([rng.uniform(-2, 2, (n, 1, 2)), np.zeros((n, 1, 1))], axis=2)
    j = np.concatenate([rng.uniform(-0.4, 0.4, (n, 133, 2)), rng.uniform(0, 1.8, (n, 133, 1))], axis=2)
    return (c + j).astype(np.float32)


# ---- CPU: the oracle against the real reference's outputs ---------------------------------------------------
def test_oracle_control_point_order_is_the_shipped_profile():
    names, fzr = _profiles()
    assert tuple(names) == bo.CONTROL_POINTS
    assert fzr.shape == (24, 3)


def test_oracle_control_points_match_reference_golden():
    g = _golden("blender_points")
    ctrl, valid = bo.control_points_batch(g["pts"])
    assert np.array_equal(valid, g["valid"].astype(bool))
    assert (~valid).sum() >= 5, "fixture must exercise invalid control points"
    assert np.array_equal(np.isnan(ctrl), np.isnan(g["ctrl"]))
    np.testing.assert_allclose(ctrl[valid], g["ctrl"][valid], rtol=0, atol=1e-14)


def test_oracle_quaternion_matches_scipy():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(5)
    for _ in range(50):
        x, y = rng.normal(size=3), rng.normal(size=3)
        x, y = x / np.linalg.norm(x), y / np.linalg.norm(y)
        z = np.cross(x, y) / np.linalg.norm(np.cross(x, y))
        R = np.array([x, y, z]).T
        np.testing.assert_allclose(bo.rotation_matrix_to_quaternion(R), Rotation.from_matrix(R).as_quat(), atol=1e-14)


def test_oracle_smooth_matches_reference_golden():
    g = _golden("blender_smooth")
    _, fzr = _profiles()
    o = bo.BlenderSmoothOracle(fzr)
    for t, n in enumerate(g["counts"]):
        ctrl, valid = bo.control_points_batch(g["clip"][t, :n])
        y = o.step(ctrl, valid, float(g["dt"]))
        want = g[f"ctrl_{t}"]
        assert y.shape == want.shape, t
        assert np.array_equal(np.isnan(y), np.isnan(want)), t
        np.testing.assert_allclose(np.nan_to_num(y), np.nan_to_num(want), rtol=0, atol=1e-13)


def test_oracle_control_points_are_rigid_motion_equivariant():
    """Moving the person by a rotation Q and a translation t moves every position by the same motion and composes
    the root rotation with Q (a property of blender.py:11-96 that needs no reference run)."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(12)
    pts = _persons(rng, 6).astype(np.float64)
    pos = [k for k in range(24) if k != bo.ROOT_ROTATION]
    for n in range(6):
        Q = Rotation.random(random_state=100 + n)
        t = rng.uniform(-3, 3, 3)
        c0, v0 = bo.control_points(pts[n])
        c1, v1 = bo.control_points(Q.apply(pts[n]) + t)
        assert v0.all() and v1.all()
        np.testing.assert_allclose(c1[pos, :3], Q.apply(c0[pos, :3]) + t, atol=1e-11)
        q0 = Rotation.from_quat(np.roll(c0[1], -1))     # stored (w, x, y, z) -> SciPy's (x, y, z, w)
        q1 = Rotation.from_quat(np.roll(c1[1], -1))
        np.testing.assert_allclose(q1.as_matrix(), (Q * q0).as_matrix(), atol=1e-11)


def test_oracle_smoothing_is_linear_in_the_inputs():
    """With every control point valid the followers are a linear filter: step(a*x + b*y) = a*step(x) + b*step(y)."""
    rng = np.random.default_rng(13)
    _, fzr = _profiles()
    x = rng.normal(size=(30, 2, 24, 4))
    y = rng.normal(size=(30, 2, 24, 4))
    valid = np.ones((2, 24), bool)
    ox, oy, oz = bo.BlenderSmoothOracle(fzr), bo.BlenderSmoothOracle(fzr), bo.BlenderSmoothOracle(fzr)
    for t in range(30):
        sx, sy = ox.step(x[t], valid, 1 / 30), oy.step(y[t], valid, 1 / 30)
        sz = oz.step(2.0 * x[t] - 0.5 * y[t], valid, 1 / 30)
        np.testing.assert_allclose(sz, 2.0 * sx - 0.5 * sy, atol=1e-11)


# ---- CPU: host logic of the drop-in names (no device is touched on these paths) --------------------------------
def test_dropin_host_checks_run_before_the_device_is_needed():
    from snowmocap_b200 import Human_Triangulation_Blender, Human_Triangulation_To_Blender_Result
    names, _ = _profiles()
    profile = {n: [] for n in names}
    empty = {"hrnet_triangulate_points": [], "hrnet_triangulate_keypoint_scores": []}
    assert Human_Triangulation_Blender(empty, profile) == {"blender_armature_control_points": [],
                                                          "blender_armature_control_points_scores": []}
    short = {"hrnet_triangulate_points": [np.zeros((30, 3))], "hrnet_triangulate_keypoint_scores": [np.ones(30)]}
    with pytest.raises(IndexError):          # the reference indexes person[129] (blender.py:103)
        Human_Triangulation_Blender(short, profile)
    whole = {"hrnet_triangulate_points": [np.zeros((133, 3))], "hrnet_triangulate_keypoint_scores": [np.ones(133)]}
    with pytest.raises(NameError):           # the reference eval()s the profile's names (blender.py:133)
        Human_Triangulation_Blender(whole, {"tail_ik": []})
    # zip() truncation of To_Blender_Result (blender.py:183): scores of persons without a follower are dropped
    res = Human_Triangulation_To_Blender_Result({"blender_armature_control_points": [{"a": [0, 0, 0]}],
                                                 "blender_armature_control_points_scores": [{"a": 1}, {"a": 1}]})
    assert res == {"armature": [{"a": [0, 0, 0]}], "score": [{"a": 1}]}


def test_control_point_table_matches_oracle_and_header():
    import re
    from snowmocap_b200.blender import CONTROL_POINTS, _pack_control
    assert CONTROL_POINTS == bo.CONTROL_POINTS
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "snowtri.h")).read()
    listed = re.findall(r"(\d+) ([a-z_]+_(?:position|rotation|ik|pole))", header)
    assert [n for _, n in sorted(((int(i), n) for i, n in listed))] == list(CONTROL_POINTS)
    cps = [{"root_rotation": [1.0, 0.0, 0.0, 0.0], "head_ik": [1.0, 2.0, 3.0]}]
    scs = [{"root_rotation": 1, "head_ik": 0}]
    ctrl, valid = _pack_control(cps, scs, ["root_rotation", "head_ik"])
    assert ctrl.shape == (1, 1, 24, 4) and valid.tolist() == [[1 << 1]]
    assert ctrl[0, 0, 1].tolist() == [1.0, 0.0, 0.0, 0.0] and ctrl[0, 0, 22].tolist() == [1.0, 2.0, 3.0, 0.0]


def test_clip_to_blender_result_list_schema(tmp_path):
    """Arrays of a clip -> the reference's per-frame {'armature': [...], 'score': [...]} list (blender.py:180-187)."""
    from snowmocap_b200 import clip_to_blender_result_list, save_blender_result
    g = _golden("blender_smooth")
    names, _ = _profiles()
    F = len(g["counts"])
    ctrl = np.zeros((F, 3, 24, 4))
    valid = np.zeros((F, 3), np.int32)
    nsm = np.zeros(F, np.int32)
    for t in range(F):
        m = g[f"ctrl_{t}"].shape[0]
        nsm[t] = m
        ctrl[t, :m] = g[f"ctrl_{t}"]
        valid[t, :m] = (g[f"valid_{t}"].astype(np.int64) << np.arange(24)).sum(-1)
    frames = clip_to_blender_result_list(ctrl, valid, nsm, {n: [] for n in names})
    assert len(frames) == F
    for t in range(F):
        assert set(frames[t]) == {"armature", "score"}
        assert len(frames[t]["armature"]) == len(frames[t]["score"]) == nsm[t]
        for k in range(nsm[t]):
            assert list(frames[t]["armature"][k]) == names
            for i, n in enumerate(names):
                v = frames[t]["armature"][k][n]
                assert len(v) == (4 if n == "root_rotation" else 3)
                assert np.array_equal(np.asarray(v), g[f"ctrl_{t}"][k, i, :len(v)], equal_nan=True)
                assert frames[t]["score"][k][n] == int(g[f"valid_{t}"][k, i])
    path = tmp_path / "clip.json"
    save_blender_result(frames, str(path))
    assert len(json.loads(path.read_text())) == F
    with pytest.raises(NameError):
        clip_to_blender_result_list(ctrl, valid, nsm, {"tail_ik": []})


# ---- GPU ----------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    return torch


def _util_engine():
    from snowmocap_b200.triangulation import _util_engine
    return _util_engine()


def _masks(valid_bits):
    return ((np.asarray(valid_bits)[..., None] >> np.arange(24)) & 1).astype(bool)


def _run_control(torch, pts4, nout=None):
    from snowmocap_b200.blender import BlenderControl
    eng = _util_engine()
    out = torch.from_numpy(np.ascontiguousarray(pts4)).cuda()
    n = None if nout is None else torch.from_numpy(np.asarray(nout, np.int32)).cuda()
    ctrl, valid = BlenderControl(eng).run(out, n)
    torch.cuda.synchronize()
    return ctrl.cpu().numpy(), valid.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(np.float64, TOL_F64), (np.float32, TOL_F32)])
def test_control_points_match_reference_golden(torch_cuda, dtype, tol):
    g = _golden("blender_points")
    N = g["pts"].shape[0]
    buf = np.zeros((1, N, 133, 4), dtype)
    buf[0, :, :, :3] = g["pts"]            # float32-representable values: the same input in both layouts
    buf[0, :, :, 3] = 1.0
    ctrl, vbits = _run_control(torch_cuda, buf)
    valid = _masks(vbits[0])
    want_valid = g["valid"].astype(bool)
    assert np.array_equal(valid, want_valid)
    assert np.array_equal(np.isnan(ctrl[0]), np.isnan(g["ctrl"]))
    assert rel_l2(ctrl[0][want_valid], g["ctrl"][want_valid]) < tol
    assert not ctrl[0, :, [k for k in range(24) if k != 1], 3].any(), "positions carry a zero fourth float"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(np.float64, TOL_F64), (np.float32, TOL_F32)])
@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (65, 5), (1000, 2)])
def test_control_points_vs_oracle_seeded(torch_cuda, dtype, tol, shape):
    """Tile edges (rows not a multiple of 32), empty slots (nout < Pout), zeroed joints."""
    F, Pout = shape
    rng = np.random.default_rng(F * 31 + Pout)
    pts = _persons(rng, F * Pout).reshape(F, Pout, 133, 3)
    kill = rng.random((F, Pout)) < 0.1
    pts[kill, 91] = pts[kill, 96]                     # degenerate hand -> NaN pole
    nout = rng.integers(0, Pout + 1, F).astype(np.int32)
    buf = np.zeros((F, Pout, 133, 4), dtype)
    buf[..., :3] = pts
    ctrl, vbits = _run_control(torch_cuda, buf, nout)
    valid = _masks(vbits)
    want, want_valid = bo.control_points_batch(pts.reshape(-1, 133, 3).astype(np.float64))
    want, want_valid = want.reshape(F, Pout, 24, 4), want_valid.reshape(F, Pout, 24)
    present = np.arange(Pout)[None, :] < nout[:, None]
    assert np.array_equal(valid[present], want_valid[present])
    assert not valid[~present].any() and not ctrl[~present].any(), "empty slots: zeros, nothing valid"
    ok = present[..., None] & want_valid
    if ok.any():
        assert rel_l2(ctrl[ok], want[ok]) < tol
        # the quaternion on its own (unit norm, not dominated by metre-sized positions)
        q_ok = ok[..., 1]
        assert np.abs(ctrl[..., 1, :][q_ok] - want[..., 1, :][q_ok]).max() < (1e-11 if dtype == np.float64 else 1e-6)
    assert np.array_equal(np.isnan(ctrl[present]), np.isnan(want[present]))


@pytest.mark.gpu
def test_control_points_nan_root_rotation_is_flagged_not_raised(torch_cuda):
    rng = np.random.default_rng(3)
    pts = _persons(rng, 4).astype(np.float64)
    pts[1, 11] = pts[1, 12]                           # zero pelvis axis: SciPy raises in the reference
    buf = np.zeros((1, 4, 133, 4))
    buf[0, :, :, :3] = pts
    ctrl, vbits = _run_control(torch_cuda, buf)
    valid = _masks(vbits[0])
    assert not valid[1, 1] and np.isnan(ctrl[0, 1, 1]).all()
    assert valid[[0, 2, 3], 1].all()


@pytest.mark.gpu
def test_control_points_need_wholebody_joints(torch_cuda):
    with pytest.raises(IndexError):
        _run_control(torch_cuda, np.zeros((1, 1, 17, 4), np.float32))
    from snowmocap_b200 import _lib
    eng = _util_engine()
    t = torch_cuda.zeros((1, 1, 17, 4), device="cuda")
    c = torch_cuda.zeros((1, 1, 24, 4), device="cuda")
    v = torch_cuda.zeros((1, 1), dtype=torch_cuda.int32, device="cuda")
    rc = eng._lib.snowtri_blender_run(eng._h, t.data_ptr(), None, 1, 1, 17, c.data_ptr(), v.data_ptr(), None)
    assert rc == _lib.E_ARG


def _ref_like_result(persons):
    return {"hrnet_triangulate_points": [p for p in persons],
            "hrnet_triangulate_keypoint_scores": [np.ones(133) for _ in persons],
            "hrnet_triangulate_person_scores": [1.0] * len(persons)}


@pytest.mark.gpu
def test_dropin_blender_matches_reference_golden(torch_cuda):
    """Human_Triangulation_Blender with the shipped armature profile: same dicts as the reference."""
    from snowmocap_b200 import Human_Triangulation_Blender
    g = _golden("blender_points")
    names, _ = _profiles()
    profile = {n: [] for n in names}
    res = Human_Triangulation_Blender(_ref_like_result(g["pts"]), profile)
    cps, scs = res["blender_armature_control_points"], res["blender_armature_control_points_scores"]
    assert len(cps) == len(scs) == g["pts"].shape[0]
    for i, (cp, sc) in enumerate(zip(cps, scs)):
        assert list(cp.keys()) == names and list(sc.keys()) == names
        for k, n in enumerate(names):
            assert isinstance(cp[n], list) and len(cp[n]) == (4 if n == "root_rotation" else 3)
            assert sc[n] == int(g["valid"][i, k])
            if sc[n]:
                np.testing.assert_allclose(cp[n], g["ctrl"][i, k, :len(cp[n])], rtol=0, atol=1e-12)
            else:
                assert np.isnan(cp[n]).any()
    assert profile == {n: [] for n in names}, "the profile must not be modified"


@pytest.mark.gpu
def test_dropin_blender_error_behaviour(torch_cuda):
    from snowmocap_b200 import Human_Triangulation_Blender
    names, _ = _profiles()
    profile = {n: [] for n in names}
    rng = np.random.default_rng(9)
    p = _persons(rng, 1)[0].astype(np.float64)
    assert Human_Triangulation_Blender(_ref_like_result([]), profile) == {
        "blender_armature_control_points": [], "blender_armature_control_points_scores": []}
    with pytest.raises(IndexError):
        Human_Triangulation_Blender(_ref_like_result([p[:30]]), profile)
    with pytest.raises(NameError):
        Human_Triangulation_Blender(_ref_like_result([p]), {"tail_ik": []})
    q = p.copy()
    q[11] = q[12]
    with pytest.raises(np.linalg.LinAlgError):
        Human_Triangulation_Blender(_ref_like_result([q]), profile)
    sub = Human_Triangulation_Blender(_ref_like_result([p]), {"head_ik": [], "root_position": []})
    assert list(sub["blender_armature_control_points"][0].keys()) == ["head_ik", "root_position"]


@pytest.mark.gpu
def test_dropin_clip_matches_reference_golden(torch_cuda, tmp_path):
    """main.py:80-87 frame after frame: Blender -> Blender_Smooth -> To_Blender_Result, JSON written at the end."""
    from snowmocap_b200 import (Human_Triangulation_Blender, Human_Triangulation_Blender_Smooth,
                                Human_Triangulation_To_Blender_Result, save_blender_result)
    g = _golden("blender_smooth")
    names, fzr = _profiles()
    armature = {n: [] for n in names}
    smooth = {n: fzr[k].tolist() for k, n in enumerate(names)}
    prev, frames = None, []
    for t, n in enumerate(g["counts"]):
        cur = Human_Triangulation_Blender(_ref_like_result([g["clip"][t, k] for k in range(n)]), armature)
        sm = Human_Triangulation_Blender_Smooth(cur, armature, smooth, prev, delta_time=float(g["dt"]))
        prev = sm
        fin = Human_Triangulation_To_Blender_Result(sm)
        frames.append(fin)
        want, want_valid = g[f"ctrl_{t}"], g[f"valid_{t}"]
        assert len(fin["armature"]) == len(fin["score"]) == want.shape[0], t
        for i in range(want.shape[0]):
            for k, name in enumerate(names):
                got = np.asarray(fin["armature"][i][name])
                assert fin["score"][i][name] == int(want_valid[i, k])
                ref = want[i, k, :got.shape[0]]
                assert np.array_equal(np.isnan(got), np.isnan(ref)), (t, i, name)
                np.testing.assert_allclose(np.nan_to_num(got), np.nan_to_num(ref), rtol=0, atol=1e-11)
    path = tmp_path / "blender_mocap_data.json"
    save_blender_result(frames, str(path))
    back = json.loads(path.read_text())
    assert len(back) == len(frames) and set(back[0].keys()) == {"armature", "score"}


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, TOL_F32)])
@pytest.mark.parametrize("split", [None, 1, 7])
def test_batch_clip_matches_reference_golden(torch_cuda, dtype, tol, split):
    """Whole clip (or streamed batches) through snowtri_blender_run + snowtri_blender_smooth_run."""
    torch = torch_cuda
    from snowmocap_b200.blender import BlenderControl, BlenderSmoothState
    g = _golden("blender_smooth")
    _, fzr = _profiles()
    counts = g["counts"]
    F, P = g["clip"].shape[:2]
    buf = np.zeros((F, P, 133, 4), dtype)
    buf[..., :3] = g["clip"]
    eng = _util_engine()
    out = torch.from_numpy(buf).cuda()
    nout = torch.from_numpy(counts).cuda()
    ctrl, valid = BlenderControl(eng).run(out, nout)
    raw = ctrl.clone()
    state = BlenderSmoothState(eng, P, fzr)
    step = F if split is None else split
    nsm = []
    for s in range(0, F, step):
        nsm.append(state.run(ctrl[s:s + step], valid[s:s + step], nout[s:s + step], float(g["dt"])))
    torch.cuda.synchronize()
    nsm = torch.cat(nsm).cpu().numpy()
    ctrl, raw = ctrl.cpu().numpy(), raw.cpu().numpy()
    for t in range(F):
        want = g[f"ctrl_{t}"]
        m = want.shape[0]
        assert nsm[t] == m, t
        assert np.array_equal(np.isnan(ctrl[t, :m]), np.isnan(want)), t
        assert rel_l2(np.nan_to_num(ctrl[t, :m]), np.nan_to_num(want)) < tol, t
        assert np.array_equal(ctrl[t, m:], raw[t, m:], equal_nan=True), "rows beyond nsmooth stay untouched"
    state.reset()
    again = raw.copy()
    c2 = torch.from_numpy(again).cuda()
    state.run(c2, valid, nout, float(g["dt"]))
    torch.cuda.synchronize()
    assert np.array_equal(c2.cpu().numpy(), ctrl, equal_nan=True), "reset starts a new clip with identical results"
    state.close()


@pytest.mark.gpu
@pytest.mark.parametrize("chunked,batches", [(True, [600]), (False, [600]), (True, [290, 310]), (True, [1, 599]),
                                             (True, [300, 40, 260]), (True, [9000]), (True, [4500, 4500]),
                                             ("scan", [9000]), ("scan", [300, 40, 260]), ("scan", [4500, 4500]),
                                             ("fast", [9000]), ("fast", [300, 40, 260]), ("fast", [4500, 4500]),
                                             ("fast", [1, 599])])
def test_batch_smooth_long_clip_vs_oracle(torch_cuda, chunked, batches):
    """Person count changing, random invalid control points, non-zero r; the chunk-parallel paths (batches of more
    than 256 frames: one pass with a warm-up per chunk by default, "scan" = the chunk scan, 9000 frames = 71 chunks in
    3 groups), the sequential kernel, and a clip streamed through them."""
    torch = torch_cuda
    from snowmocap_b200.blender import BlenderSmoothState
    rng = np.random.default_rng(77)
    F, P = sum(batches), 3
    fzr = np.stack([rng.uniform(1.0, 3.0, 24), rng.uniform(0.5, 1.0, 24), rng.uniform(0.0, 0.5, 24)], axis=1)
    if chunked == "fast":
        # followers that all forget their state within 256 frames (like the shipped blender_smooth_profile.json, f = 1.5 ...
        # 3, z = 0.75): batches of more than 256 frames take the one-pass path; the slow followers above keep the scan
        fzr = np.stack([rng.uniform(1.5, 3.0, 24), rng.uniform(0.7, 0.8, 24), rng.uniform(0.0, 0.5, 24)], axis=1)
        chunked = True
    ctrl = np.cumsum(rng.normal(0, 0.01, (F, P, 24, 4)), axis=0) + rng.uniform(-2, 2, (1, P, 24, 4))
    ctrl[:, :, [k for k in range(24) if k != 1], 3] = 0.0
    vmask = rng.random((F, P, 24)) > 0.05
    ctrl[~vmask] = np.nan
    vbits = (vmask.astype(np.int64) << np.arange(24)).sum(-1).astype(np.int32)
    nout = rng.integers(1, P + 1, F).astype(np.int32)
    nout[0] = P
    o = bo.BlenderSmoothOracle(fzr)
    want = [o.step(ctrl[t, :nout[t]], vmask[t, :nout[t]], 1 / 30) for t in range(F)]
    eng = _util_engine()
    c = torch.from_numpy(ctrl.copy()).cuda()
    state = BlenderSmoothState(eng, P, fzr)
    state.set_chunked(chunked)
    vb, no = torch.from_numpy(vbits).cuda(), torch.from_numpy(nout).cuda()
    nsm, s0 = [], 0
    for n in batches:
        nsm.append(state.run(c[s0:s0 + n], vb[s0:s0 + n], no[s0:s0 + n], 1 / 30))
        s0 += n
    torch.cuda.synchronize()
    got = c.cpu().numpy()
    assert np.array_equal(torch.cat(nsm).cpu().numpy(), nout)
    for t in range(F):
        m = nout[t]
        assert np.array_equal(np.isnan(got[t, :m]), np.isnan(want[t])), t
        assert rel_l2(np.nan_to_num(got[t, :m]), np.nan_to_num(want[t])) < 1e-10, t
    state.close()


@pytest.mark.gpu
def test_pipeline_run_to_control_points(torch_cuda):
    """snowtri_run (cfg2 rig, float32 output) -> snowtri_blender_run on the same device buffers, against the
    oracle applied to the kernel's own float32 joints."""
    torch = torch_cuda
    from conftest import floor_rig
    from snowmocap_b200 import synth
    from snowmocap_b200.blender import BlenderControl
    from snowmocap_b200.engine import TriangulationEngine
    rig = floor_rig()
    data = synth.make_frames(rig, 96, 1, 133, seed=5)
    eng = TriangulationEngine(rig.K, rig.R, rig.t, device=0, precision="f64", **synth.DEFAULT_PARAMS)
    kp, sc, cn = [torch.from_numpy(np.ascontiguousarray(data[k])).cuda() for k in ("kpts", "scores", "counts")]
    res = eng.run(kp, sc, cn, Pout=1)
    ctrl, vbits = BlenderControl(eng).run(res["out"], res["nout"])
    torch.cuda.synchronize()
    out = res["out"].cpu().numpy()
    assert (res["nout"].cpu().numpy() == 1).all()
    want, want_valid = bo.control_points_batch(out.reshape(-1, 133, 4)[..., :3].astype(np.float64))
    valid = _masks(vbits.cpu().numpy().reshape(-1))
    assert np.array_equal(valid, want_valid)
    got = ctrl.cpu().numpy().reshape(-1, 24, 4)
    assert rel_l2(got[want_valid], want[want_valid]) < TOL_F32
    eng.close()


# ---- the whole per-frame body of the reference's main.py (:50-87) ----------------------------------------------
def _pipeline_golden():
    g = _golden("pipeline_main")
    return g, json.loads(str(g["params"]))


@pytest.mark.gpu
def test_main_py_sequence_through_dropin_names(torch_cuda):
    """main.py:50-87 with only the import changed: add_human_2D_points -> Human_Triangulation -> Condense -> Smooth
    -> Blender -> Blender_Smooth -> To_Blender_Result, frame after frame, against the real reference's output."""
    import snowmocap_b200 as sv
    from conftest import floor_rig
    g, p = _pipeline_golden()
    names, fzr = _profiles()
    armature = {n: [] for n in names}
    smooth = {n: fzr[k].tolist() for k, n in enumerate(names)}
    rig = floor_rig()
    group = sv.CameraGroup(cap_ids=list(range(rig.C)), resolutions=[(1280, 720)] * rig.C)
    for c in range(rig.C):
        group.cameras[c].K, group.cameras[c].R, group.cameras[c].t = rig.K[c], rig.R[c], rig.t[c].reshape(3, 1)
    prev_tri, prev_bl = None, None
    F = g["kpts"].shape[0]
    for f in range(F):
        for c in range(rig.C):
            group.add_human_2D_points(g["kpts"][f, c, 0], g["scores"][f, c, 0], c)
        tri = sv.Human_Triangulation(group, keypoint_score_threshold=p["kst"], average_score_threshold=p["ast"],
                                     distance_threshold=p["dthr"])
        tri = sv.Human_Triangulation_Condense(tri, condense_distance_tol=p["cond_tol"],
                                              condense_person_num_tol=p["num_tol"], condense_score_tol=p["score_tol"],
                                              center_point_index=p["center"], keypoint_num=p["keypoint_num"])
        tri = sv.Human_Triangulation_Smooth(tri, prev_tri, f=p["smooth_f"], z=p["smooth_z"], r=p["smooth_r"],
                                            delta_time=p["smooth_delta_time"])
        prev_tri = tri
        bl = sv.Human_Triangulation_Blender(tri, armature)
        bl = sv.Human_Triangulation_Blender_Smooth(bl, armature, smooth, prev_bl, delta_time=p["smooth_delta_time"])
        prev_bl = bl
        fin = sv.Human_Triangulation_To_Blender_Result(bl)
        group.clear_2D_points()
        want, want_valid = g[f"ctrl_{f}"], g[f"valid_{f}"]
        assert rel_l2(np.array(tri["hrnet_triangulate_points"]).reshape(-1, 133, 3), g[f"joints_{f}"]) < 1e-9
        assert len(fin["armature"]) == len(fin["score"]) == want.shape[0] == 1
        got = np.zeros((24, 4))
        for k, name in enumerate(names):
            v = np.asarray(fin["armature"][0][name])
            got[k, :v.shape[0]] = v
            assert fin["score"][0][name] == int(want_valid[0, k]) == 1
        assert rel_l2(got, want[0]) < 1e-9, f


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("f64", 1e-6), ("f32", 1e-4)])
def test_main_py_sequence_device_resident_batch(torch_cuda, precision, tol):
    """The same clip as one batch that never leaves the device: snowtri_run -> snowtri_smooth_run ->
    snowtri_blender_run -> snowtri_blender_smooth_run, against the real reference's final control points
    (north_star bound 1e-4 in float32 arithmetic; float64 arithmetic with float32 storage 1e-6)."""
    torch = torch_cuda
    from conftest import floor_rig
    from snowmocap_b200.blender import BlenderControl, BlenderSmoothState
    from snowmocap_b200.engine import SmoothState, TriangulationEngine
    g, p = _pipeline_golden()
    _, fzr = _profiles()
    rig = floor_rig()
    eng = TriangulationEngine(rig.K, rig.R, rig.t, device=0, precision=precision, kst=p["kst"], ast=p["ast"],
                              dthr=p["dthr"], cond_tol=p["cond_tol"], num_tol=p["num_tol"], score_tol=p["score_tol"],
                              center=p["center"])
    kp, sc = torch.from_numpy(g["kpts"]).cuda(), torch.from_numpy(g["scores"]).cuda()
    res = eng.run(kp, sc, None, Pout=1, keypoint_num=p["keypoint_num"])
    out, nout = res["out"], res["nout"]
    sm = SmoothState(eng, 1, 133, p["smooth_f"], p["smooth_z"], p["smooth_r"])
    nsm = sm.run(out, nout, p["smooth_delta_time"])
    ctrl, valid = BlenderControl(eng).run(out, nsm)
    bs = BlenderSmoothState(eng, 1, fzr)
    nfin = bs.run(ctrl, valid, nsm, p["smooth_delta_time"])
    torch.cuda.synchronize()
    F = g["kpts"].shape[0]
    assert (nfin.cpu().numpy() == 1).all()
    got = ctrl.cpu().numpy()[:, 0]
    want = np.stack([g[f"ctrl_{f}"][0] for f in range(F)])
    joints = np.stack([g[f"joints_{f}"][0] for f in range(F)])
    assert rel_l2(out.cpu().numpy()[:, 0, :, :3], joints) < tol
    assert rel_l2(got, want) < tol
    assert (valid.cpu().numpy() == (1 << 24) - 1).all()
    sm.close(); bs.close(); eng.close()


@pytest.mark.gpu
def test_control_points_full_size_properties(torch_cuda):
    """BASELINE cfg2 size (131 072 frames x 1 person): translation equivariance (positions move with the person,
    the root quaternion does not change at all), determinism, and a random sample of rows against the oracle."""
    torch = torch_cuda
    from snowmocap_b200.blender import BlenderControl
    eng = _util_engine()
    F = 131072
    g = torch.Generator(device="cuda").manual_seed(11)
    # coordinates on a 1/1024 grid in [0, 4): adding an integer offset is exact in float32
    pts = torch.randint(0, 4096, (F, 1, 133, 4), generator=g, device="cuda").to(torch.float32) / 1024.0
    bc = BlenderControl(eng)
    ctrl, valid = bc.run(pts, None)
    ctrl2, valid2 = bc.run(pts, None)
    assert torch.equal(ctrl, ctrl2) and torch.equal(valid, valid2)
    off = torch.tensor([3.0, -2.0, 1.0, 0.0], device="cuda")
    ctrl_s, valid_s = bc.run(pts + off, None)
    torch.cuda.synchronize()
    assert torch.equal(valid, valid_s)
    full = (valid == (1 << 24) - 1).view(-1)
    assert full.float().mean() > 0.99
    q, q_s = ctrl[:, 0, 1][full], ctrl_s[:, 0, 1][full]
    assert torch.equal(q, q_s), "differences of exactly shifted inputs are identical"
    assert torch.allclose(q.double().norm(dim=1), torch.ones(1, dtype=torch.float64, device="cuda"), atol=1e-6)
    pos = [k for k in range(24) if k != 1]
    d = (ctrl_s[:, 0, pos] - ctrl[:, 0, pos])[full]
    assert (d - off).abs().max().item() < 2e-6
    rows = np.random.default_rng(0).choice(F, 200, replace=False)
    want, want_valid = bo.control_points_batch(pts[rows, 0, :, :3].cpu().numpy().astype(np.float64))
    got = ctrl[rows, 0].cpu().numpy()
    assert np.array_equal(_masks(valid[rows, 0].cpu().numpy()), want_valid)
    assert rel_l2(got[want_valid], want[want_valid]) < TOL_F32
