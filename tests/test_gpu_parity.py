"""Parity of the CUDA path against the golden vectors and the CPU oracles (run on the B200 box).

Tolerances: the candidate/condense entry points compute and return float64 -> 1e-9 relative L2
against the reference.  The fused path computes in float64 but stores float32 -> 1e-6.  The
north_star bound for the fused path in any precision mode is 1e-4 relative L2 on the 3D points.
"""
import numpy as np
import pytest

from conftest import Golden, floor_rig, golden_names, rel_l2
from snowmocap_b200 import synth

pytestmark = pytest.mark.gpu

TOL_F64 = 1e-9
TOL_FUSED = 1e-6
TOL_NORTH_STAR = 1e-4


F32_DIST_ERR = 1e-5   # metres: bound of the float32 error of a ray-to-ray distance (kDistDelta in the kernels)


def score_error_bound(kscore, cameras):
    """Relative error bound of an all-float32 keypoint score.  kscore = sum over the cluster's pairs of
    (sm+ss)/2 / (1000 dist_e), divided by the member count n <= C(C-1)/2; detector scores that count are >= 0.5, so the
    closest pair has 1/dist <= 1000 n kscore / 0.5 and its term -- hence the sum -- is off by at most
    F32_DIST_ERR / dist relatively.  The additive 1e-3 covers pairs at ordinary (millimetre) distances."""
    n = cameras * (cameras - 1) // 2
    return F32_DIST_ERR * 2000.0 * n * np.asarray(kscore) + 1e-3


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    return torch


def _engine(g_or_rig, params, precision="f64"):
    from snowmocap_b200.engine import TriangulationEngine
    p = {k: v for k, v in params.items() if k != "keypoint_num"}
    return TriangulationEngine(g_or_rig.K, g_or_rig.R, g_or_rig.t, device=0, precision=precision, **p)


def _to_dev(torch, *arrays):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]


@pytest.mark.parametrize("name", golden_names())
def test_fused_matches_reference_golden(torch_cuda, name):
    torch = torch_cuda
    g = Golden(name)
    eng = _engine(g, g.params)
    pout = max(1, max(c[0].shape[0] for c in g.con))
    kp, sc, cn = _to_dev(torch, g.kpts, g.scores, g.counts)
    res = eng.run(kp, sc, cn, Pout=pout, keypoint_num=g.params["keypoint_num"])
    torch.cuda.synchronize()
    out, ps, nout = res["out"].cpu().numpy(), res["pscores"].cpu().numpy(), res["nout"].cpu().numpy()
    for f in range(g.F):
        pts, ks, pscore = g.con[f]
        n = pts.shape[0]
        assert nout[f] == n, f"{name} frame {f}: persons {nout[f]} != {n}"
        if n:
            assert rel_l2(out[f, :n, :, :3], pts) < TOL_FUSED
            np.testing.assert_allclose(out[f, :n, :, 3], ks, rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(ps[f, :n], pscore, rtol=1e-5, atol=1e-7)
        assert not out[f, n:].any() and not ps[f, n:].any(), "unused slots must be zero"


@pytest.mark.parametrize("name", golden_names())
def test_candidates_match_reference_golden(torch_cuda, name):
    torch = torch_cuda
    g = Golden(name)
    eng = _engine(g, g.params)
    kp, sc, cn = _to_dev(torch, g.kpts, g.scores, g.counts)
    res = eng.candidates(kp, sc, cn)
    torch.cuda.synchronize()
    cand, avg, keep = res["cand"].cpu().numpy(), res["avg"].cpu().numpy(), res["keep"].cpu().numpy().astype(bool)
    for f in range(g.F):
        pts, ks, pscore = g.tri[f]
        assert keep[f].sum() == pts.shape[0]
        if pts.shape[0]:
            assert rel_l2(cand[f][keep[f]][:, :, :3], pts) < TOL_F64
            np.testing.assert_allclose(cand[f][keep[f]][:, :, 3], ks, rtol=1e-8, atol=1e-12)
            np.testing.assert_allclose(avg[f][keep[f]], pscore, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("name", golden_names())
def test_condense_matches_reference_golden(torch_cuda, name):
    torch = torch_cuda
    g = Golden(name)
    eng = _engine(g, g.params)
    nmax = max(1, max(t[0].shape[0] for t in g.tri))
    J = g.kpts.shape[3]
    cand = np.zeros((g.F, nmax, J, 4))
    ncand = np.zeros(g.F, np.int32)
    for f in range(g.F):
        n = g.tri[f][0].shape[0]
        ncand[f] = n
        cand[f, :n, :, :3] = g.tri[f][0]
        cand[f, :n, :, 3] = g.tri[f][1]
    dc, dn = _to_dev(torch, cand, ncand)
    res = eng.condense(dc, dn, keypoint_num=g.params["keypoint_num"], Pout=nmax)
    torch.cuda.synchronize()
    out, ps, nout = res["out"].cpu().numpy(), res["pscores"].cpu().numpy(), res["nout"].cpu().numpy()
    for f in range(g.F):
        pts, ks, pscore = g.con[f]
        n = pts.shape[0]
        assert nout[f] == n
        if n:
            assert rel_l2(out[f, :n, :, :3], pts) < TOL_F64
            np.testing.assert_allclose(out[f, :n, :, 3], ks, rtol=1e-9, atol=1e-13)
            np.testing.assert_allclose(ps[f, :n], pscore, rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("name", golden_names())
def test_dropin_api_matches_reference_golden(torch_cuda, name):
    """The reference's own call sequence (main.py:55-71, 106) through the drop-in names."""
    import snowmocap_b200 as sv
    g, p = Golden(name), Golden(name).params
    C = g.K.shape[0]
    group = sv.CameraGroup(cap_ids=list(range(C)), resolutions=[(1280, 720)] * C)
    for c in range(C):
        group.cameras[c].K, group.cameras[c].R, group.cameras[c].t = g.K[c], g.R[c], g.t[c].reshape(3, 1)
    for f in range(g.F):
        for c in range(C):
            for q in range(int(g.counts[f, c])):
                group.add_human_2D_points(g.kpts[f, c, q], g.scores[f, c, q], c)
        tri = sv.Human_Triangulation(group, keypoint_score_threshold=p["kst"], average_score_threshold=p["ast"],
                                     distance_threshold=p["dthr"])
        con = sv.Human_Triangulation_Condense(tri, condense_distance_tol=p["cond_tol"],
                                              condense_person_num_tol=p["num_tol"], condense_score_tol=p["score_tol"],
                                              center_point_index=p["center"], keypoint_num=p["keypoint_num"])
        group.clear_2D_points()
        for got, want in ((tri, g.tri[f]), (con, g.con[f])):
            assert set(got) == {"hrnet_triangulate_points", "hrnet_triangulate_keypoint_scores",
                                "hrnet_triangulate_person_scores"}
            n = want[0].shape[0]
            assert len(got["hrnet_triangulate_points"]) == n
            if n:
                assert got["hrnet_triangulate_points"][0].dtype == np.float64
                assert rel_l2(np.array(got["hrnet_triangulate_points"]), want[0]) < TOL_F64
                np.testing.assert_allclose(np.array(got["hrnet_triangulate_keypoint_scores"]), want[1], rtol=1e-8, atol=1e-12)
                np.testing.assert_allclose(np.array(got["hrnet_triangulate_person_scores"]), want[2], rtol=1e-8, atol=1e-12)


CASES = [
    # rig, F, P, J, params, Pout, data kwargs
    ("floor2", 300, 1, 17, synth.DEFAULT_PARAMS, 2, {}),                       # BASELINE configs[0]
    ("floor4", 1500, 1, 133, synth.DEFAULT_PARAMS, 2, {}),                     # BASELINE configs[1]
    ("ring8", 120, 4, 133, synth.MULTI_PARAMS, 8, {}),                         # BASELINE configs[2] geometry
    ("ring8", 60, 4, 133, synth.MULTI_PARAMS, 8, {"low_score_frac": 0.1, "drop_prob": 0.15}),
    ("ring6", 200, 3, 17, dict(synth.MULTI_PARAMS, ast=0.0), 64, {"low_score_frac": 0.2}),   # ghost clusters
    ("ring6", 200, 3, 17, dict(synth.MULTI_PARAMS, score_tol=0.3, num_tol=2, center=4), 8, {"low_score_frac": 0.1}),
    ("ring5", 150, 2, 33, dict(synth.MULTI_PARAMS, kst=-1.0, ast=-5.0), 16, {}),              # kst<0: general paths
]


def _rig(name):
    if name.startswith("floor"):
        return floor_rig().subset(int(name[5:]))
    return synth.ring_rig(int(name[4:]))


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("tuning", [(0, 0, 0), (1, 3, 256), (5, 7, 512), (3, 0, -256), (8, 0, 0)])
def test_fused_matches_c_oracle(torch_cuda, case, tuning):
    """Seeded synthetic batches vs the C oracle, with different frames-per-group / CTA counts
    (exercises the TMA and the plain-load staging paths, group tails and the persistent loop)."""
    torch = torch_cuda
    from oracle import c_oracle
    rname, F, P, J, prm, pout, kw = CASES[case]
    rig = _rig(rname)
    d = synth.make_frames(rig, F, P, J, seed=100 + case, **kw)
    ref = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, prm, Pout=pout)
    eng = _engine(rig, prm)
    eng.set_tuning(*tuning)
    kp, sc, cn = _to_dev(torch, d["kpts"], d["scores"], d["counts"])
    res = eng.run(kp, sc, cn, Pout=pout)
    torch.cuda.synchronize()
    out, ps, nout = res["out"].cpu().numpy(), res["pscores"].cpu().numpy(), res["nout"].cpu().numpy()
    assert np.array_equal(nout, ref["nout"])
    m = np.minimum(ref["nout"], pout)
    valid = np.arange(pout)[None, :] < m[:, None]
    assert rel_l2(out[valid][:, :, :3], ref["points"][valid]) < TOL_FUSED
    np.testing.assert_allclose(out[valid][:, :, 3], ref["kscores"][valid], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(ps[valid], ref["pscores"][valid], rtol=2e-5, atol=1e-7)
    assert not out[~valid].any()


# Single-person kernel (snowtri_p1.cuh): every camera count it is compiled for, absent cameras,
# low scores, clusters that split (small condense_distance_tol), num_tol rejections, truncated
# keypoint_num, tiles of 1..32 frames, more clusters than output slots.
P1_CASES = [
    # rig, F, J, params, Pout, keypoint_num, data kwargs
    ("ring2", 97, 17, synth.DEFAULT_PARAMS, 1, None, {}),
    ("ring3", 130, 17, dict(synth.DEFAULT_PARAMS, cond_tol=0.004), 3, None, {"low_score_frac": 0.1}),
    ("floor4", 700, 133, synth.DEFAULT_PARAMS, 1, None, {"low_score_frac": 0.1, "drop_prob": 0.2}),
    ("floor4", 300, 133, dict(synth.DEFAULT_PARAMS, cond_tol=0.003, center=5), 4, 100, {"drop_prob": 0.1}),
    ("ring5", 260, 33, dict(synth.DEFAULT_PARAMS, cond_tol=0.004, num_tol=2), 2, None, {"low_score_frac": 0.05}),
    ("ring6", 150, 133, dict(synth.DEFAULT_PARAMS, cond_tol=0.005), 6, None, {"drop_prob": 0.3}),
    ("ring7", 90, 40, dict(synth.DEFAULT_PARAMS, dthr=0.004), 2, None, {"low_score_frac": 0.1}),
    ("ring8", 200, 133, dict(synth.DEFAULT_PARAMS, cond_tol=0.004), 5, None, {"low_score_frac": 0.1, "drop_prob": 0.15}),
]


@pytest.mark.parametrize("case", range(len(P1_CASES)))
@pytest.mark.parametrize("precision,tile", [("f64", 0), ("f64", 1), ("f64", 5), ("f64", 32), ("f32", 0), ("f32", 3),
                                            ("mixed", 0), ("mixed", 7)])
def test_single_person_kernel_matches_c_oracle(torch_cuda, case, precision, tile):
    torch = torch_cuda
    from oracle import c_oracle
    rname, F, J, prm, pout, knum, kw = P1_CASES[case]
    rig = _rig(rname)
    d = synth.make_frames(rig, F, 1, J, seed=300 + case, **kw)
    jout = knum or J
    ref = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, prm, Pout=pout, keypoint_num=jout)
    eng = _engine(rig, prm, precision=precision)
    eng.set_tuning(tile, 0, 0)
    kp, sc, cn = _to_dev(torch, d["kpts"], d["scores"], d["counts"])
    res = eng.run(kp, sc, cn, Pout=pout, keypoint_num=jout)
    torch.cuda.synchronize()
    assert eng.last_launch_info()["kernel"] == "p1"
    out, ps, nout = res["out"].cpu().numpy(), res["pscores"].cpu().numpy(), res["nout"].cpu().numpy()
    assert np.array_equal(nout, ref["nout"])
    m = np.minimum(ref["nout"], pout)
    valid = np.arange(pout)[None, :] < m[:, None]
    assert not out[~valid].any() and not ps[~valid].any()
    if valid.any():
        # zero / non-zero pattern of the keypoint scores is a set of discrete decisions: exact in both modes
        assert np.array_equal(out[valid][:, :, 3] == 0, ref["kscores"][valid] == 0)
        if precision == "f64":
            assert rel_l2(out[valid][:, :, :3], ref["points"][valid]) < TOL_FUSED
            np.testing.assert_allclose(out[valid][:, :, 3], ref["kscores"][valid], rtol=2e-5, atol=1e-7)
            np.testing.assert_allclose(ps[valid], ref["pscores"][valid], rtol=2e-5, atol=1e-7)
        elif precision == "mixed":
            # float32 bulk, float64 distance numerator: scores are as good as float32 storage allows
            assert rel_l2(out[valid][:, :, :3], ref["points"][valid]) < TOL_FUSED * 5
            np.testing.assert_allclose(out[valid][:, :, 3], ref["kscores"][valid], rtol=1e-4, atol=1e-7)
            np.testing.assert_allclose(ps[valid], ref["pscores"][valid], rtol=1e-4, atol=1e-7)
        else:
            # all-float32: the points hold the north_star bound with margin.  A keypoint score is 1/distance of two
            # nearly intersecting rays, ill-conditioned in float32: the float32 ray distance is good
            # to F32_DIST_ERR metres, so a pair's score s = (sm+ss)/2 / (1000 dist) is off by at most
            # s * F32_DIST_ERR / dist relatively -- every score is bounded by score_error_bound(), the typical one
            # (median, 99th percentile) much tighter
            assert rel_l2(out[valid][:, :, :3], ref["points"][valid]) < TOL_NORTH_STAR / 10
            ks, kr = out[valid][:, :, 3].astype(np.float64), ref["kscores"][valid]
            nz = kr != 0
            rel = np.abs(ks[nz] - kr[nz]) / kr[nz]
            assert np.median(rel) < 1e-3 and np.quantile(rel, 0.99) < 0.25
            assert (rel <= score_error_bound(kr[nz], rig.C)).all(), float((rel / score_error_bound(kr[nz], rig.C)).max())


def test_single_person_kernel_equals_general_kernel(torch_cuda):
    """Same batch through p1_kernel and through fused_kernel (explicit block size selects the latter)."""
    torch = torch_cuda
    rig = floor_rig()
    d = synth.make_frames(rig, 333, 1, 133, seed=5, low_score_frac=0.1, drop_prob=0.1)
    kp, sc, cn = _to_dev(torch, d["kpts"], d["scores"], d["counts"])
    outs = []
    for threads in (0, 256):
        eng = _engine(rig, synth.DEFAULT_PARAMS)
        eng.set_tuning(0, 0, threads)
        res = eng.run(kp, sc, cn, Pout=2)
        torch.cuda.synchronize()
        assert eng.last_launch_info()["kernel"].startswith("p1" if threads == 0 else "fused")
        outs.append([res[k].cpu().numpy() for k in ("out", "pscores", "nout")])
    assert np.array_equal(outs[0][2], outs[1][2])
    assert rel_l2(outs[0][0], outs[1][0]) < TOL_FUSED
    np.testing.assert_allclose(outs[0][1], outs[1][1], rtol=2e-5, atol=1e-7)


def test_fused_without_counts_and_host_path(torch_cuda):
    torch = torch_cuda
    from oracle import c_oracle
    rig = synth.ring_rig(4)
    d = synth.make_frames(rig, 257, 2, 133, seed=3)
    prm = synth.MULTI_PARAMS
    ref = c_oracle.fused(d["kpts"], d["scores"], None, rig.K, rig.R, rig.t, prm, Pout=4)
    eng = _engine(rig, prm)
    res = eng.run_host(d["kpts"], d["scores"], None, Pout=4)
    assert np.array_equal(res["nout"], ref["nout"])
    assert rel_l2(res["out"][..., :3], ref["points"]) < TOL_FUSED
    assert eng.last_launch_info()["kernel"] == "general" and eng.launch_count == 6   # one chunk: keep, centre, cluster, members, fuse, pscore


@pytest.mark.parametrize("chunk", [0, 1, 37, 100000])
def test_host_pipeline_chunks_match_device_path(torch_cuda, chunk):
    """snowtri_run_host cuts the batch into chunks over two streams: same bytes as one device-side run."""
    torch = torch_cuda
    rig = floor_rig()
    d = synth.make_frames(rig, 301, 1, 133, seed=11, low_score_frac=0.05, drop_prob=0.1)
    eng = _engine(rig, synth.DEFAULT_PARAMS, precision="f32")
    kp, sc, cn = _to_dev(torch, d["kpts"], d["scores"], d["counts"])
    ref = eng.run(kp, sc, cn, Pout=2)
    torch.cuda.synchronize()
    eng.set_pipeline(chunk)
    n0 = eng.launch_count
    res = eng.run_host(d["kpts"], d["scores"], d["counts"], Pout=2)
    assert eng.launch_count - n0 == (1 if chunk in (0, 100000) else -(-301 // chunk))
    for k in ("out", "nout"):
        assert np.array_equal(res[k], ref[k].cpu().numpy()), k
    # the person score is a float32 sum whose lane partition depends on the frame's position in its tile
    np.testing.assert_allclose(res["pscores"], ref["pscores"].cpu().numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("case", [0, 1, 2, 3])
@pytest.mark.parametrize("precision", ["f32", "mixed", "f32x"])
def test_fused_float_modes_within_north_star(torch_cuda, case, precision):
    """Float modes of snowtri_run on the general cases: persons emitted must be identical (every discrete
    decision is float64-guarded) and the joints within the north_star bound.  "f32x" also runs the general
    kernel in float32 with several persons, where ghost clusters are only good to ~1e-3."""
    torch = torch_cuda
    from oracle import c_oracle
    rname, F, P, J, prm, pout, kw = CASES[case]
    rig = _rig(rname)
    d = synth.make_frames(rig, F, P, J, seed=200 + case, **kw)
    ref = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, prm, Pout=pout)
    eng = _engine(rig, prm, precision=precision)
    kp, sc, cn = _to_dev(torch, d["kpts"], d["scores"], d["counts"])
    res = eng.run(kp, sc, cn, Pout=pout)
    torch.cuda.synchronize()
    out, nout = res["out"].cpu().numpy(), res["nout"].cpu().numpy()
    assert np.array_equal(nout, ref["nout"])
    m = np.minimum(ref["nout"], pout)
    valid = np.arange(pout)[None, :] < m[:, None]
    tol = 2e-3 if (precision == "f32x" and P > 1) else TOL_NORTH_STAR
    assert rel_l2(out[valid][:, :, :3], ref["points"][valid]) < tol


def test_fused_equals_candidates_then_condense(torch_cuda):
    """K1 -> K2 composition must reproduce the fused kernel."""
    torch = torch_cuda
    rig = synth.ring_rig(5)
    d = synth.make_frames(rig, 40, 3, 33, seed=9, low_score_frac=0.1, drop_prob=0.1)
    prm = synth.MULTI_PARAMS
    eng = _engine(rig, prm)
    kp, sc, cn = _to_dev(torch, d["kpts"], d["scores"], d["counts"])
    fused = eng.run(kp, sc, cn, Pout=8)
    c = eng.candidates(kp, sc, cn)
    torch.cuda.synchronize()
    keep = c["keep"].cpu().numpy().astype(bool)
    cand = c["cand"].cpu().numpy()
    F, N = keep.shape
    packed = np.zeros_like(cand)
    ncand = keep.sum(1).astype(np.int32)
    for f in range(F):
        packed[f, :ncand[f]] = cand[f][keep[f]]
    dc, dn = _to_dev(torch, packed, ncand)
    con = eng.condense(dc, dn, Pout=8)
    torch.cuda.synchronize()
    assert np.array_equal(con["nout"].cpu().numpy(), fused["nout"].cpu().numpy())
    assert rel_l2(fused["out"].cpu().numpy(), con["out"].cpu().numpy()) < TOL_FUSED


def test_skew_ray_solver(torch_cuda):
    import snowmocap_b200 as sv
    from oracle import c_oracle
    dist, W = sv.Skew_Ray_Solver(np.array([[1.0], [0], [0]]), np.array([[0], [1.0], [0]]),
                                 np.array([[-1.0], [0], [0]]), np.array([[0], [-1.0], [2.0]]))
    assert abs(dist - 2.0) < 1e-13 and np.allclose(W, [0, 0, 1.0], atol=1e-13)
    rng = np.random.default_rng(0)
    hm, hs, tm, ts = (rng.standard_normal((1000, 3)) for _ in range(4))
    eng = sv.triangulation._util_engine()
    torch = torch_cuda
    d, m = eng.skew_ray(*_to_dev(torch, hm, hs, tm, ts))
    d0, m0 = c_oracle.skew_ray(hm, hs, tm, ts)
    np.testing.assert_allclose(d.cpu().numpy(), d0, rtol=1e-9)
    np.testing.assert_allclose(m.cpu().numpy(), m0, rtol=1e-9, atol=1e-9)


def test_error_behaviour(torch_cuda):
    import snowmocap_b200 as sv
    from snowmocap_b200._lib import SnowtriError
    torch = torch_cuda
    rig = synth.ring_rig(3)
    d = synth.make_frames(rig, 2, 1, 17, seed=1)
    eng = _engine(rig, dict(synth.DEFAULT_PARAMS, center=40))
    kp, sc, cn = _to_dev(torch, d["kpts"], d["scores"], d["counts"])
    with pytest.raises(SnowtriError):
        eng.run(kp, sc, cn)                      # centre joint >= J: the reference raises IndexError
    eng.set_params(center=0)
    with pytest.raises(SnowtriError):
        eng.run(kp, sc, cn, keypoint_num=18)     # keypoint_num > J
    tri = {"hrnet_triangulate_points": [np.zeros((17, 3))] * 3,
           "hrnet_triangulate_keypoint_scores": [np.ones(17)] * 3, "hrnet_triangulate_person_scores": [1.0] * 3}
    with pytest.raises(IndexError):
        sv.Human_Triangulation_Condense(tri, keypoint_num=30)


# ---- temporal smoothing -----------------------------------------------------------------
from conftest import GOLDEN, smooth_golden_names  # noqa: E402


def _smooth_golden(name):
    import json
    import os
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    counts = z["counts"]
    return z["pts"], counts, json.loads(str(z["params"])), [z[f"out_{t}"] for t in range(len(counts))]


@pytest.mark.parametrize("name", smooth_golden_names())
def test_smooth_dropin_matches_reference_golden(torch_cuda, name):
    """main.py:72-78 call pattern through the drop-in name, frame after frame."""
    import snowmocap_b200 as sv
    pts, counts, prm, want = _smooth_golden(name)
    J = pts.shape[2]
    prev = None
    for t in range(len(counts)):
        n = int(counts[t])
        res = {"hrnet_triangulate_points": [pts[t, k].copy() for k in range(n)],
               "hrnet_triangulate_keypoint_scores": [np.ones(J) for _ in range(n)],
               "hrnet_triangulate_person_scores": [1.0] * n}
        res = sv.Human_Triangulation_Smooth(res, prev, f=prm["f"], z=prm["z"], r=prm["r"], delta_time=prm["dt"])
        prev = res
        got = res["hrnet_triangulate_points"]
        assert set(res) == {"hrnet_triangulate_points", "hrnet_triangulate_keypoint_scores",
                            "hrnet_triangulate_person_scores", "second_order_dynamics"}
        assert len(got) == want[t].shape[0], f"frame {t}"
        assert len(res["hrnet_triangulate_keypoint_scores"]) == n      # scores pass through unaligned
        if want[t].size:
            np.testing.assert_allclose(np.array(got), want[t], rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("name", smooth_golden_names())
@pytest.mark.parametrize("split", [None, 1, 6])
def test_smooth_batch_matches_reference_golden(torch_cuda, name, split):
    """Whole clip (or streamed chunks) through snowtri_smooth_run_f64."""
    torch = torch_cuda
    from snowmocap_b200.engine import SmoothState
    import snowmocap_b200 as sv
    pts, counts, prm, want = _smooth_golden(name)
    F, P, J, _ = pts.shape
    eng = sv.triangulation._util_engine()
    st = SmoothState(eng, P, J, prm["f"], prm["z"], prm["r"])
    dense = np.zeros((F, P, J, 4))
    dense[..., :3] = pts
    dense[..., 3] = 0.25
    step = F if split is None else split
    for t0 in range(0, F, step):
        out = torch.from_numpy(dense[t0:t0 + step].copy()).cuda()
        nsm = st.run(out, torch.from_numpy(counts[t0:t0 + step].copy()).cuda(), prm["dt"]).cpu().numpy()
        o = out.cpu().numpy()
        for i, t in enumerate(range(t0, min(F, t0 + step))):
            assert nsm[i] == want[t].shape[0], f"frame {t}"
            if want[t].size:
                np.testing.assert_allclose(o[i, :nsm[i], :, :3], want[t], rtol=1e-11, atol=1e-13)
            assert np.all(o[i, :, :, 3] == 0.25)                                  # score untouched
            np.testing.assert_array_equal(o[i, nsm[i]:, :, :3], pts[t, nsm[i]:])   # dropped persons untouched


def test_smooth_float32_layout_long_clip_vs_c_oracle(torch_cuda):
    """snowtri_run layout (float32 x,y,z,score) over a long clip with a ragged person count."""
    torch = torch_cuda
    from oracle import c_oracle
    from snowmocap_b200.engine import SmoothState
    import snowmocap_b200 as sv
    rng = np.random.default_rng(5)
    F, P, J = 5000, 3, 133
    pts = (rng.uniform(-2, 2, (1, P, 1, 3)) + np.cumsum(rng.normal(0, 0.01, (F, P, J, 3)), axis=0)).astype(np.float32)
    nout = rng.integers(0, P + 1, F).astype(np.int32)
    nout[0] = 2
    ref, nsm_ref, _ = c_oracle.smooth(pts.astype(np.float64), nout, 2.5, 0.75, 0.3, 1 / 30)
    dense = np.zeros((F, P, J, 4), np.float32)
    dense[..., :3] = pts
    st = SmoothState(sv.triangulation._util_engine(), P, J, 2.5, 0.75, 0.3)
    out = torch.from_numpy(dense).cuda()
    nsm = st.run(out, torch.from_numpy(nout).cuda(), 1 / 30).cpu().numpy()
    assert np.array_equal(nsm, nsm_ref)
    o = out.cpu().numpy()
    m = np.arange(P)[None, :] < nsm[:, None]
    assert rel_l2(o[m][..., :3], ref[m]) < 1e-6          # float32 storage of the result
    st.reset()
    out2 = torch.from_numpy(dense).cuda()
    nsm2 = st.run(out2, torch.from_numpy(nout).cuda(), 1 / 30).cpu().numpy()
    assert np.array_equal(nsm2, nsm) and torch.equal(out2, out)   # reset starts the same clip again


@pytest.mark.parametrize("chunked", [True, False, "scan"])
@pytest.mark.parametrize("batches", [(3000,), (1, 700, 299, 2000), (300, 2700)])
def test_smooth_long_clip_float64_vs_c_oracle(torch_cuda, chunked, batches):
    """Chunk-parallel and sequential kernels, whole clip or streamed batches, ragged person count."""
    torch = torch_cuda
    from oracle import c_oracle
    from snowmocap_b200.engine import SmoothState
    import snowmocap_b200 as sv
    rng = np.random.default_rng(8)
    F, P, J = sum(batches), 3, 17
    pts = rng.uniform(-2, 2, (1, P, 1, 3)) + np.cumsum(rng.normal(0, 0.01, (F, P, J, 3)), axis=0)
    nout = rng.integers(0, P + 2, F).astype(np.int32)      # sometimes more persons than slots
    nout[0] = 2
    ref, nsm_ref, _ = c_oracle.smooth(pts, np.minimum(nout, P), 2.5, 0.75, 0.4, 0.03333333333)
    st = SmoothState(sv.triangulation._util_engine(), P, J, 2.5, 0.75, 0.4)
    st.set_chunked(chunked)
    dense = np.zeros((F, P, J, 4))
    dense[..., :3] = pts
    t0, got, nsm = 0, [], []
    for b in batches:
        out = torch.from_numpy(dense[t0:t0 + b].copy()).cuda()
        nsm.append(st.run(out, torch.from_numpy(nout[t0:t0 + b].copy()).cuda(), 0.03333333333).cpu().numpy())
        got.append(out.cpu().numpy())
        t0 += b
    got, nsm = np.concatenate(got), np.concatenate(nsm)
    assert np.array_equal(nsm, nsm_ref)
    m = np.arange(P)[None, :] < nsm[:, None]
    np.testing.assert_allclose(got[m][..., :3], ref[m], rtol=1e-10, atol=1e-12)
    np.testing.assert_array_equal(got[~m][..., :3], pts[~m])


# ---- ragged ingestion ---------------------------------------------------------------------
@pytest.mark.parametrize("P", [None, 2])
def test_pack_detections_equals_reference_call_sequence(torch_cuda, P):
    """Ragged detector outputs -> dense batch on the device == what add_human_2D_points builds frame by frame,
    and the fused path on it == the loop oracle fed the same lists (persons beyond P slots are dropped)."""
    torch = torch_cuda
    from oracle import loop_oracle
    rng = np.random.default_rng(3)
    rig = synth.ring_rig(5)
    F, J, Pmax = 9, 17, 3
    d = synth.make_frames(rig, F, Pmax, J, seed=31, drop_prob=0.4)
    dets = [[(d["kpts"][f, c, :d["counts"][f, c]], d["scores"][f, c, :d["counts"][f, c]]) for c in range(rig.C)]
            for f in range(F)]
    prm = synth.MULTI_PARAMS
    eng = _engine(rig, prm)
    kp, sc, cn = eng.pack_detections(dets, P=P)
    torch.cuda.synchronize()
    Pd = Pmax if P is None else P
    want_counts = np.minimum(d["counts"], Pd)
    assert np.array_equal(cn.cpu().numpy(), want_counts)
    want_k = np.zeros((F, rig.C, Pd, J, 2), np.float32)
    want_s = np.zeros((F, rig.C, Pd, J), np.float32)
    for f in range(F):
        for c in range(rig.C):
            n = want_counts[f, c]
            want_k[f, c, :n], want_s[f, c, :n] = d["kpts"][f, c, :n], d["scores"][f, c, :n]
    assert np.array_equal(kp.cpu().numpy(), want_k) and np.array_equal(sc.cpu().numpy(), want_s)
    res = eng.run(kp, sc, cn, Pout=8)
    torch.cuda.synchronize()
    out, nout = res["out"].cpu().numpy(), res["nout"].cpu().numpy()
    for f in range(F):
        ref = loop_oracle.fused_frame(want_k[f], want_s[f], want_counts[f], rig.K, rig.R, rig.t, prm)
        n = len(ref["hrnet_triangulate_points"])
        assert nout[f] == n
        if n:
            assert rel_l2(out[f, :n, :, :3], np.array(ref["hrnet_triangulate_points"])) < TOL_FUSED


# ---- properties --------------------------------------------------------------------------
def test_property_score_scaling(torch_cuda):
    """Scaling every detector score by k > 0 (thresholds that do not gate on the score) scales the keypoint and
    person scores by k and leaves the 3D points unchanged (weights enter only as ratios)."""
    torch = torch_cuda
    rig = floor_rig()
    d = synth.make_frames(rig, 200, 1, 133, seed=41)
    prm = dict(synth.DEFAULT_PARAMS, kst=0.0)
    kp, sc, cn = _to_dev(torch, d["kpts"], d["scores"], d["counts"])
    for precision in ("f64", "mixed"):
        eng = _engine(rig, prm, precision=precision)
        a = eng.run(kp, sc, cn, Pout=1)
        b = eng.run(kp, sc * 4.0, cn, Pout=1)          # power of two: exact in floating point
        torch.cuda.synchronize()
        oa, ob = a["out"].cpu().numpy(), b["out"].cpu().numpy()
        assert np.array_equal(a["nout"].cpu().numpy(), b["nout"].cpu().numpy())
        np.testing.assert_allclose(ob[..., :3], oa[..., :3], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(ob[..., 3], 4.0 * oa[..., 3], rtol=1e-6)
        np.testing.assert_allclose(b["pscores"].cpu().numpy(), 4.0 * a["pscores"].cpu().numpy(), rtol=1e-5)


def test_property_frames_are_independent(torch_cuda):
    """The path has no cross-frame state: any permutation of the frames permutes the outputs (this is what lets
    frames shard across GPUs), for the single-person and the general kernel."""
    torch = torch_cuda
    rng = np.random.default_rng(0)
    for rig, P, prm, pout in ((floor_rig(), 1, synth.DEFAULT_PARAMS, 1), (synth.ring_rig(5), 3, synth.MULTI_PARAMS, 6)):
        d = synth.make_frames(rig, 173, P, 33, seed=43, low_score_frac=0.1, drop_prob=0.1)
        perm = rng.permutation(173)
        eng = _engine(rig, prm)
        a = eng.run(*_to_dev(torch, d["kpts"], d["scores"], d["counts"]), Pout=pout)
        b = eng.run(*_to_dev(torch, d["kpts"][perm], d["scores"][perm], d["counts"][perm]), Pout=pout)
        torch.cuda.synchronize()
        assert np.array_equal(a["nout"].cpu().numpy()[perm], b["nout"].cpu().numpy())
        assert np.array_equal(a["out"].cpu().numpy()[perm], b["out"].cpu().numpy())
        np.testing.assert_allclose(a["pscores"].cpu().numpy()[perm], b["pscores"].cpu().numpy(), rtol=1e-5, atol=1e-7)


def test_property_person_order_within_a_camera(torch_cuda):
    """Detector order inside a camera is arbitrary (main.py:54): permuting it may reorder the emitted persons but
    the kernel must keep agreeing with the oracle, which follows the reference's list order."""
    torch = torch_cuda
    from oracle import c_oracle
    rig = synth.ring_rig(6)
    d = synth.make_frames(rig, 60, 3, 17, seed=47, shuffle=False)
    rng = np.random.default_rng(1)
    k2, s2 = d["kpts"].copy(), d["scores"].copy()
    for f in range(60):
        for c in range(rig.C):
            p = rng.permutation(3)
            k2[f, c], s2[f, c] = d["kpts"][f, c][p], d["scores"][f, c][p]
    prm = synth.MULTI_PARAMS
    ref = c_oracle.fused(k2, s2, d["counts"], rig.K, rig.R, rig.t, prm, Pout=8)
    eng = _engine(rig, prm)
    res = eng.run(*_to_dev(torch, k2, s2, d["counts"]), Pout=8)
    torch.cuda.synchronize()
    nout = res["nout"].cpu().numpy()
    assert np.array_equal(nout, ref["nout"])
    valid = np.arange(8)[None, :] < np.minimum(nout, 8)[:, None]
    assert rel_l2(res["out"].cpu().numpy()[valid][..., :3], ref["points"][valid]) < TOL_FUSED


# ---- BASELINE configs[3] and [4] geometry: rigs far beyond what fits in shared memory -------------------------
@pytest.mark.parametrize("C,P,F,precision", [(16, 8, 3, "f64"), (16, 8, 3, "mixed"), (32, 16, 1, "f64"), (32, 16, 1, "mixed")])
def test_large_rigs_streaming_path(torch_cuda, C, P, F, precision):
    """16 cameras x 8 persons and 32 cameras x 16 persons x 133 joints (7 680 / 126 976 candidates per frame):
    the streaming general path against the C oracle."""
    torch = torch_cuda
    from oracle import c_oracle
    rig = synth.ring_rig(C)
    d = synth.make_frames(rig, F, P, 133, seed=400 + C, low_score_frac=0.05)
    prm = synth.MULTI_PARAMS
    pout = 2 * P
    ref = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, prm, Pout=pout)
    eng = _engine(rig, prm, precision=precision)
    res = eng.run(*_to_dev(torch, d["kpts"], d["scores"], d["counts"]), Pout=pout)
    torch.cuda.synchronize()
    # float modes: second-generation matching kernel (ray pre-pass + tiles from global memory), first-generation fuse
    assert eng.last_launch_info()["kernel"] == ("general" if precision == "f64" else "general2m")
    out, nout = res["out"].cpu().numpy(), res["nout"].cpu().numpy()
    assert np.array_equal(nout, ref["nout"])
    valid = np.arange(pout)[None, :] < np.minimum(nout, pout)[:, None]
    tol = TOL_FUSED if precision == "f64" else TOL_NORTH_STAR / 10
    assert rel_l2(out[valid][:, :, :3], ref["points"][valid]) < tol
    assert np.array_equal(out[valid][:, :, 3] == 0, ref["kscores"][valid] == 0)


# ---- second-generation kernels of the general path (snowtri_match.cuh, snowtri_mfuse.cuh) -----------------------
def _run_general(torch, rig, d, prm, pout, precision, generation, jout=None, tune_g=0):
    eng = _engine(rig, prm, precision=precision)
    eng.set_general_kernels(generation)
    if tune_g:
        eng.set_tuning(tune_g, 0, 0)
    n0 = eng.launch_count
    res = eng.run(*_to_dev(torch, d["kpts"], d["scores"], d["counts"]), Pout=pout, keypoint_num=jout)
    torch.cuda.synchronize()
    return ({k: v.cpu().numpy() for k, v in res.items()}, eng.last_launch_info()["kernel"], eng.launch_count - n0)


@pytest.mark.parametrize("C,P,J,F,drop,low", [(8, 4, 133, 40, 0.0, 0.05), (8, 4, 133, 24, 0.25, 0.1), (3, 5, 33, 50, 0.2, 0.1),
                                              (6, 8, 17, 30, 0.1, 0.3), (4, 2, 133, 64, 0.0, 0.0), (2, 9, 5, 20, 0.3, 0.0),
                                              (7, 1, 64, 33, 0.2, 0.1), (5, 3, 1, 40, 0.1, 0.0)])
def test_general_second_generation_vs_oracle_and_first(torch_cuda, C, P, J, F, drop, low):
    """BASELINE configs[2] geometry and ragged / partial-tile shapes (P not a multiple of the 4 x 4 tile, J below a
    warp, absent cameras): the three-launch path (matching + centres, clustering + member decode, clique fuse + person
    score) against the C oracle and against the first-generation six-launch path on the same batch."""
    torch = torch_cuda
    from oracle import c_oracle
    rig = synth.ring_rig(C, seed=C)
    d = synth.make_frames(rig, F, P, J, seed=900 + C * 10 + P, low_score_frac=low, drop_prob=drop)
    prm = dict(synth.MULTI_PARAMS, center=min(synth.MULTI_PARAMS["center"], J - 1))
    pout = min(2 * P, 32)
    ref = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, prm, Pout=pout)
    new, kn, ln = _run_general(torch, rig, d, prm, pout, "mixed", 2)
    old, ko, lo = _run_general(torch, rig, d, prm, pout, "mixed", 1)
    assert kn == "general2" and ln == 3, (kn, ln)
    assert ko == "general" and lo == 6, (ko, lo)
    valid = np.arange(pout)[None, :] < np.minimum(ref["nout"], pout)[:, None]
    for res, tag in ((new, "gen2"), (old, "gen1")):
        assert np.array_equal(res["nout"], ref["nout"]), tag
        assert not res["out"][~valid].any() and not res["pscores"][~valid].any(), tag
        if valid.any():
            assert np.array_equal(res["out"][valid][:, :, 3] == 0, ref["kscores"][valid] == 0), tag
            assert rel_l2(res["out"][valid][:, :, :3], ref["points"][valid]) < TOL_FUSED * 5, tag
            np.testing.assert_allclose(res["out"][valid][:, :, 3], ref["kscores"][valid], rtol=1e-4, atol=1e-7, err_msg=tag)
            np.testing.assert_allclose(res["pscores"][valid], ref["pscores"][valid], rtol=1e-4, atol=1e-7, err_msg=tag)


@pytest.mark.parametrize("prm_over", [dict(ast=0.0), dict(ast=0.0, cond_tol=10.0), dict(ast=0.05, cond_tol=1.0, num_tol=3),
                                      dict(dthr=0.01, ast=0.5), dict(kst=0.0, ast=0.3), dict(ast=5.0)])
def test_general_second_generation_thresholds(torch_cuda, prm_over):
    """Everything kept (ghost clusters that are not cliques take the rolled member loop), huge merge radius (one
    cluster with every candidate), cluster-size filter, tight gate, no keypoint threshold, nothing kept."""
    torch = torch_cuda
    from oracle import c_oracle
    rig = synth.ring_rig(6, seed=3)
    d = synth.make_frames(rig, 48, 3, 133, seed=77, low_score_frac=0.1, drop_prob=0.15)
    prm = dict(synth.MULTI_PARAMS, **prm_over)
    pout = 7
    ref = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, prm, Pout=pout)
    res, kn, ln = _run_general(torch, rig, d, prm, pout, "mixed", 2)
    assert kn == "general2" and ln == 3
    assert np.array_equal(res["nout"], ref["nout"])
    valid = np.arange(pout)[None, :] < np.minimum(ref["nout"], pout)[:, None]
    assert not res["out"][~valid].any()
    if valid.any():
        assert np.array_equal(res["out"][valid][:, :, 3] == 0, ref["kscores"][valid] == 0)
        # clusters that swallow wrongly matched candidates fuse midpoints metres apart: float32 weights hold the
        # north_star bound there, not the 1e-5 of correctly matched persons
        assert rel_l2(res["out"][valid][:, :, :3], ref["points"][valid]) < TOL_NORTH_STAR
        np.testing.assert_allclose(res["pscores"][valid], ref["pscores"][valid], rtol=2e-4, atol=1e-7)


@pytest.mark.parametrize("noise,prm_over", [(0.0, dict(ast=4e4)), (0.0, dict(ast=2e4, cond_tol=0.05)), (0.0, dict(dthr=1e-5, ast=0.2)),
                                            (0.02, dict(ast=25.0)), (0.5, dict(dthr=0.0502, ast=0.21))])
def test_general_matching_bounds_edge_cases(torch_cuda, noise, prm_over):
    """The matching kernel decides from float32 bounds and falls back to float64 where they cannot decide: noise-free
    detections (rays that pass within micrometres: joints without an upper bound, heavy-tailed scores, thresholds in
    the middle of the score distribution), a distance threshold below the float32 error of a ray distance (no joint
    surely passes), thresholds a hair off the defaults.  The decisions (persons per frame, zero pattern) must be the
    float64 oracle's in every case."""
    torch = torch_cuda
    from oracle import c_oracle
    rig = synth.ring_rig(8, seed=2)
    d = synth.make_frames(rig, 32, 4, 133, seed=4242, noise_px=noise, low_score_frac=0.05, drop_prob=0.1)
    prm = dict(synth.MULTI_PARAMS, **prm_over)
    pout = 8
    ref = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, prm, Pout=pout)
    res, kn, ln = _run_general(torch, rig, d, prm, pout, "mixed", 2)
    assert kn == "general2" and ln == 3
    assert np.array_equal(res["nout"], ref["nout"])
    valid = np.arange(pout)[None, :] < np.minimum(ref["nout"], pout)[:, None]
    assert not res["out"][~valid].any()
    if valid.any():
        assert np.array_equal(res["out"][valid][:, :, 3] == 0, ref["kscores"][valid] == 0)
        assert rel_l2(res["out"][valid][:, :, :3], ref["points"][valid]) < TOL_NORTH_STAR
        # a candidate kept or dropped differently would change its cluster's member count, i.e. every keypoint score
        np.testing.assert_allclose(res["out"][valid][:, :, 3], ref["kscores"][valid], rtol=2e-3, atol=1e-7)


def test_general_second_generation_chunks_and_truncation(torch_cuda):
    """Several scratch chunks per batch (frames_per_group caps the chunk) and keypoint_num < J give the same bytes /
    the truncated rows of the one-chunk run."""
    torch = torch_cuda
    rig = synth.ring_rig(8)
    d = synth.make_frames(rig, 70, 4, 133, seed=5, low_score_frac=0.05, drop_prob=0.1)
    prm = synth.MULTI_PARAMS
    one, _, l1 = _run_general(torch, rig, d, prm, 8, "mixed", 2)
    many, _, l2 = _run_general(torch, rig, d, prm, 8, "mixed", 2, tune_g=16)
    assert l1 == 3 and l2 == 3 * 5
    assert np.array_equal(one["nout"], many["nout"]) and np.array_equal(one["out"], many["out"])
    np.testing.assert_allclose(one["pscores"], many["pscores"], rtol=1e-5, atol=1e-7)
    trunc, _, _ = _run_general(torch, rig, d, dict(prm, center=3), 8, "mixed", 2, jout=40)
    full, _, _ = _run_general(torch, rig, d, dict(prm, center=3), 8, "mixed", 2)
    assert np.array_equal(trunc["nout"], full["nout"]) and np.array_equal(trunc["out"], full["out"][:, :, :40])


@pytest.mark.parametrize("name", __import__("conftest").big_golden_names())
@pytest.mark.parametrize("precision", ["f64", "mixed"])
def test_full_size_reference_goldens(torch_cuda, name, precision):
    """What the REAL reference emitted at BASELINE configs[2] (8 x 4 x 133, 2 frames) and configs[3] (16 x 8 x 133,
    1 frame) sizes (tests/golden/make_golden_big.py) against the fused path."""
    torch = torch_cuda
    from conftest import BigGolden
    g = BigGolden(name)
    pout = max(c[0].shape[0] for c in g.con)
    if pout > 32:
        pout = 32
    eng = _engine(g, g.params, precision=precision)
    res = eng.run(*_to_dev(torch, g.kpts, g.scores, g.counts), Pout=pout, keypoint_num=g.params["keypoint_num"])
    torch.cuda.synchronize()
    out, ps, nout = res["out"].cpu().numpy(), res["pscores"].cpu().numpy(), res["nout"].cpu().numpy()
    for f in range(g.F):
        pts, ks, pscore = g.con[f]
        assert nout[f] == pts.shape[0]
        n = min(pts.shape[0], pout)
        assert rel_l2(out[f, :n, :, :3], pts[:n]) < (TOL_FUSED if precision == "f64" else TOL_FUSED * 5)
        np.testing.assert_allclose(out[f, :n, :, 3], ks[:n], rtol=1e-5 if precision == "f64" else 1e-4, atol=1e-7)
        np.testing.assert_allclose(ps[f, :n], pscore[:n], rtol=1e-5 if precision == "f64" else 1e-4, atol=1e-7)


# ---- empty and degenerate inputs ---------------------------------------------------------------------------------
def test_empty_batch_and_empty_frames(torch_cuda):
    """F = 0 is a no-op; frames in which no camera sees anybody emit no person (all kernels, all outputs zeroed)."""
    torch = torch_cuda
    rig = floor_rig()
    for P, prm in ((1, synth.DEFAULT_PARAMS), (3, synth.MULTI_PARAMS)):
        eng = _engine(rig, prm, precision="mixed")
        z = eng.run(torch.empty((0, 4, P, 17, 2), device="cuda"), torch.empty((0, 4, P, 17), device="cuda"),
                    torch.empty((0, 4), dtype=torch.int32, device="cuda"), Pout=2)
        assert z["out"].shape == (0, 2, 17, 4) and z["nout"].shape == (0,)
        d = synth.make_frames(rig, 40, P, 17, seed=51)
        counts = d["counts"].copy()
        counts[::3] = 0                      # every third frame: nobody detected
        counts[1::3, 1:] = 0                 # every third frame: one camera only -> no pair, no candidate
        res = eng.run(*_to_dev(torch, d["kpts"], d["scores"], counts), Pout=2)
        torch.cuda.synchronize()
        nout, out, ps = res["nout"].cpu().numpy(), res["out"].cpu().numpy(), res["pscores"].cpu().numpy()
        assert not nout[::3].any() and not nout[1::3].any()
        assert not out[::3].any() and not out[1::3].any() and not ps[::3].any() and not ps[1::3].any()
        from oracle import c_oracle
        ref = c_oracle.fused(d["kpts"], d["scores"], counts, rig.K, rig.R, rig.t, prm, Pout=2)
        assert np.array_equal(nout, ref["nout"])


def test_single_camera_rig_has_no_candidates(torch_cuda):
    """One camera: no camera pair, so Human_Triangulation yields nothing (reference :56-58 loops are empty)."""
    torch = torch_cuda
    rig = synth.ring_rig(4).subset(1)
    d = synth.make_frames(rig, 5, 2, 17, seed=52)
    eng = _engine(rig, synth.MULTI_PARAMS)
    res = eng.run(*_to_dev(torch, d["kpts"], d["scores"], d["counts"]), Pout=2)
    torch.cuda.synchronize()
    assert not res["nout"].cpu().numpy().any() and not res["out"].cpu().numpy().any()


# ---- rig-specialised single-person kernel (NVRTC) ----------------------------------------------------------------
@pytest.mark.parametrize("precision", ["f32", "mixed"])
@pytest.mark.parametrize("case", [0, 2, 3, 7])
def test_jit_specialised_kernel_matches_precompiled(torch_cuda, precision, case):
    """snowtri_set_jit(always): the kernel compiled at run time for this rig and batch shape must reproduce the
    precompiled kernel (same source, constants as immediates) and the oracle; a second run hits the cache."""
    torch = torch_cuda
    from oracle import c_oracle
    rname, F, J, prm, pout, knum, kw = P1_CASES[case]
    rig = _rig(rname)
    d = synth.make_frames(rig, F, 1, J, seed=600 + case, **kw)
    jout = knum or J
    kp, sc, cn = _to_dev(torch, d["kpts"], d["scores"], d["counts"])
    eng = _engine(rig, prm, precision=precision)
    eng.set_jit("off")
    base = eng.run(kp, sc, cn, Pout=pout, keypoint_num=jout)
    torch.cuda.synchronize()
    assert eng.last_launch_info()["kernel"] == "p1"
    base = {k: v.clone() for k, v in base.items()}
    eng.set_jit("always")
    for rep in range(2):
        res = eng.run(kp, sc, cn, Pout=pout, keypoint_num=jout)
        torch.cuda.synchronize()
        assert eng.last_launch_info()["kernel"] == "p1-jit", eng.jit_status
        assert eng.jit_status.startswith("compiled")
        assert torch.equal(res["nout"], base["nout"])
        np.testing.assert_allclose(res["out"].cpu().numpy(), base["out"].cpu().numpy(), rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(res["pscores"].cpu().numpy(), base["pscores"].cpu().numpy(), rtol=2e-5, atol=1e-6)
    ref = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, prm, Pout=pout, keypoint_num=jout)
    nout = res["nout"].cpu().numpy()
    assert np.array_equal(nout, ref["nout"])
    valid = np.arange(pout)[None, :] < np.minimum(nout, pout)[:, None]
    if valid.any():
        assert rel_l2(res["out"].cpu().numpy()[valid][:, :, :3], ref["points"][valid]) < TOL_NORTH_STAR / 10
    eng.set_params(dthr=0.04)      # a threshold is baked in too: changing it must compile a new kernel, not reuse
    res2 = eng.run(kp, sc, cn, Pout=pout, keypoint_num=jout)
    torch.cuda.synchronize()
    ref2 = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, dict(prm, dthr=0.04), Pout=pout,
                          keypoint_num=jout)
    assert np.array_equal(res2["out"].cpu().numpy()[..., 3] == 0, np.where(np.arange(pout)[None, :, None] < np.minimum(ref2["nout"], pout)[:, None, None], ref2["kscores"] == 0, True))


def test_jit_failure_falls_back_to_precompiled_kernel(torch_cuda, monkeypatch):
    """A runtime-compilation failure is soft: the precompiled kernel runs, the result is unchanged and the reason is
    reported.  (The failure is provoked with an experiment hook that injects a macro into the generated source.)"""
    torch = torch_cuda
    rig = floor_rig()
    d = synth.make_frames(rig, 50, 1, 133, seed=61)
    kp, sc, cn = _to_dev(torch, d["kpts"], d["scores"], d["counts"])
    eng = _engine(rig, synth.DEFAULT_PARAMS, precision="f32")
    eng.set_jit("off")
    base = eng.run(kp, sc, cn, Pout=1)["out"].clone()
    monkeypatch.setenv("SNOWTRI_JIT_DEFINES", "P1_KEEP=(")
    eng.set_jit("always")
    res = eng.run(kp, sc, cn, Pout=1)
    torch.cuda.synchronize()
    assert eng.last_launch_info()["kernel"] == "p1"
    assert eng.jit_status.startswith("failed: nvrtc"), eng.jit_status
    assert torch.equal(res["out"], base)


# ---- optional DLT mode ------------------------------------------------------------------------
@pytest.mark.parametrize("rname,f64acc,tol", [("floor4", False, 2e-5), ("floor4", True, 1e-6), ("ring8", False, 2e-5),
                                                ("ring2", True, 1e-6)])
def test_dlt_mode_matches_numpy_svd(torch_cuda, rname, f64acc, tol):
    """Homogeneous linear triangulation against np.linalg.svd of the same 2V x 4 system (not a reference parity
    test: snowvision has no DLT).  Also recovers the synthetic ground truth to the noise level."""
    torch = torch_cuda
    from oracle import dlt_oracle
    rig = _rig(rname)
    d = synth.make_frames(rig, 40, 1, 133, seed=71, low_score_frac=0.15, shuffle=False)
    ref = dlt_oracle.dlt_points(d["kpts"], d["scores"], rig.K, rig.R, rig.t, kst=0.5)
    eng = _engine(rig, synth.DEFAULT_PARAMS)
    out = eng.dlt(*_to_dev(torch, d["kpts"], d["scores"]), accumulate_f64=f64acc).cpu().numpy()
    assert np.array_equal(out[..., 3], ref[..., 3])
    m = ref[..., 3] >= 2
    assert not out[~m][..., :3].any()
    assert rel_l2(out[m][..., :3], ref[m][..., :3]) < tol
    full = ref[..., 3] == rig.C
    assert np.abs(ref[full][..., :3] - d["truth"][:, 0][full]).max() < 0.05     # 0.5 px noise at ~5 m: centimetres


# ---- randomized differential test: random rigs, shapes and thresholds against the C oracle ---------------------
@pytest.mark.parametrize("seed", range(60))
def test_random_configurations_vs_c_oracle(torch_cuda, seed):
    """Random camera count, persons, joints, ragged counts and thresholds (including the reference's defaults,
    zero / tiny / huge tolerances, num_tol above the cluster sizes, truncated keypoint_num).  float64 must match the
    oracle to storage precision; the float modes must emit the same persons and hold the north_star bound."""
    torch = torch_cuda
    from oracle import c_oracle
    rng = np.random.default_rng(1000 + seed)
    C = int(rng.integers(2, 9))
    P = int(rng.choice([1, 1, 2, 3]))
    J = int(rng.choice([1, 5, 17, 33, 64, 133]))
    F = int(rng.integers(1, 70))
    rig = synth.ring_rig(C, seed=seed)
    d = synth.make_frames(rig, F, P, J, seed=2000 + seed, low_score_frac=float(rng.choice([0.0, 0.1, 0.5])),
                          drop_prob=float(rng.choice([0.0, 0.2, 0.6])), noise_px=float(rng.choice([0.3, 1.0, 3.0])))
    prm = dict(kst=float(rng.choice([0.0, 0.5, 0.7])), ast=float(rng.choice([0.0, 0.0, 0.05, 0.3])),
               dthr=float(rng.choice([0.05, 0.01, 0.5])), cond_tol=float(rng.choice([0.1, 0.3, 10.0, 0.005])),
               num_tol=int(rng.choice([0, 0, 2, 4])), score_tol=float(rng.choice([0.0, 0.0, 0.0, 0.05])),
               center=int(rng.integers(0, J)))
    jout = int(rng.integers(1, J + 1)) if rng.random() < 0.3 else J
    pout = int(rng.integers(1, 6))
    ref = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, prm, Pout=pout, keypoint_num=jout)
    kp, sc, cn = _to_dev(torch, d["kpts"], d["scores"], d["counts"])
    m = np.minimum(ref["nout"], pout)
    valid = np.arange(pout)[None, :] < m[:, None]
    for precision in ("f64", "mixed", "f32"):
        eng = _engine(rig, prm, precision=precision)
        if seed % 3 == 0:
            eng.set_jit("always")
        res = eng.run(kp, sc, cn, Pout=pout, keypoint_num=jout)
        torch.cuda.synchronize()
        out, nout = res["out"].cpu().numpy(), res["nout"].cpu().numpy()
        tag = f"C={C} P={P} J={J} F={F} {precision} {prm} jout={jout} pout={pout} kernel={eng.last_launch_info()['kernel']}"
        assert np.array_equal(nout, ref["nout"]), tag
        assert not out[~valid].any(), tag
        if valid.any():
            assert np.array_equal(out[valid][:, :, 3] == 0, ref["kscores"][valid] == 0), tag
            tol = TOL_FUSED if precision == "f64" else TOL_NORTH_STAR
            assert rel_l2(out[valid][:, :, :3], ref["points"][valid]) < tol, tag


# ---- final all-gather through the C ABI -----------------------------------------------------------
def test_native_allgather_single_rank(torch_cuda):
    """snowtri_comm_unique_id / snowtri_comm_init / snowtri_allgather with a one-rank communicator (the round-end GPU
    box has one GPU; tools/mgpu_check.py under tests/test_multi_gpu.py is the multi-rank check)."""
    torch = torch_cuda
    import ctypes as ct
    from snowmocap_b200 import _lib
    rig = floor_rig()
    eng = _engine(rig, synth.DEFAULT_PARAMS)
    lib = eng._lib
    ident = ct.create_string_buffer(128)
    _lib.check(lib.snowtri_comm_unique_id(ident))
    assert any(ident.raw)
    x = torch.arange(4096, dtype=torch.float32, device="cuda")
    y = torch.zeros_like(x)
    rc = lib.snowtri_allgather(eng._h, x.data_ptr(), y.data_ptr(), x.numel() * 4, None, None)
    assert rc == _lib.E_ARG, "no communicator yet"
    _lib.check(lib.snowtri_comm_init(eng._h, ident, 1, 0), eng._h)
    assert lib.snowtri_comm_init(eng._h, ident, 1, 0) == _lib.E_ARG, "a handle owns at most one communicator"
    _lib.check(lib.snowtri_allgather(eng._h, x.data_ptr(), y.data_ptr(), x.numel() * 4, None,
                                     torch.cuda.current_stream().cuda_stream), eng._h)
    torch.cuda.synchronize()
    assert torch.equal(x, y)
    assert lib.snowtri_comm_destroy(eng._h) == _lib.OK
    eng.close()
