"""Pin both CPU oracles to the golden vectors produced by the real reference (CPU only)."""
import os

import numpy as np
import pytest

from conftest import Golden, golden_names, rel_l2
from oracle import c_oracle, loop_oracle
from snowmocap_b200 import synth

TOL = 1e-10   # FP64 restatement vs reference (adjugate inverse instead of LAPACK getri)


def _check(got, want, what):
    pts, ks, ps = want
    assert len(got[0]) == pts.shape[0], f"{what}: count {len(got[0])} != {pts.shape[0]}"
    if pts.shape[0] == 0:
        return
    assert rel_l2(np.array(got[0]), pts) < TOL, what
    # scores are ~1/dist: compare with a relative tolerance per element
    np.testing.assert_allclose(np.array(got[1]), ks, rtol=1e-8, atol=1e-12, err_msg=what)
    np.testing.assert_allclose(np.array(got[2]), ps, rtol=1e-8, atol=1e-12, err_msg=what)


def test_loop_oracle_matches_reference(golden):
    g, p = golden, golden.params
    for f in range(g.F):
        tri = loop_oracle.triangulate_frame(g.kpts[f], g.scores[f], g.counts[f], g.K, g.R, g.t,
                                            kst=p["kst"], ast=p["ast"], dthr=p["dthr"])
        _check((tri[loop_oracle.POINTS], tri[loop_oracle.KSCORES], tri[loop_oracle.PSCORES]), g.tri[f],
               f"{g.name} frame {f} triangulate")
        con = loop_oracle.condense_frame(tri, tol=p["cond_tol"], num_tol=p["num_tol"], score_tol=p["score_tol"],
                                         center=p["center"], keypoint_num=p["keypoint_num"])
        _check((con[loop_oracle.POINTS], con[loop_oracle.KSCORES], con[loop_oracle.PSCORES]), g.con[f],
               f"{g.name} frame {f} condense")


def test_c_oracle_candidates_match_reference(golden):
    g, p = golden, golden.params
    for f in range(g.F):
        c = c_oracle.candidates(g.kpts[f], g.scores[f], g.counts[f], g.K, g.R, g.t, p["kst"], p["ast"], p["dthr"])
        _check((c["points"], c["kscores"], c["pscores"]), g.tri[f], f"{g.name} frame {f}")


def test_c_oracle_condense_matches_reference(golden):
    g, p = golden, golden.params
    for f in range(g.F):
        pts, ks, _ = g.tri[f]
        c = c_oracle.condense(pts, ks, p["cond_tol"], p["num_tol"], p["score_tol"], p["center"], p["keypoint_num"])
        _check((c["points"], c["kscores"], c["pscores"]), g.con[f], f"{g.name} frame {f}")


def test_c_oracle_fused_matches_reference(golden):
    g, p = golden, golden.params
    pout = max(1, max(c[0].shape[0] for c in g.con))
    r = c_oracle.fused(g.kpts, g.scores, g.counts, g.K, g.R, g.t, p, Pout=pout, keypoint_num=p["keypoint_num"])
    for f in range(g.F):
        n = g.con[f][0].shape[0]
        assert r["nout"][f] == n
        assert r["ncand"][f] == g.tri[f][0].shape[0]
        _check((r["points"][f, :n], r["kscores"][f, :n], r["pscores"][f, :n]), g.con[f], f"{g.name} frame {f}")


@pytest.mark.parametrize("name", __import__("conftest").big_golden_names())
def test_c_oracle_fused_matches_reference_at_full_size(name):
    """BASELINE configs[2] (8 cameras x 4 persons x 133 joints, 2 frames) and configs[3] (16 x 8 x 133, 1 frame):
    the C oracle against what the REAL reference emitted (tests/golden/make_golden_big.py)."""
    from conftest import BigGolden
    g = BigGolden(name)
    p = g.params
    pout = max(1, max(c[0].shape[0] for c in g.con))
    r = c_oracle.fused(g.kpts, g.scores, g.counts, g.K, g.R, g.t, p, Pout=pout, keypoint_num=p["keypoint_num"])
    for f in range(g.F):
        n = g.con[f][0].shape[0]
        assert r["nout"][f] == n
        assert r["ncand"][f] == g.tri_ps[f].shape[0]
        _check((r["points"][f, :n], r["kscores"][f, :n], r["pscores"][f, :n]), g.con[f], f"{name} frame {f}")


def test_fused_threads_agree():
    rig = synth.ring_rig(6)
    d = synth.make_frames(rig, 24, 2, 17, seed=5, low_score_frac=0.1, drop_prob=0.1)
    a = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, synth.MULTI_PARAMS, Pout=6, nthreads=1)
    b = c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, synth.MULTI_PARAMS, Pout=6, nthreads=4)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_skew_ray_known_answer():
    # two perpendicular rays through (1,0,0)/(0,1,0) offset by 2 along z: midpoint z=1, distance 2
    hm, hs = np.array([[1.0, 0, 0]]), np.array([[0, 1.0, 0]])
    tm, ts = np.array([[-1.0, 0, 0]]), np.array([[0, -1.0, 2.0]])
    dist, W = c_oracle.skew_ray(hm, hs, tm, ts)
    assert abs(dist[0] - 2.0) < 1e-14 and np.allclose(W[0], [0, 0, 1.0], atol=1e-14)
    d2, W2 = loop_oracle.skew_ray_solve(hm.T, hs.T, tm.T, ts.T)
    assert abs(d2 - 2.0) < 1e-14 and np.allclose(W2, [0, 0, 1.0], atol=1e-14)


def test_single_candidate_condense_is_empty():
    # BASELINE config 1 (C=2, P=1) yields one candidate and zero condensed persons
    g = Golden("cfg1_c2p1j17")
    assert all(t[0].shape[0] == 1 for t in g.tri) and all(c[0].shape[0] == 0 for c in g.con)


def test_noise_free_input_gives_nan_like_reference():
    # exact projections -> dist==0 -> inf score -> NaN fused point (not special-cased)
    rig = synth.ring_rig(3)
    X = np.array([[[0.25, -0.5, 1.0]]])
    uv = synth.project(rig, X)                     # (C,1,1,2) exact float64
    hm = loop_oracle.back_project(rig.K[0], rig.R[0], uv[0, 0, 0])
    hs = loop_oracle.back_project(rig.K[1], rig.R[1], uv[1, 0, 0])
    dist, _ = loop_oracle.skew_ray_solve(hm, hs, rig.t[0].reshape(3, 1), rig.t[1].reshape(3, 1))
    assert dist < 1e-12


# ---- temporal smoothing: both restatements against the real reference's outputs ----
from conftest import GOLDEN, smooth_golden_names  # noqa: E402


def _smooth_golden(name):
    import json
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    counts = z["counts"]
    return z["pts"], counts, json.loads(str(z["params"])), [z[f"out_{t}"] for t in range(len(counts))]


@pytest.mark.parametrize("name", smooth_golden_names())
def test_loop_oracle_smooth_matches_reference(name):
    from oracle import loop_oracle
    pts, counts, prm, want = _smooth_golden(name)
    got = loop_oracle.smooth_sequence([[pts[t, k] for k in range(counts[t])] for t in range(len(counts))],
                                      f=prm["f"], z=prm["z"], r=prm["r"], delta_time=prm["dt"])
    for t, (g, w) in enumerate(zip(got, want)):
        assert g.shape[0] == w.shape[0], f"frame {t}"
        if w.size:
            np.testing.assert_allclose(g, w, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name", smooth_golden_names())
@pytest.mark.parametrize("split", [None, 1, 7])
def test_c_oracle_smooth_matches_reference(name, split):
    """Dense C restatement, also fed in chunks with the state carried over (streaming a clip)."""
    from oracle import c_oracle
    pts, counts, prm, want = _smooth_golden(name)
    F = len(counts)
    step = F if split is None else split
    state, outs, nsm = None, [], []
    for t0 in range(0, F, step):
        o, n, state = c_oracle.smooth(pts[t0:t0 + step], counts[t0:t0 + step], prm["f"], prm["z"], prm["r"],
                                      prm["dt"], state=state)
        outs.append(o)
        nsm.append(n)
    out, nsm = np.concatenate(outs), np.concatenate(nsm)
    for t, w in enumerate(want):
        assert nsm[t] == w.shape[0], f"frame {t}"
        if w.size:
            np.testing.assert_allclose(out[t, :nsm[t]], w, rtol=1e-12, atol=1e-14)


def _forget_frames(f, z, r, T, limit=256):
    """Frames after which max|A^n| < 1e-19 for the state matrix of one present frame (what snowtri_smooth_run computes
    on the host to choose the one-pass path; csrc/snowtri_smooth.cu)."""
    k1, k2, k3 = z / (np.pi * f), 1 / (2 * np.pi * f) ** 2, r * z / (2 * np.pi * f)
    g = T / k2
    A = np.array([[0, 0, 0], [0, 1, T], [-k3 / k2, -g, 1 - g * (T + k1)]])
    pw = np.eye(3)
    for n in range(1, limit + 1):
        pw = A @ pw
        if np.abs(pw).max() < 1e-19:
            return n
    return 0


def test_follower_forgets_its_state_and_warm_up_chunks_equal_the_sequential_recurrence():
    """The one-pass smoothing path: with the reference's default filter a follower has forgotten any
    start state after 80 present frames, so a chunk that starts from a ZERO state that many frames early gives the
    sequential recurrence's output on its own frames -- checked here on the CPU with the C oracle as the walker."""
    from oracle import c_oracle
    assert _forget_frames(2.5, 0.75, 0.0, 1 / 30) == 80 and _forget_frames(2.0, 0.75, 0.0, 1 / 30) == 95
    assert _forget_frames(0.01, 0.75, 0.0, 1 / 30) == 0          # a slow filter keeps the chunk scan
    f, z, r, T = 2.5, 0.75, 0.4, 1 / 30
    W = _forget_frames(f, z, r, T)
    assert 0 < W <= 256
    rng = np.random.default_rng(5)
    F, P, J, L = 1000, 2, 5, 256
    pts = rng.uniform(-2, 2, (1, P, 1, 3)) + np.cumsum(rng.normal(0, 0.01, (F, P, J, 3)), axis=0)
    nout = np.full(F, P, np.int32)
    want, _, _ = c_oracle.smooth(pts, nout, f, z, r, T)
    for t0 in range(L, F, L):
        tw = t0 - W
        state = np.zeros(2 + 3 * P * J * 3)
        state[0], state[1] = 1.0, P                                # initialised clip, zero followers
        got, _, _ = c_oracle.smooth(pts[tw:t0 + L], nout[tw:t0 + L], f, z, r, T, state=state)
        np.testing.assert_allclose(got[W:], want[t0:t0 + L], rtol=1e-13, atol=1e-15)
