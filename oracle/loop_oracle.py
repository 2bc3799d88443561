"""TEST INFRASTRUCTURE ONLY -- scalar CPU restatement of the SnowMocap triangulation path.

This module is the *oracle*: a step-by-step FP64 restatement of the reference's
algorithm, written against array inputs instead of the reference's Camera objects.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it; the product package ``snowmocap_b200`` never does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference
(``/root/reference/snowvision``) in the build container, runs it on seeded inputs and
commits inputs + outputs under ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks
this module against every one of those fixtures (agreement <= 1e-12 relative).

Reference lines restated here (paths relative to /root/reference):
  * back-projection of a 2D keypoint to a world ray  snowvision/camera.py:234-253
  * closest points of two skew rays                  snowvision/triangulation.py:24-31
  * all camera-pair x person-pair candidates         snowvision/triangulation.py:50-93
  * greedy clustering + score-weighted fuse          snowvision/triangulation.py:95-162

Input convention shared by every oracle and by the CUDA path (one frame):
  kpts   (C, P, J, 2)  2D keypoints in undistorted pixels
  scores (C, P, J)     detector confidences
  counts (C,)          persons actually present for camera c (slots [0, counts[c]) valid)
  K, R   (C, 3, 3)     intrinsics; R is camera->world (camera.py:41-44)
  t      (C, 3)        camera centre in world metres
All arithmetic is float64, like the reference fed with float64 arrays.
"""
from __future__ import annotations

import numpy as np

POINTS = "hrnet_triangulate_points"
KSCORES = "hrnet_triangulate_keypoint_scores"
PSCORES = "hrnet_triangulate_person_scores"


def back_project(K, R, uv):
    """World-frame (unnormalised) ray direction of pixel ``uv``; camera.py:240-244."""
    pix = np.array([uv[0], uv[1], 1])
    in_cam = np.dot(np.linalg.inv(K), pix)
    return np.dot(R, in_cam).reshape((-1, 1))


def build_rays(kpts, counts, K, R):
    """rays[c][p][j] -> (3,1) array, for p < counts[c]; camera.py:234-253."""
    C = kpts.shape[0]
    rays = []
    for c in range(C):
        per_cam = []
        for p in range(int(counts[c])):
            per_cam.append([back_project(K[c], R[c], uv) for uv in kpts[c, p]])
        rays.append(per_cam)
    return rays


def skew_ray_solve(hm, hs, tm, ts):
    """(distance, midpoint) of the closest points of two rays; triangulation.py:24-31."""
    H = np.hstack((hm, hs))
    gram_inv = np.linalg.inv(np.dot(H.T, H))
    rhs = np.dot(H.T, (ts - tm))
    S = np.dot(gram_inv, rhs)
    Wm = hm * S[0] + tm
    Ws = -hs * S[1] + ts
    return np.linalg.norm(Wm - Ws), ((Wm + Ws) / 2).reshape(3)


def triangulate_frame(kpts, scores, counts, K, R, t, kst=0.5, ast=0.0, dthr=0.05,
                      with_index=False):
    """Human_Triangulation for one frame; triangulation.py:50-93.

    Returns the reference's 3-key dict.  ``with_index=True`` adds ``'index'``: the
    (mc, sc, pm, ps) tuple of every surviving candidate (for kernel unit tests).
    """
    kpts = np.asarray(kpts, dtype=np.float64)
    scores = np.asarray(scores, dtype=np.float64)
    C = kpts.shape[0]
    rays = build_rays(kpts, counts, K, R)
    centres = [np.asarray(t[c], dtype=np.float64).reshape(3, 1) for c in range(C)]
    pts_out, ks_out, ps_out, idx_out = [], [], [], []
    for mc in range(C - 1):
        tm = centres[mc]
        for sc in range(mc + 1, C):
            ts = centres[sc]
            for pm in range(int(counts[mc])):
                for ps in range(int(counts[sc])):
                    cand_pts, cand_sc = [], []
                    for j in range(kpts.shape[2]):
                        dist, W = skew_ray_solve(rays[mc][pm][j], rays[sc][ps][j], tm, ts)
                        sm, ss = scores[mc, pm, j], scores[sc, ps, j]
                        with np.errstate(divide="ignore", invalid="ignore"):
                            score = ((sm + ss) / 2) / (dist * 1000)
                        if sm < kst or ss < kst or dist > dthr:
                            score = 0
                        cand_pts.append(W)
                        cand_sc.append(score)
                    avg = np.mean(cand_sc)
                    if avg < ast:
                        continue
                    pts_out.append(np.array(cand_pts))
                    ks_out.append(np.array(cand_sc))
                    ps_out.append(avg)
                    idx_out.append((mc, sc, pm, ps))
    out = {POINTS: pts_out, KSCORES: ks_out, PSCORES: ps_out}
    if with_index:
        out["index"] = idx_out
    return out


def condense_frame(result, tol=0.1, num_tol=0, score_tol=0.0, center=18, keypoint_num=30):
    """Human_Triangulation_Condense; triangulation.py:95-162 (quirks Q1-Q7 of SURVEY 8a)."""
    cands = result[POINTS]
    cand_scores = result[KSCORES]
    n_cand = len(cands)
    out_pts, out_ks, out_ps = [], [], []
    absorbed = []
    for mc in range(n_cand - 1):            # the last candidate is never a main (Q1/Q2)
        if mc in absorbed:
            continue
        main_centre = cands[mc][center]
        members = [mc]
        for sc in range(mc + 1, n_cand):
            if sc in absorbed:
                continue
            if np.linalg.norm(main_centre - cands[sc][center]) > tol:   # distance to MAIN (Q3)
                continue
            absorbed.append(sc)
            members.append(sc)
        n = len(members)
        if n < num_tol:                      # members stay absorbed (Q5)
            continue
        person = np.zeros((keypoint_num, 3))
        person_ks = np.zeros(keypoint_num)
        for j in range(keypoint_num):
            w = np.array([cand_scores[m][j] for m in members])
            total = np.sum(w)
            if total == 0.0:                 # joint stays (0,0,0), score 0 (Q7)
                continue
            w = w / total
            person[j] = np.sum(np.array([cands[m][j] * w[i] for i, m in enumerate(members)]), axis=0)
            person_ks[j] = total / n         # n counts zero-score members (Q6)
        avg = np.mean(person_ks)
        if avg < score_tol:
            continue
        out_pts.append(person)
        out_ks.append(person_ks)
        out_ps.append(avg)
    return {POINTS: out_pts, KSCORES: out_ks, PSCORES: out_ps}


def fused_frame(kpts, scores, counts, K, R, t, params):
    """triangulate_frame followed by condense_frame with a params dict (main.py:62-71)."""
    tri = triangulate_frame(kpts, scores, counts, K, R, t,
                            kst=params["kst"], ast=params["ast"], dthr=params["dthr"])
    J = kpts.shape[2]
    return condense_frame(tri, tol=params["cond_tol"], num_tol=params["num_tol"],
                          score_tol=params["score_tol"], center=params["center"],
                          keypoint_num=params.get("keypoint_num", J))


# ---------------------------------------------------------------------------------------------
# Temporal smoothing (the step after the hot path)
class SecondOrderDynamic:
    """Restatement of reference snowvision/triangulation.py:4-22 (semi-implicit Euler follower)."""

    def __init__(self, f, z, r, x0):
        pi = np.pi
        self.k1 = z / (pi * f)                            # :7
        self.k2 = 1 / ((2 * pi * f) * (2 * pi * f))       # :8
        self.k3 = r * z / (2 * pi * f)                    # :9
        self.xp = x0                                      # :11
        self.y = x0                                       # :12
        self.yd = 0                                       # :13

    def update(self, T, x):
        xd = (x - self.xp) / T                            # :17
        self.xp = x                                       # :18
        self.y = self.y + T * self.yd                     # :20
        self.yd = self.yd + T * (x + self.k3 * xd - self.y - self.k1 * self.yd) / self.k2   # :21
        return self.y


def smooth_sequence(points_per_frame, f=2, z=0.75, r=0, delta_time=1 / 30):
    """Human_Triangulation_Smooth (reference snowvision/triangulation.py:164-186) applied frame after frame
    the way main.py:72-78 does.  points_per_frame: list of (n_t, J, 3) float64 arrays.  Returns the list of
    smoothed (m_t, J, 3) arrays: the first frame passes through and creates one follower per (person, joint);
    later frames zip persons with the followers of the FIRST frame by list position (persons beyond that are
    dropped, followers of absent persons are not advanced)."""
    sods, out = None, []
    for pts in points_per_frame:
        pts = [np.asarray(p, np.float64) for p in pts]
        if sods is None:                                  # :177-184 first frame
            sods = [[SecondOrderDynamic(f, z, r, p) for p in person] for person in pts]
            out.append(np.array(pts, np.float64).reshape(len(pts), -1, 3) if pts else np.zeros((0, 0, 3)))
        else:                                             # :168-176
            damped = [[sod.update(delta_time, p) for p, sod in zip(person, psods)]
                      for person, psods in zip(pts, sods)]
            out.append(np.array(damped, np.float64).reshape(len(damped), -1, 3) if damped else np.zeros((0, 0, 3)))
    return out
