"""Generate golden vectors by running the REAL reference (``/root/reference/snowvision``).

Run in the build container only (the reference does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every fixture ``tests/golden/<case>.npz`` holds the inputs (float32 keypoints/scores,
counts, camera parameters, thresholds) and, per frame, the reference's own outputs of
``Human_Triangulation`` (``tri_*``) and ``Human_Triangulation_Condense`` (``con_*``)
obtained through the exact ``main.py:55-71`` call sequence.  The reference is fed
float64 arrays holding the float32 values, so its arithmetic is float64 under any NumPy.
``floor_rig.npz`` is the 4-camera calibration of ``configs/camera_group_floor.json``.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SNOW_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import snowvision as ref  # noqa: E402  (the real reference)
from snowmocap_b200 import synth  # noqa: E402


def ref_group(rig):
    C = rig.C
    group = ref.CameraGroup(cap_ids=list(range(C)), resolutions=[(1280, 720)] * C)
    for c in range(C):
        group.cameras[c].K = rig.K[c].copy()
        group.cameras[c].R = rig.R[c].copy()
        group.cameras[c].t = rig.t[c].reshape(3, 1).copy()
    return group


def run_reference(rig, data, params, keypoint_num=None):
    """main.py:55-71 sequence per frame -> lists of (tri dict, condensed dict)."""
    group = ref_group(rig)
    F, C, P, J = data["scores"].shape
    out = []
    for f in range(F):
        for c in range(C):
            for p in range(int(data["counts"][f, c])):
                group.add_human_2D_points(data["kpts"][f, c, p].astype(np.float64),
                                          data["scores"][f, c, p].astype(np.float64), c)
        with np.errstate(all="ignore"):
            tri = ref.Human_Triangulation(group, keypoint_score_threshold=params["kst"],
                                          average_score_threshold=params["ast"],
                                          distance_threshold=params["dthr"])
            con = ref.Human_Triangulation_Condense(
                tri, condense_distance_tol=params["cond_tol"],
                condense_person_num_tol=params["num_tol"], condense_score_tol=params["score_tol"],
                center_point_index=params["center"],
                keypoint_num=J if keypoint_num is None else keypoint_num)
        group.clear_2D_points()
        out.append((tri, con))
    return out


def pack(res, J, jout):
    d = {}
    for f, (tri, con) in enumerate(res):
        for tag, r, j in (("tri", tri, J), ("con", con, jout)):
            n = len(r["hrnet_triangulate_points"])
            d[f"{tag}_pts_{f}"] = np.array(r["hrnet_triangulate_points"], np.float64).reshape(n, j, 3)
            d[f"{tag}_ks_{f}"] = np.array(r["hrnet_triangulate_keypoint_scores"], np.float64).reshape(n, j)
            d[f"{tag}_ps_{f}"] = np.array(r["hrnet_triangulate_person_scores"], np.float64).reshape(n)
    return d


def save_case(name, rig, data, params, keypoint_num=None):
    J = data["scores"].shape[3]
    jout = J if keypoint_num is None else keypoint_num
    res = run_reference(rig, data, params, keypoint_num)
    p = dict(params)
    p["keypoint_num"] = jout
    np.savez_compressed(os.path.join(HERE, name + ".npz"), K=rig.K, R=rig.R, t=rig.t,
                        kpts=data["kpts"], scores=data["scores"], counts=data["counts"],
                        params=json.dumps(p), **pack(res, J, jout))
    ncand = [len(r[0]["hrnet_triangulate_points"]) for r in res]
    nout = [len(r[1]["hrnet_triangulate_points"]) for r in res]
    print(f"{name}: frames={len(res)} candidates={ncand} persons={nout}")


def floor_rig():
    with open(os.path.join(REF, "configs", "camera_group_floor.json")) as fh:
        info = json.load(fh)["camera_group_info"]
    return synth.Rig([c["K"] for c in info], [c["R"] for c in info],
                     [np.array(c["t"]).reshape(3) for c in info])


def save_smooth_case(name, counts, J, f, z, r, dt, seed):
    """Human_Triangulation_Smooth of the real reference, frame after frame like main.py:72-78, on a synthetic
    walk of persons with a varying person count per frame."""
    rng = np.random.default_rng(seed)
    F, P = len(counts), max(max(counts), 1)
    base = rng.uniform(-2, 2, (P, 1, 3)) + rng.uniform(-0.4, 0.4, (P, J, 3))
    walk = np.cumsum(rng.normal(0, 0.01, (F, P, J, 3)), axis=0)
    pts = (base[None] + walk + rng.normal(0, 0.004, (F, P, J, 3))).astype(np.float32).astype(np.float64)
    prev, d = None, {}
    for t in range(F):
        n = counts[t]
        res = {"hrnet_triangulate_points": [pts[t, k].copy() for k in range(n)],
               "hrnet_triangulate_keypoint_scores": [np.ones(J) for _ in range(n)],
               "hrnet_triangulate_person_scores": [1.0] * n}
        res = ref.Human_Triangulation_Smooth(res, prev, f=f, z=z, r=r, delta_time=dt)
        prev = res
        out = res["hrnet_triangulate_points"]
        d[f"out_{t}"] = np.array(out, np.float64).reshape(len(out), J, 3)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pts=pts, counts=np.array(counts, np.int32),
                        params=json.dumps(dict(f=f, z=z, r=r, dt=dt)), **d)
    print(f"{name}: frames={F} persons in={counts} out={[d[f'out_{t}'].shape[0] for t in range(F)]}")


def main():
    floor = floor_rig()
    np.savez(os.path.join(HERE, "floor_rig.npz"), K=floor.K, R=floor.R, t=floor.t)
    D, M = synth.DEFAULT_PARAMS, synth.MULTI_PARAMS

    # BASELINE.json configs[0]: 2 cameras, 1 person, 17 keypoints (Condense emits 0 persons, Q1)
    r = floor.subset(2)
    save_case("cfg1_c2p1j17", r, synth.make_frames(r, 6, 1, 17, seed=11), D)
    # BASELINE.json configs[1]: 4 cameras, 1 person, 133 keypoints, shipped thresholds
    save_case("cfg2_c4p1j133", floor, synth.make_frames(floor, 3, 1, 133, seed=12), D)
    save_case("cfg2_lowscore", floor,
              synth.make_frames(floor, 2, 1, 133, seed=13, low_score_frac=0.15), D)
    # multi-person matching on the demo calibration
    save_case("multi_c4p3j133", floor,
              synth.make_frames(floor, 1, 3, 133, seed=14, low_score_frac=0.1), M)
    # BASELINE.json configs[2] geometry at J=17 to keep the fixture small
    ring8 = synth.ring_rig(8)
    save_case("multi_c8p4j17", ring8, synth.make_frames(ring8, 2, 4, 17, seed=15), M)
    # ragged person counts (a camera may see 0 persons)
    ring5 = synth.ring_rig(5)
    save_case("ragged_c5p3j17", ring5,
              synth.make_frames(ring5, 5, 3, 17, seed=16, low_score_frac=0.1, drop_prob=0.35), M)
    # quirk exercises: everything kept (ast=0) so ghosts cluster; cluster-size and score filters;
    # non-default centre joint; keypoint_num < J truncation
    ring4 = synth.ring_rig(4)
    dq = synth.make_frames(ring4, 2, 3, 17, seed=17, low_score_frac=0.2)
    save_case("quirk_allkept", ring4, dq, dict(M, ast=0.0))
    save_case("quirk_numtol", ring4, dq, dict(M, ast=0.0, num_tol=3, center=5))
    save_case("quirk_scoretol", ring4, dq, dict(M, ast=0.0, score_tol=0.08, cond_tol=0.5))
    save_case("quirk_truncate", ring4, dq, dict(M, ast=0.1, center=3), keypoint_num=11)
    save_case("quirk_bigtol", ring4, dq, dict(M, ast=0.05, cond_tol=10.0))
    # temporal smoothing: shipped parameters (configs/snowmocap_default_config.json:18-21),
    # a person count that grows, shrinks and hits zero; non-zero r; a clip whose first frame is empty
    save_smooth_case("smooth_default", [2, 2, 3, 3, 1, 2, 0, 2, 2, 1, 3, 2] + [2] * 28, 17, 2.5, 0.75, 0, 0.03333333333, 21)
    save_smooth_case("smooth_r", [1] * 30, 133, 2.0, 0.75, 0.5, 1 / 30, 22)
    save_smooth_case("smooth_emptyfirst", [0, 2, 2, 1, 2], 17, 2.5, 0.75, 0, 1 / 30, 23)


if __name__ == "__main__":
    main()
