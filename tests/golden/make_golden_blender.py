"""Golden vectors of the Blender control-point stage from the REAL reference (``/root/reference/snowvision``).

Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_blender.py

``blender_points.npz``  persons (N,133,3) -> ``Human_Triangulation_Blender`` (blender.py:93-143) control points
                        (N,24,4) and 0/1 scores (N,24), with the shipped ``configs/blender_armature_profile.json``;
                        some persons have joints zeroed (Condense's "no observation" value) so poles go NaN.
``blender_smooth.npz``  a clip run frame after frame through ``Human_Triangulation_Blender`` +
                        ``Human_Triangulation_Blender_Smooth`` + ``Human_Triangulation_To_Blender_Result``
                        (main.py:80-87) with the shipped ``configs/blender_smooth_profile.json``; person counts
                        vary and some control points are invalid on some frames (also on the first).
``pipeline_main.npz``   the whole per-frame body of the reference's ``main.py`` (:50-87) on 10 synthetic frames of
                        the shipped 4-camera calibration with the shipped thresholds: add_human_2D_points ->
                        Human_Triangulation -> Condense -> Smooth -> Blender -> Blender_Smooth ->
                        To_Blender_Result; inputs (float32 2D keypoints) and the final per-frame control points.
``blender_profiles.npz`` the two shipped profiles (names in order; f, z, r per control point).
Positions are stored as (x, y, z, 0), ``root_rotation`` as (w, x, y, z) -- the layout of the CUDA path.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SNOW_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import snowvision as ref  # noqa: E402  (the real reference)
import snowvision.blender as refb  # noqa: E402
from snowmocap_b200 import synth  # noqa: E402

with open(os.path.join(REF, "configs", "blender_armature_profile.json")) as fh:
    ARMATURE = json.load(fh)
with open(os.path.join(REF, "configs", "blender_smooth_profile.json")) as fh:
    SMOOTH = json.load(fh)
NAMES = list(ARMATURE.keys())


def body(rng, n):
    """n roughly human-shaped persons: joints = centre + U([-0.4,0.4]^2 x [0,1.8]), as float32 values."""
    c = np.concatenate([rng.uniform(-2, 2, (n, 1, 2)), np.zeros((n, 1, 1))], axis=2)
    j = np.concatenate([rng.uniform(-0.4, 0.4, (n, 133, 2)), rng.uniform(0, 1.8, (n, 133, 1))], axis=2)
    return (c + j).astype(np.float32).astype(np.float64)


def pack(cp_list, sc_list):
    ctrl = np.zeros((len(cp_list), 24, 4))
    valid = np.zeros((len(sc_list), 24), np.uint8)
    for n, cp in enumerate(cp_list):
        for i, name in enumerate(NAMES):
            v = np.asarray(cp[name], np.float64)
            ctrl[n, i, :len(v)] = v
    for n, sc in enumerate(sc_list):
        for i, name in enumerate(NAMES):
            valid[n, i] = sc[name]
    return ctrl, valid


def run_blender(persons):
    res = {"hrnet_triangulate_points": [p for p in persons],
           "hrnet_triangulate_keypoint_scores": [np.ones(133) for _ in persons]}
    with np.errstate(all="ignore"):
        return refb.Human_Triangulation_Blender(res, ARMATURE)


def main():
    rng = np.random.default_rng(31)
    np.savez(os.path.join(HERE, "blender_profiles.npz"), names=np.array(NAMES),
             fzr=np.array([SMOOTH[n] for n in NAMES], np.float64))

    pts = body(rng, 24)
    # zeroed joints (Condense leaves (0,0,0) where nothing was observed, triangulation.py:136-143): a zero hand
    # makes the hand pole NaN, coincident toes make the foot pole NaN; the root joints stay intact because the
    # reference raises LinAlgError on a NaN root rotation
    pts[3, [112, 117, 129]] = 0.0
    pts[5, [91, 96, 108]] = 0.0
    pts[7, 20] = pts[7, 21]
    pts[9, [8, 6, 10]] = 0.0
    pts[11, 3] = pts[11, 4]
    out = run_blender(pts)
    ctrl, valid = pack(out["blender_armature_control_points"], out["blender_armature_control_points_scores"])
    np.savez_compressed(os.path.join(HERE, "blender_points.npz"), pts=pts, ctrl=ctrl, valid=valid)
    print("blender_points: persons", len(pts), "invalid control points", int((valid == 0).sum()))

    # clip: main.py:80-87 frame after frame
    counts = [2, 2, 3, 1, 2, 0, 2, 2, 3, 2] + [2] * 30
    F, P = len(counts), 3
    base = body(rng, P)
    walk = np.cumsum(rng.normal(0, 0.01, (F, P, 133, 3)), axis=0)
    clip = (base[None] + walk).astype(np.float32).astype(np.float64)
    clip[0, 1, [112, 117, 129]] = 0.0          # invalid on the FIRST frame: follower starts from zeros
    clip[4:7, 0, [91, 96, 108]] = 0.0          # invalid for a few frames: follower holds its last input
    clip[12, 1, 20] = clip[12, 1, 21]
    prev, d = None, {}
    for t in range(F):
        cur = run_blender([clip[t, k] for k in range(counts[t])])
        sm = refb.Human_Triangulation_Blender_Smooth(cur, ARMATURE, SMOOTH, prev, delta_time=0.03333333333)
        prev = sm
        fin = refb.Human_Triangulation_To_Blender_Result(sm)
        c, v = pack(fin["armature"], fin["score"])
        d[f"ctrl_{t}"], d[f"valid_{t}"] = c, v
    np.savez_compressed(os.path.join(HERE, "blender_smooth.npz"), clip=clip, counts=np.array(counts, np.int32),
                        dt=0.03333333333, **d)
    print("blender_smooth: frames", F, "persons out", [d[f"ctrl_{t}"].shape[0] for t in range(F)])


def pipeline():
    """main.py:50-87 of the reference, frame after frame."""
    with open(os.path.join(REF, "configs", "snowmocap_default_config.json")) as fh:
        cfg = json.load(fh)
    z = np.load(os.path.join(HERE, "floor_rig.npz"))
    rig = synth.Rig(z["K"], z["R"], z["t"])
    F = 10
    data = synth.make_frames(rig, F, 1, 133, seed=41)
    group = ref.CameraGroup(cap_ids=list(range(rig.C)), resolutions=[(1280, 720)] * rig.C)
    for c in range(rig.C):
        group.cameras[c].K, group.cameras[c].R = rig.K[c].copy(), rig.R[c].copy()
        group.cameras[c].t = rig.t[c].reshape(3, 1).copy()
    kst = 0.5   # the shipped file says 3.0, which rejects every rtmlib score <= 1; SURVEY 8d uses 0.5
    prev_tri, prev_bl, d = None, None, {}
    for f in range(F):
        for c in range(rig.C):
            group.add_human_2D_points(data["kpts"][f, c, 0].astype(np.float64), data["scores"][f, c, 0].astype(np.float64), c)
        with np.errstate(all="ignore"):
            tri = ref.Human_Triangulation(group, keypoint_score_threshold=kst,
                                          average_score_threshold=cfg["average_score_threshold"],
                                          distance_threshold=cfg["distance_threshold"])
            tri = ref.Human_Triangulation_Condense(tri, condense_distance_tol=cfg["condense_distance_tol"],
                                                   condense_person_num_tol=cfg["condense_person_num_tol"],
                                                   condense_score_tol=cfg["condense_score_tol"],
                                                   center_point_index=cfg["center_point_index"],
                                                   keypoint_num=cfg["keypoint_num"])
            tri = ref.Human_Triangulation_Smooth(tri, prev_tri, f=cfg["smooth_f"], z=cfg["smooth_z"], r=cfg["smooth_r"],
                                                 delta_time=cfg["smooth_delta_time"])
            prev_tri = tri
            bl = refb.Human_Triangulation_Blender(tri, ARMATURE)
            bl = refb.Human_Triangulation_Blender_Smooth(bl, ARMATURE, SMOOTH, prev_bl, delta_time=cfg["smooth_delta_time"])
            prev_bl = bl
            fin = refb.Human_Triangulation_To_Blender_Result(bl)
        group.clear_2D_points()
        d[f"ctrl_{f}"], d[f"valid_{f}"] = pack(fin["armature"], fin["score"])
        d[f"joints_{f}"] = np.array(tri["hrnet_triangulate_points"], np.float64).reshape(-1, 133, 3)
    params = dict(kst=kst, ast=cfg["average_score_threshold"], dthr=cfg["distance_threshold"],
                  cond_tol=cfg["condense_distance_tol"], num_tol=cfg["condense_person_num_tol"],
                  score_tol=cfg["condense_score_tol"], center=cfg["center_point_index"], keypoint_num=cfg["keypoint_num"],
                  smooth_f=cfg["smooth_f"], smooth_z=cfg["smooth_z"], smooth_r=cfg["smooth_r"],
                  smooth_delta_time=cfg["smooth_delta_time"])
    np.savez_compressed(os.path.join(HERE, "pipeline_main.npz"), kpts=data["kpts"], scores=data["scores"],
                        params=json.dumps(params), **d)
    print("pipeline_main: frames", F, "persons", [d[f"ctrl_{f}"].shape[0] for f in range(F)])


if __name__ == "__main__":
    pipeline()
    main()
