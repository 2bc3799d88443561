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

import snowvision.blender as refb  # noqa: E402  (the real reference)

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


if __name__ == "__main__":
    main()
