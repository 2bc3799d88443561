"""CPU oracle of the Blender control-point stage (TEST INFRASTRUCTURE -- only ``tests/``, ``smoke()`` and the
CPU legs of bench scripts may import this; the product path never does).

A float64 NumPy restatement of the reference's ``snowvision/blender.py``:

* ``control_points``            <- ``Human_Triangulation_Blender`` (blender.py:93-143) with its helpers
  ``Get_Root_Position`` (:11-13), ``Get_Root_Rotation`` (:15-36), ``Get_Chest_IK`` (:38-40), ``Get_Chest_Pole``
  (:42-48), ``Get_Head_IK`` (:50-55), ``Get_Head_Pole`` (:57-63), ``Get_Hand_Pole`` (:65-73), ``Get_Foot_IK``
  (:74-76), ``Get_Foot_Pole`` (:78-86), ``Get_Joint_Pole`` (:88-96);
* ``rotation_matrix_to_quaternion`` <- ``snowvision/util.py:26-28`` = SciPy ``Rotation.from_matrix(R).as_quat()``.
  SciPy is a third-party dependency of the reference (``requirements.txt``: scipy==1.15.2; 1.18.1 in this image).
  Its published algorithm, restated here: a matrix whose Gram matrix is not the identity (``isclose`` with
  rtol 1e-5 / atol 1e-12) is replaced by its nearest rotation ``U @ Vt`` from the SVD (orthogonal Procrustes),
  then Markley's method picks the largest of (m00, m11, m22, trace) and normalises; the result is (x, y, z, w)
  and is NOT sign-canonicalised.  blender.py:27 reorders to (w, x, y, z);
* ``BlenderSmoothOracle``       <- ``Human_Triangulation_Blender_Smooth`` (blender.py:145-178) on top of
  ``SecondOrderDynamic`` (triangulation.py:4-22).

Pinned by ``tests/golden/blender_*.npz`` (outputs of the real reference, ``tests/golden/make_golden_blender.py``).
Layout used by the CUDA path: 24 control points in the order of ``configs/blender_armature_profile.json``, four
floats each -- (x, y, z, 0) for positions, (w, x, y, z) for ``root_rotation`` -- and one validity flag each
(the reference's 0/1 "score": 0 when any component is NaN, blender.py:135-138).
"""
import math

import numpy as np

CONTROL_POINTS = ("root_position", "root_rotation", "clavicle_r_ik", "clavicle_l_ik", "arm_r_ik", "arm_r_pole",
                  "arm_l_ik", "arm_l_pole", "leg_r_ik", "leg_r_pole", "leg_l_ik", "leg_l_pole", "hand_r_ik",
                  "hand_r_pole", "hand_l_ik", "hand_l_pole", "foot_r_ik", "foot_r_pole", "foot_l_ik", "foot_l_pole",
                  "chest_ik", "chest_pole", "head_ik", "head_pole")
ROOT_ROTATION = CONTROL_POINTS.index("root_rotation")


def _unit(v):
    return v / np.linalg.norm(v)


def rotation_matrix_to_quaternion(R):
    """SciPy's ``Rotation.from_matrix(R).as_quat()`` -> (x, y, z, w); NaN in, NaN out (SciPy raises instead)."""
    R = np.asarray(R, np.float64)
    if not np.all(np.isfinite(R)):
        return np.full(4, np.nan)
    gram = R @ R.T
    if not np.all(np.isclose(gram, np.eye(3), rtol=1e-5, atol=1e-12)):
        U, _, Vt = np.linalg.svd(R, full_matrices=False)
        R = U @ Vt
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    choice = int(np.argmax([R[0, 0], R[1, 1], R[2, 2], tr]))
    if choice == 3:
        q = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], 1 + tr])
    else:
        i, j, k = choice, (choice + 1) % 3, (choice + 2) % 3
        q = np.empty(4)
        q[i] = 1 - tr + 2 * R[i, i]
        q[j] = R[j, i] + R[i, j]
        q[k] = R[k, i] + R[i, k]
        q[3] = R[k, j] - R[j, k]
    return q / np.linalg.norm(q)


def _hand_or_foot_pole(a_point, b_point, root_point, is_left):   # blender.py:65-73, 78-86
    a, b = a_point - root_point, b_point - root_point
    n = np.cross(b, a) if is_left else np.cross(a, b)
    return root_point + n / np.linalg.norm(n)


def _joint_pole(joint, upper, lower):   # blender.py:88-96
    a, b, c = upper - joint, lower - joint, upper - lower
    n = np.cross(np.cross(b, a), c)
    return joint + n / np.linalg.norm(n)


def control_points(person):
    """One person (J>=130, 3) -> (ctrl (24,4) float64, valid (24,) bool)."""
    p = np.asarray(person, np.float64)
    if p.shape[0] < 130:
        raise IndexError("Human_Triangulation_Blender needs the 133 Wholebody keypoints")
    with np.errstate(all="ignore"):
        pelvis_mid, shoulder_mid, ear_mid = (p[11] + p[12]) / 2, (p[5] + p[6]) / 2, (p[3] + p[4]) / 2
        x = _unit(p[11] - p[12])
        y = _unit(shoulder_mid - pelvis_mid)
        z = _unit(np.cross(x, y))
        q = rotation_matrix_to_quaternion(np.array([x, y, z]).T)
        spine, neck = shoulder_mid - pelvis_mid, ear_mid - shoulder_mid
        vals = {
            "root_position": pelvis_mid,
            "root_rotation": np.array([q[3], q[0], q[1], q[2]]),
            "clavicle_r_ik": p[6], "clavicle_l_ik": p[5],
            "arm_r_ik": p[10], "arm_r_pole": _joint_pole(p[8], p[6], p[10]),
            "arm_l_ik": p[9], "arm_l_pole": _joint_pole(p[7], p[5], p[9]),
            "leg_r_ik": p[16], "leg_r_pole": _joint_pole(p[14], p[12], p[16]),
            "leg_l_ik": p[15], "leg_l_pole": _joint_pole(p[13], p[11], p[15]),
            "hand_r_ik": p[121], "hand_r_pole": _hand_or_foot_pole(p[117], p[129], p[112], False),
            "hand_l_ik": p[100], "hand_l_pole": _hand_or_foot_pole(p[96], p[108], p[91], True),
            "foot_r_ik": (p[20] + p[21]) / 2, "foot_r_pole": _hand_or_foot_pole(p[20], p[21], p[22], False),
            "foot_l_ik": (p[17] + p[18]) / 2, "foot_l_pole": _hand_or_foot_pole(p[17], p[18], p[19], True),
            "chest_ik": shoulder_mid,
            "chest_pole": shoulder_mid + _unit(np.cross(p[5] - p[6], spine)),
            "head_ik": shoulder_mid + _unit(neck),
            "head_pole": ear_mid + _unit(np.cross(p[3] - p[4], neck)),
        }
    ctrl = np.zeros((24, 4))
    valid = np.zeros(24, bool)
    for i, name in enumerate(CONTROL_POINTS):
        v = vals[name]
        ctrl[i, :len(v)] = v
        valid[i] = not np.any(np.isnan(v))
    return ctrl, valid


def control_points_batch(points):
    """(N, J, >=3) -> ctrl (N,24,4), valid (N,24)."""
    pts = np.asarray(points, np.float64)
    ctrl = np.zeros((pts.shape[0], 24, 4))
    valid = np.zeros((pts.shape[0], 24), bool)
    for n in range(pts.shape[0]):
        ctrl[n], valid[n] = control_points(pts[n, :, :3])
    return ctrl, valid


class BlenderSmoothOracle:
    """``Human_Triangulation_Blender_Smooth`` frame after frame (blender.py:145-178).  ``fzr`` (24,3): the
    (f, z, r) of every control point (``configs/blender_smooth_profile.json``)."""

    def __init__(self, fzr):
        fzr = np.asarray(fzr, np.float64).reshape(24, 3)
        f, z, r = fzr[:, 0:1], fzr[:, 1:2], fzr[:, 2:3]
        self.k1 = z / (math.pi * f)                                # triangulation.py:7-9
        self.k2 = 1 / ((2 * math.pi * f) * (2 * math.pi * f))
        self.k3 = r * z / (2 * math.pi * f)
        self.state = None

    def step(self, ctrl, valid, T):
        """ctrl (n,24,4), valid (n,24) of one frame -> the smoothed (m,24,4) of that frame."""
        ctrl = np.asarray(ctrl, np.float64)
        valid = np.asarray(valid, bool)
        if self.state is None:                                     # blender.py:165-176
            x0 = np.where(valid[..., None], ctrl, 0.0)
            self.state = {"xp": x0.copy(), "y": x0.copy(), "yd": np.zeros_like(x0)}
            return ctrl.copy()
        s = self.state
        m = min(ctrl.shape[0], s["xp"].shape[0])                   # zip() of blender.py:151-153
        out = np.empty((m, 24, 4))
        for k in range(m):
            x = np.where(valid[k][:, None], ctrl[k], s["xp"][k])   # blender.py:157-160
            xd = (x - s["xp"][k]) / T                              # triangulation.py:16-22
            s["xp"][k] = x
            s["y"][k] = s["y"][k] + T * s["yd"][k]
            s["yd"][k] = s["yd"][k] + T * (x + self.k3 * xd - s["y"][k] - self.k1 * s["yd"][k]) / self.k2
            out[k] = s["y"][k]
        return out
