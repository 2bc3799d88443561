"""Drop-in mirror of the hot-path part of ``snowvision.camera`` (reference camera.py:16-59,
141-170, 234-261): the camera parameter containers and the per-frame observation store.

Capture, recording, chessboard calibration and undistortion (reference camera.py:61-139,
172-218) are out of scope of this engine and are not provided.

Unlike the reference, ``add_human_2D_points`` does not back-project on the host: it only
records the detector output; rays are computed on the GPU inside ``Human_Triangulation``.
"""
from __future__ import annotations

import json

import numpy as np


_warned_cast = [False]


class Camera:
    """Parameter container: K (3,3), R (3,3) camera->world, t (3,1) camera centre, D (1,5)."""

    def __init__(self, cap_id=0, frame_width=1280, frame_height=720, camera_info_path=None, camera_info_dict=None):
        self.cap_id = cap_id
        self.frame_width = frame_width
        self.frame_height = frame_height
        self.K = np.zeros((3, 3))
        self.R = np.eye(3)
        self.t = np.zeros((3, 1))
        self.D = np.zeros((1, 5))
        self._reset_observations()
        if camera_info_path is not None or camera_info_dict is not None:
            if camera_info_path is not None:
                with open(camera_info_path, "r") as fh:
                    camera_info_dict = json.load(fh)
            self.cap_id = camera_info_dict["cap_id"]
            self.frame_width = camera_info_dict["frame_width"]
            self.frame_height = camera_info_dict["frame_height"]
            self.K = np.array(camera_info_dict["K"])
            self.R = np.array(camera_info_dict["R"])
            self.t = np.array(camera_info_dict["t"])
            self.D = np.array(camera_info_dict["D"])

    def _reset_observations(self):
        # same attribute names as the reference; rays live on the device, so that list stays empty
        self.points = []
        self.point_rays = []
        self.hrnet_points = []
        self.hrnet_point_rays = []
        self.hrnet_point_score = []

    def camera_info_dict(self):
        return {"cap_id": self.cap_id, "frame_width": self.frame_width, "frame_height": self.frame_height,
                "K": np.asarray(self.K).tolist(), "R": np.asarray(self.R).tolist(),
                "t": np.asarray(self.t).tolist(), "D": np.asarray(self.D).tolist()}

    def save_camera_info(self, camera_info_path):
        with open(camera_info_path, "w") as fh:
            fh.write(json.dumps(self.camera_info_dict()))


class CameraGroup:
    """Same constructor and attributes as the reference (camera.py:141-157)."""

    def __init__(self, cap_ids=[0, 1], resolutions=[(1280, 720), (1280, 720)], camera_group_info_path=None):
        self.cameras = []
        self._engine = None
        self._engine_key = None
        if camera_group_info_path is None:
            self.camera_num = len(cap_ids)
            for cap_id, resolution in zip(cap_ids, resolutions):
                self.cameras.append(Camera(cap_id=cap_id, frame_width=resolution[0], frame_height=resolution[1]))
        else:
            with open(camera_group_info_path, "r") as fh:
                info = json.load(fh)
            self.camera_num = info["camera_num"]
            for camera_info_dict in info["camera_group_info"]:
                self.cameras.append(Camera(camera_info_dict=camera_info_dict))

    def camera_group_info_dict(self):
        return {"camera_num": self.camera_num,
                "camera_group_info": [self.cameras[i].camera_info_dict() for i in range(self.camera_num)]}

    def save_camera_group_info(self, camera_group_info_path):
        with open(camera_group_info_path, "w") as fh:
            fh.write(json.dumps(self.camera_group_info_dict(), indent=4))

    # -- per-frame observation store (reference camera.py:234-261) -----------------------------
    def add_human_2D_points(self, person, scores, camera_index, ax=None):
        """Record one detected person (J,2) with per-keypoint scores (J,) for ``camera_index``.

        Input contract: keypoints and scores are stored as float32 (what pose detectors emit; the engine's input
        layout).  The reference keeps whatever dtype it is handed and back-projects in float64, so float64 keypoints
        lose ~1e-7 relative pixel precision here and a score within one float32 ulp of ``keypoint_score_threshold``
        can fall on the other side of it.  A lossy cast warns once."""
        person = np.asarray(person)
        if person.dtype == np.float64 and not _warned_cast[0] and person.size and \
                not np.array_equal(person.astype(np.float32).astype(np.float64), person):
            import warnings
            warnings.warn("snowmocap_b200: float64 keypoints are rounded to float32 (the engine's input precision); "
                          "see CameraGroup.add_human_2D_points", stacklevel=2)
            _warned_cast[0] = True
        cam = self.cameras[camera_index]
        cam.hrnet_points.append(np.asarray(person, dtype=np.float32).reshape(-1, 2))
        cam.hrnet_point_score.append(np.asarray(scores, dtype=np.float32).reshape(-1))

    def clear_2D_points(self):
        for cam in self.cameras:
            cam._reset_observations()

    # -- device side ---------------------------------------------------------------------------
    def parameters(self):
        """(K (C,3,3), R (C,3,3), t (C,3)) float64 of the first camera_num cameras."""
        cams = self.cameras[:self.camera_num]
        K = np.stack([np.asarray(c.K, np.float64).reshape(3, 3) for c in cams])
        R = np.stack([np.asarray(c.R, np.float64).reshape(3, 3) for c in cams])
        t = np.stack([np.asarray(c.t, np.float64).reshape(3) for c in cams])
        return K, R, t

    def engine(self, device=0):
        """TriangulationEngine for the current camera parameters (rebuilt if they changed)."""
        from .engine import TriangulationEngine
        K, R, t = self.parameters()
        key = (K.tobytes(), R.tobytes(), t.tobytes(), device)
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            self._engine = TriangulationEngine(K, R, t, device=device)
            self._engine_key = key
        return self._engine

    def pack_frame(self):
        """Dense (1,C,P,J,2) / (1,C,P,J) / (1,C) arrays of the observations added since the last clear."""
        cams = self.cameras[:self.camera_num]
        P = max((len(c.hrnet_points) for c in cams), default=0)
        Js = {p.shape[0] for c in cams for p in c.hrnet_points}
        if len(Js) > 1:
            raise ValueError(f"all persons must have the same number of keypoints, got {sorted(Js)}")
        J = Js.pop() if Js else 0
        kpts = np.zeros((1, len(cams), P, J, 2), np.float32)
        scores = np.zeros((1, len(cams), P, J), np.float32)
        counts = np.zeros((1, len(cams)), np.int32)
        for c, cam in enumerate(cams):
            counts[0, c] = len(cam.hrnet_points)
            for p, (pts, sc) in enumerate(zip(cam.hrnet_points, cam.hrnet_point_score)):
                if sc.shape[0] != J:
                    raise ValueError("scores and keypoints disagree on the number of keypoints")
                kpts[0, c, p] = pts
                scores[0, c, p] = sc
        return kpts, scores, counts
