"""Golden vectors of the REAL reference at the full multi-person sizes of BASELINE configs[2] and [3] (SURVEY 8d:
>= 2 true frames of cfg3 at J=133, 1 frame of cfg4).  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_big.py

The candidate lists of these sizes are large (448 / 7 680 candidates x 133 joints), so only the condensed output of
``Human_Triangulation_Condense`` (what main.py:66-71 hands on), the candidate count and the candidates' person scores
are stored, next to the float32 inputs.  Fixtures: ``tests/golden/big_*.npz``.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import ref_runner  # noqa: E402
from snowmocap_b200 import synth  # noqa: E402


def save(name, rig, data, prm):
    F, C, P, J = data["scores"].shape
    res, dt = ref_runner.run_frames(rig.K, rig.R, rig.t, data["kpts"], data["scores"], data["counts"], prm, keep_tri=True)
    d = {}
    for f, (tri, con) in enumerate(res):
        n = len(con["hrnet_triangulate_points"])
        d[f"con_pts_{f}"] = np.array(con["hrnet_triangulate_points"], np.float64).reshape(n, J, 3)
        d[f"con_ks_{f}"] = np.array(con["hrnet_triangulate_keypoint_scores"], np.float64).reshape(n, J)
        d[f"con_ps_{f}"] = np.array(con["hrnet_triangulate_person_scores"], np.float64).reshape(n)
        d[f"tri_ps_{f}"] = np.array(tri["hrnet_triangulate_person_scores"], np.float64)
    p = dict(prm, keypoint_num=J)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), K=rig.K, R=rig.R, t=rig.t, kpts=data["kpts"],
                        scores=data["scores"], counts=data["counts"], params=json.dumps(p),
                        seconds_per_frame=dt / F, **d)
    print(f"{name}: frames={F} {dt / F:.2f} s/frame candidates={[len(d[f'tri_ps_{f}']) for f in range(F)]} "
          f"persons={[d[f'con_ps_{f}'].shape[0] for f in range(F)]}")


def main():
    assert ref_runner.make_ref(), "needs /root/reference"
    M = synth.MULTI_PARAMS
    ring8 = synth.ring_rig(8)
    save("big_cfg3_c8p4j133", ring8, synth.make_frames(ring8, 2, 4, 133, seed=31, low_score_frac=0.05), M)
    ring16 = synth.ring_rig(16)
    save("big_cfg4_c16p8j133", ring16, synth.make_frames(ring16, 1, 8, 133, seed=32), M)


if __name__ == "__main__":
    main()
