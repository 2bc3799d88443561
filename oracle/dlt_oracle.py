"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the optional DLT mode.

The reference (snowvision) has no DLT; BASELINE.json's north_star names it, so this mode's oracle is the
textbook method itself: build the 2V x 4 matrix of the normalised homogeneous equations and take the right
singular vector of the smallest singular value (``np.linalg.svd``), float64.  Parity status: this restates a
published algorithm (Hartley & Zisserman, Multiple View Geometry, section 12.2), not reference code; it is
cross-checked against the reference-pinned midpoint oracle only in the sense that both recover the synthetic truth.
"""
from __future__ import annotations

import numpy as np


def dlt_points(kpts, scores, K, R, t, kst=0.5):
    """kpts (F,C,1,J,2) f32, scores (F,C,1,J) f32 -> (F,J,4) f64: x, y, z, views used (zeros for < 2 views)."""
    F, C, _, J, _ = kpts.shape
    Kinv = np.linalg.inv(np.asarray(K, np.float64))
    Rm = np.asarray(R, np.float64)
    Q = np.concatenate([Rm.transpose(0, 2, 1), -np.einsum("cji,cj->ci", Rm, np.asarray(t, np.float64))[:, :, None]], axis=2)
    kf = np.float32(kst) if np.float64(np.float32(kst)) >= kst else np.nextafter(np.float32(kst), np.float32(np.inf))
    out = np.zeros((F, J, 4))
    uv1 = np.concatenate([kpts[:, :, 0].astype(np.float64), np.ones((F, C, J, 1))], axis=-1)       # (F,C,J,3)
    n = np.einsum("cij,fcpj->fcpi", Kinv, uv1)                                                       # (F,C,J,3)
    xn, yn = n[..., 0] / n[..., 2], n[..., 1] / n[..., 2]
    rows_x = xn[..., None] * Q[None, :, None, 2, :] - Q[None, :, None, 0, :]                          # (F,C,J,4)
    rows_y = yn[..., None] * Q[None, :, None, 2, :] - Q[None, :, None, 1, :]
    ok = scores[:, :, 0] >= kf                                                                       # (F,C,J)
    for f in range(F):
        for j in range(J):
            use = ok[f, :, j]
            v = int(use.sum())
            out[f, j, 3] = v
            if v < 2:
                continue
            A = np.concatenate([rows_x[f, use, j], rows_y[f, use, j]], axis=0)
            _, _, vt = np.linalg.svd(A)
            x = vt[-1]
            out[f, j, :3] = x[:3] / x[3]
    return out
