"""Synthetic camera rigs and 2D observations for tests and benchmarks.

Nothing here is on the hot path: it only manufactures inputs with the same shape and
statistics the reference's ``main.py:50-55`` loop feeds to ``add_human_2D_points``.

Conventions (same as the reference, ``snowvision/camera.py:41-44``):
  K (C,3,3) intrinsics, R (C,3,3) camera->world rotation, t (C,3) camera centre (metres).
2D inputs are always float32 and never noise-free: the reference divides by the ray
distance (``snowvision/triangulation.py:72``), so exact projections give inf/NaN.
"""
from __future__ import annotations

import numpy as np

CHUNK = 256  # frames per independently-seeded RNG chunk (lets shards generate their own slice)


class Rig:
    """Plain container of camera parameters: K (C,3,3), R (C,3,3), t (C,3), float64."""

    def __init__(self, K, R, t):
        self.K = np.ascontiguousarray(K, dtype=np.float64).reshape(-1, 3, 3)
        self.R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 3, 3)
        self.t = np.ascontiguousarray(t, dtype=np.float64).reshape(-1, 3)
        assert self.K.shape[0] == self.R.shape[0] == self.t.shape[0]

    @property
    def C(self):
        return self.K.shape[0]

    def subset(self, n):
        return Rig(self.K[:n], self.R[:n], self.t[:n])


def ring_rig(C, radius=4.5, height=2.6, seed=0):
    """C cameras on a ring looking at (0,0,1); the rig SURVEY.md 8(d) specifies for C >= 8."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    K = np.tile(np.array([[690.0, 0, 640.0], [0, 695.0, 360.0], [0, 0, 1.0]]), (C, 1, 1))
    R = np.zeros((C, 3, 3))
    t = np.zeros((C, 3))
    for i in range(C):
        az = 2 * np.pi * (i + 0.5) / C
        t[i] = [radius * np.cos(az), radius * np.sin(az), height + 0.1 * rng.standard_normal()]
        z = np.array([0.0, 0.0, 1.0]) - t[i]
        z /= np.linalg.norm(z)
        x = np.cross(z, [0.0, 0.0, 1.0])
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R[i] = np.stack([x, y, z], axis=1)
    return Rig(K, R, t)


def project(rig, X):
    """Pinhole projection of world points X (..., 3) into every camera -> (C, ..., 2)."""
    Xc = np.einsum("cji,c...j->c...i", rig.R, X[None] - rig.t.reshape((rig.C,) + (1,) * (X.ndim - 1) + (3,)))
    uvw = np.einsum("cij,c...j->c...i", rig.K, Xc)
    return uvw[..., :2] / uvw[..., 2:3]


def make_frames(rig, F, P, J, seed=1234, frame0=0, noise_px=0.5, low_score_frac=0.0,
                drop_prob=0.0, shuffle=True, kst=0.5):
    """Synthetic observations for frames [frame0, frame0+F).

    Returns dict(kpts (F,C,P,J,2) f32, scores (F,C,P,J) f32, counts (F,C) i32,
    truth (F,P,J,3) f64).  Person slots >= counts[f,c] hold zeros.
    """
    C = rig.C
    kpts = np.zeros((F, C, P, J, 2), np.float32)
    scores = np.zeros((F, C, P, J), np.float32)
    counts = np.zeros((F, C), np.int32)
    truth = np.zeros((F, P, J, 3), np.float64)
    f = 0
    while f < F:
        g = frame0 + f
        chunk, off = divmod(g, CHUNK)
        n = min(CHUNK - off, F - f)
        kc, sc, cc, tc = _make_chunk(rig, chunk, P, J, seed, noise_px, low_score_frac,
                                     drop_prob, shuffle, kst)
        kpts[f:f + n], scores[f:f + n] = kc[off:off + n], sc[off:off + n]
        counts[f:f + n], truth[f:f + n] = cc[off:off + n], tc[off:off + n]
        f += n
    return {"kpts": kpts, "scores": scores, "counts": counts, "truth": truth}


def _make_chunk(rig, chunk, P, J, seed, noise_px, low_score_frac, drop_prob, shuffle, kst):
    C, n = rig.C, CHUNK
    rng = np.random.Generator(np.random.Philox(key=[seed, chunk]))
    centre = np.concatenate([rng.uniform(-2, 2, (n, P, 1, 2)), np.zeros((n, P, 1, 1))], -1)
    body = np.concatenate([rng.uniform(-0.4, 0.4, (n, P, J, 2)), rng.uniform(0, 1.8, (n, P, J, 1))], -1)
    X = centre + body                                            # (n,P,J,3)
    uv = project(rig, X) + noise_px * rng.standard_normal((C, n, P, J, 2))
    uv = np.moveaxis(uv, 0, 1).astype(np.float32)                # (n,C,P,J,2)
    sc = rng.uniform(0.6, 1.0, (n, C, P, J)).astype(np.float32)
    if low_score_frac > 0:
        low = rng.random((n, C, P, J)) < low_score_frac
        sc = np.where(low, np.float32(0.5 * kst) * rng.random((n, C, P, J), dtype=np.float32), sc)
    present = rng.random((n, C, P)) >= drop_prob
    order = rng.random((n, C, P)) if shuffle else np.tile(np.arange(P, dtype=np.float64), (n, C, 1))
    # absent persons sort last; present ones keep a random (detector-like) order (main.py:54)
    perm = np.argsort(np.where(present, order, 2.0 + order), axis=-1)
    uv = np.take_along_axis(uv, perm[..., None, None], axis=2)
    sc = np.take_along_axis(sc, perm[..., None], axis=2)
    counts = present.sum(-1).astype(np.int32)
    valid = np.arange(P)[None, None, :] < counts[..., None]
    uv = np.where(valid[..., None, None], uv, np.float32(0))
    sc = np.where(valid[..., None], sc, np.float32(0))
    return uv, sc.astype(np.float32), counts, X


DEFAULT_PARAMS = dict(kst=0.5, ast=0.0, dthr=0.05, cond_tol=10.0, num_tol=0, score_tol=0.0, center=0)
MULTI_PARAMS = dict(kst=0.5, ast=0.2, dthr=0.05, cond_tol=0.3, num_tol=0, score_tol=0.0, center=0)


def make_frames_torch(rig, F, P, J, seed=1234, device="cuda", noise_px=0.5, chunk=8192):
    """Same distribution as ``make_frames`` generated on the device with torch (benchmark inputs only;
    not bit-identical to the NumPy generator).  Returns (kpts (F,C,P,J,2) f32, scores (F,C,P,J) f32)."""
    import torch
    C = rig.C
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    Rt = torch.as_tensor(rig.R, device=device).transpose(1, 2).contiguous()      # world->camera
    K = torch.as_tensor(rig.K, device=device)
    t = torch.as_tensor(rig.t, device=device)
    kpts = torch.empty((F, C, P, J, 2), dtype=torch.float32, device=device)
    scores = torch.empty((F, C, P, J), dtype=torch.float32, device=device)
    for f0 in range(0, F, chunk):
        n = min(chunk, F - f0)
        centre = torch.zeros((n, P, 1, 3), dtype=torch.float64, device=device)
        centre[..., :2] = torch.rand((n, P, 1, 2), generator=gen, dtype=torch.float64, device=device) * 4 - 2
        body = torch.rand((n, P, J, 3), generator=gen, dtype=torch.float64, device=device)
        body[..., :2] = body[..., :2] * 0.8 - 0.4
        body[..., 2] *= 1.8
        X = centre + body                                                           # (n,P,J,3)
        Xc = torch.einsum("cij,ncpqj->ncpqi", Rt, X[:, None] - t[None, :, None, None, :])
        uvw = torch.einsum("cij,ncpqj->ncpqi", K, Xc)
        uv = uvw[..., :2] / uvw[..., 2:3]
        uv = uv + noise_px * torch.randn(uv.shape, generator=gen, dtype=torch.float64, device=device)
        # independent person order per camera, like a detector (main.py:54)
        perm = torch.argsort(torch.rand((n, C, P), generator=gen, device=device), dim=-1)
        uv = torch.gather(uv, 2, perm[..., None, None].expand(-1, -1, -1, J, 2))
        kpts[f0:f0 + n] = uv.to(torch.float32)
        scores[f0:f0 + n] = torch.rand((n, C, P, J), generator=gen, device=device) * 0.4 + 0.6
    return kpts, scores


def frames_block(C, P, J):
    """Frames per independently seeded block of ``make_frames_torch_range`` (a function of the shape only, so that
    every rank of a sharded run generates the same frames as the single-GPU run)."""
    return int(max(16, min(1024, (1 << 24) // max(1, C * P * J))))


def make_frames_torch_range(rig, f_lo, f_hi, P, J, seed=1234, device="cuda", noise_px=0.5):
    """Frames [f_lo, f_hi) of the endless synthetic clip ``seed``: same distribution as ``make_frames_torch``, but a
    frame's content depends only on (seed, global frame index), not on how the clip is cut into shards or batches.
    Returns (kpts (n,C,P,J,2) f32, scores (n,C,P,J) f32)."""
    import torch
    C = rig.C
    B = frames_block(C, P, J)
    Rt = torch.as_tensor(rig.R, device=device).transpose(1, 2).contiguous()
    K = torch.as_tensor(rig.K, device=device)
    t = torch.as_tensor(rig.t, device=device)
    n_tot = max(0, f_hi - f_lo)
    kpts = torch.empty((n_tot, C, P, J, 2), dtype=torch.float32, device=device)
    scores = torch.empty((n_tot, C, P, J), dtype=torch.float32, device=device)
    gen = torch.Generator(device=device)
    for b in range(f_lo // B, (f_hi + B - 1) // B if n_tot else 0):
        gen.manual_seed((seed * 1000003 + b) & 0x7FFFFFFFFFFF)
        centre = torch.zeros((B, P, 1, 3), dtype=torch.float64, device=device)
        centre[..., :2] = torch.rand((B, P, 1, 2), generator=gen, dtype=torch.float64, device=device) * 4 - 2
        body = torch.rand((B, P, J, 3), generator=gen, dtype=torch.float64, device=device)
        body[..., :2] = body[..., :2] * 0.8 - 0.4
        body[..., 2] *= 1.8
        X = centre + body
        Xc = torch.einsum("cij,ncpqj->ncpqi", Rt, X[:, None] - t[None, :, None, None, :])
        uvw = torch.einsum("cij,ncpqj->ncpqi", K, Xc)
        uv = uvw[..., :2] / uvw[..., 2:3]
        uv = uv + noise_px * torch.randn(uv.shape, generator=gen, dtype=torch.float64, device=device)
        perm = torch.argsort(torch.rand((B, C, P), generator=gen, device=device), dim=-1)
        uv = torch.gather(uv, 2, perm[..., None, None].expand(-1, -1, -1, J, 2)).to(torch.float32)
        sc = torch.rand((B, C, P, J), generator=gen, device=device) * 0.4 + 0.6
        lo, hi = max(f_lo, b * B), min(f_hi, (b + 1) * B)
        kpts[lo - f_lo:hi - f_lo] = uv[lo - b * B:hi - b * B]
        scores[lo - f_lo:hi - f_lo] = sc[lo - b * B:hi - b * B]
    return kpts, scores
