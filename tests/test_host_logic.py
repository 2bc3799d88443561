"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the header
declares, the drop-in containers mirror the reference's attributes (camera.py:16-59, 141-170,
234-261), and frame sharding + all-gather work at world_size 2 over gloo."""
import json
import os
import re
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, floor_rig
from snowmocap_b200 import _lib, dist as sdist, synth
from snowmocap_b200.camera import Camera, CameraGroup


# ---- C ABI ---------------------------------------------------------------------------------
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "snowtri.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snowtri_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"libsnowtri.so does not export {n}"
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and include/snowtri.h disagree"
    assert lib.snowtri_version() >= 1


def test_create_without_gpu_fails_loudly():
    import ctypes as ct
    import torch
    if torch.cuda.is_available():
        pytest.skip("only meaningful on a box without a GPU")
    lib = _lib.load()
    h = ct.c_void_p()
    eye = np.eye(3)[None].copy()
    rc = lib.snowtri_create(ct.byref(h), 0, 1, eye.ctypes.data, eye.ctypes.data, np.zeros((1, 3)).ctypes.data)
    assert rc == _lib.E_CUDA and not h.value
    assert b"no CPU path" in lib.snowtri_last_error(None)
    from snowmocap_b200.engine import TriangulationEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TriangulationEngine(eye, eye, np.zeros((1, 3)))


def test_null_handle_is_an_error_not_a_crash():
    lib = _lib.load()
    assert lib.snowtri_set_params(None, 0.5, 0.0, 0.05, 0.1, 0, 0.0, 0) == _lib.E_ARG
    assert lib.snowtri_run(None, None, None, None, 1, 1, 1, 1, 1, None, None, None, None) == _lib.E_ARG
    assert lib.snowtri_destroy(None) == _lib.OK
    assert lib.snowtri_launch_count(None) == 0


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "snowmocap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "snow_oracle" not in src, f


# ---- drop-in containers -----------------------------------------------------------------------
def _group_json(tmp_path, rig):
    info = {"camera_num": rig.C, "camera_group_info": [
        {"cap_id": c, "frame_width": 1280, "frame_height": 720, "K": rig.K[c].tolist(), "R": rig.R[c].tolist(),
         "t": rig.t[c].reshape(3, 1).tolist(), "D": [[0.0] * 5]} for c in range(rig.C)]}
    p = tmp_path / "group.json"
    p.write_text(json.dumps(info))
    return str(p), info


def test_camera_group_round_trip(tmp_path):
    rig = floor_rig()
    path, info = _group_json(tmp_path, rig)
    g = CameraGroup(camera_group_info_path=path)
    assert g.camera_num == rig.C and len(g.cameras) == rig.C
    assert g.cameras[1].K.shape == (3, 3) and g.cameras[1].t.shape == (3, 1) and g.cameras[1].D.shape == (1, 5)
    assert g.camera_group_info_dict() == info
    out = tmp_path / "saved.json"
    g.save_camera_group_info(str(out))
    assert json.loads(out.read_text()) == info
    K, R, t = g.parameters()
    assert np.array_equal(K, rig.K) and np.array_equal(R, rig.R) and np.array_equal(t, rig.t)
    one = Camera(camera_info_dict=info["camera_group_info"][2])
    assert one.cap_id == 2 and np.array_equal(one.R, rig.R[2])


def test_default_constructor_matches_reference_defaults():
    g = CameraGroup()
    assert g.camera_num == 2 and [c.cap_id for c in g.cameras] == [0, 1]
    c = g.cameras[0]
    assert (c.frame_width, c.frame_height) == (1280, 720)
    for name in ("points", "point_rays", "hrnet_points", "hrnet_point_rays", "hrnet_point_score"):
        assert getattr(c, name) == []


def test_observation_store_and_pack_frame():
    g = CameraGroup(cap_ids=[0, 1, 2], resolutions=[(1280, 720)] * 3)
    rng = np.random.default_rng(0)
    people = {0: 2, 1: 0, 2: 1}
    ref = {}
    for c, n in people.items():
        for p in range(n):
            kp, sc = rng.random((17, 2)) * 100, rng.random(17)
            g.add_human_2D_points(kp, sc, c)
            ref[(c, p)] = (kp.astype(np.float32), sc.astype(np.float32))
    kpts, scores, counts = g.pack_frame()
    assert kpts.shape == (1, 3, 2, 17, 2) and scores.shape == (1, 3, 2, 17)
    assert counts.tolist() == [[2, 0, 1]]
    for (c, p), (kp, sc) in ref.items():
        assert np.array_equal(kpts[0, c, p], kp) and np.array_equal(scores[0, c, p], sc)
    assert not kpts[0, 1].any() and not kpts[0, 2, 1].any()
    g.clear_2D_points()
    assert all(c.hrnet_points == [] and c.hrnet_point_score == [] for c in g.cameras)
    k2, s2, c2 = g.pack_frame()
    assert k2.shape[2] == 0 and c2.tolist() == [[0, 0, 0]]


def test_pack_frame_rejects_mixed_keypoint_counts():
    g = CameraGroup()
    g.add_human_2D_points(np.zeros((17, 2)), np.zeros(17), 0)
    g.add_human_2D_points(np.zeros((133, 2)), np.zeros(133), 1)
    with pytest.raises(ValueError):
        g.pack_frame()


# ---- synthetic generator -----------------------------------------------------------------------
def test_synth_is_shardable_and_float32():
    rig = synth.ring_rig(4)
    full = synth.make_frames(rig, 600, 2, 17, seed=3, drop_prob=0.1)
    part = synth.make_frames(rig, 200, 2, 17, seed=3, drop_prob=0.1, frame0=300)
    for k in ("kpts", "scores", "counts"):
        assert np.array_equal(full[k][300:500], part[k]), k
    assert full["kpts"].dtype == np.float32 and full["scores"].dtype == np.float32
    # reprojection of the truth is within the noise of the observations (sanity of project())
    uv = synth.project(rig, full["truth"][:1])                   # (C,1,P,J,2)
    noise = full["kpts"][0][full["counts"][0] == 2]
    assert noise.size == 0 or np.isfinite(uv).all()


def test_shard_range_partitions_frames():
    for F in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            spans = [sdist.shard_range(F, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_cyclic_pieces_tile_the_clip_in_gather_order():
    """Piece k of every rank, concatenated in rank order, is the k-th contiguous stretch of the clip: an all-gather of
    piece k lands in place."""
    for world in (1, 2, 4, 8):
        for nch in (1, 3, 5):
            F = world * nch * 7
            spans = [sdist.shard_cyclic(F, r, world, nch) for r in range(world)]
            c = F // (world * nch)
            seen = []
            for k in range(nch):
                for r in range(world):
                    lo, hi = spans[r][k]
                    assert hi - lo == c
                    seen.append((lo, hi))
            assert seen[0][0] == 0 and seen[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    with pytest.raises(ValueError):
        sdist.shard_cyclic(10, 0, 4, 3)


# ---- world_size-2 gloo: sharding + all-gather reproduces the single-process result -------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _FakeEngine:
    """Stands in for the CUDA engine on CPU: a deterministic per-frame function of the inputs, so the
    test exercises only the sharding / gather plumbing (frames are independent)."""

    def run(self, kpts, scores, counts, Pout=None, keypoint_num=None):
        import torch
        F = kpts.shape[0]
        out = (kpts.sum(dim=(1, 2)).mean(dim=-1, keepdim=True) * scores.sum(dim=(1, 2))[..., None]).float()
        return {"out": out.reshape(F, 1, -1, 1).expand(F, 1, out.shape[1], 4).contiguous(),
                "pscores": scores.mean(dim=(1, 2, 3)).reshape(F, 1).float(),
                "nout": counts.sum(dim=1).to(torch.int32)}


def _gloo_worker(rank, world, port, F, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rig = synth.ring_rig(3)
        lo, hi = sdist.shard_range(F, rank, world)
        d = synth.make_frames(rig, hi - lo, 2, 17, seed=11, frame0=lo)
        res = sdist.triangulate_sharded(_FakeEngine(), torch.from_numpy(d["kpts"]), torch.from_numpy(d["scores"]),
                                        torch.from_numpy(d["counts"]), F)
        q.put((rank, {k: v.numpy() for k, v in res.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("F", [64, 37])
def test_sharded_gather_world2_gloo(F):
    import torch
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, F, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rig = synth.ring_rig(3)
    d = synth.make_frames(rig, F, 2, 17, seed=11)
    want = _FakeEngine().run(torch.from_numpy(d["kpts"]), torch.from_numpy(d["scores"]), torch.from_numpy(d["counts"]))
    for r in range(2):
        for k, v in want.items():
            assert np.array_equal(got[r][k], v.numpy()), (r, k)


def test_bind_host_to_gpu_is_a_noop_without_a_device():
    """The NUMA helper never raises: without a CUDA device (or NVML) it reports 0 bound cores and leaves the
    process affinity alone."""
    import os
    import torch
    from snowmocap_b200.dist import bind_host_to_gpu
    if torch.cuda.is_available():
        pytest.skip("CPU-side check")
    before = os.sched_getaffinity(0)
    assert bind_host_to_gpu(0) == 0
    assert os.sched_getaffinity(0) == before


def test_magic_division_of_the_single_person_kernel_is_exact():
    """snowtri_p1.cuh walks a tile by the flat item index q and gets the item's frame as umulhi(q, ceil(2^32 / Jout)).
    The header claims exactness for q < 32 * Jout + 64 as long as Jout <= kP1MaxJout = 8192 (32 * Jout^2 < 2^32): check
    the claim at the frame boundaries, where an off-by-one would show, for every Jout the kernel accepts."""
    for d in list(range(2, 300)) + [511, 512, 513, 1000, 4095, 4096, 8191, 8192]:   # one keypoint takes the general path
        m = ((1 << 32) + d - 1) // d
        assert m < (1 << 32)
        qmax = 32 * d + 64
        qs = set()
        for g in range(0, 34):
            for off in (-2, -1, 0, 1):
                q = g * d + off
                if 0 <= q <= qmax:
                    qs.add(q)
        qs.update((0, 1, qmax - 1, qmax))
        for q in qs:
            assert (q * m) >> 32 == q // d, (d, q)
