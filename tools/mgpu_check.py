#!/usr/bin/env python
"""Multi-rank parity check (run under torchrun on a box with >= 2 GPUs; tests/test_multi_gpu.py launches it):
a fixed synthetic clip is sharded over the ranks -- contiguous blocks and block-cyclic pieces with the all-gather
overlapped with the compute -- and the gathered result on EVERY rank must be byte-identical to the same clip run on
one GPU (frames are independent, SURVEY.md 8e / section 4 "Multi-GPU test").  Two cases: BASELINE configs[2]
geometry (several persons per camera, second-generation general kernels) and configs[1] geometry (single-person
kernel).  Prints one JSON line on rank 0; exit code 1 on any mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snowmocap_b200 import synth  # noqa: E402
from snowmocap_b200.dist import (all_gather_frames, all_gather_frames_native, init_native_comm, shard_cyclic,  # noqa: E402
                                 shard_range, triangulate_cyclic_overlapped)
from snowmocap_b200.engine import TriangulationEngine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def new_out(F, pout, J):
    return {"out": torch.empty((F, pout, J, 4), dtype=torch.float32, device=dev),
            "pscores": torch.empty((F, pout), dtype=torch.float32, device=dev),
            "nout": torch.empty((F,), dtype=torch.int32, device=dev)}


def case(name, rig, P, J, pout, prm, precision, Ft, pieces_n, exact_pscores):
    eng = TriangulationEngine(rig.K, rig.R, rig.t, device=local, precision=precision, **prm)
    init_native_comm(eng)
    # the whole clip on this GPU alone
    kp, sc = synth.make_frames_torch_range(rig, 0, Ft, P, J, seed=55, device=dev)
    one = eng.run(kp, sc, None, Pout=pout)
    torch.cuda.synchronize()
    res = {"kernel": eng.last_launch_info()["kernel"], "frames": Ft}
    # (a) contiguous blocks, gathered with torch.distributed and with the C ABI
    lo, hi = shard_range(Ft, rank, world)
    part = eng.run(kp[lo:hi].contiguous(), sc[lo:hi].contiguous(), None, Pout=pout)
    ok = True
    for k in ("out", "nout"):
        g = all_gather_frames(part[k], Ft)
        ok = ok and bool(torch.equal(g, one[k]))
        if Ft % world == 0:
            ok = ok and bool(torch.equal(all_gather_frames_native(eng, part[k], world), one[k]))
    res["contiguous_blocks_equal"] = ok
    # (b) block-cyclic pieces, the gather of piece k overlapped with the compute of piece k+1
    spans = shard_cyclic(Ft, rank, world, pieces_n)
    pieces = [(kp[a:b].contiguous(), sc[a:b].contiguous(), None) for a, b in spans]
    c = spans[0][1] - spans[0][0]
    loc = [new_out(c, pout, J) for _ in spans]
    full = new_out(Ft, pout, J)
    comm = torch.cuda.Stream()
    triangulate_cyclic_overlapped(eng, pieces, loc, full, world, comm, Pout=pout)
    torch.cuda.synchronize()
    ok2 = bool(torch.equal(full["out"], one["out"])) and bool(torch.equal(full["nout"], one["nout"]))
    if exact_pscores:
        ok2 = ok2 and bool(torch.equal(full["pscores"], one["pscores"]))
    else:  # the single-person kernel's person score is a float32 sum whose lane partition depends on the tile
        ok2 = ok2 and bool(torch.allclose(full["pscores"], one["pscores"], rtol=1e-5, atol=1e-7))
    res["cyclic_overlapped_equal"] = ok2
    # the same with one local result array (small arrays gathered once after the last piece)
    loc_all, full2 = new_out(c * len(spans), pout, J), new_out(Ft, pout, J)
    triangulate_cyclic_overlapped(eng, pieces, loc_all, full2, world, comm, Pout=pout)
    torch.cuda.synchronize()
    ok3 = all(bool(torch.equal(full2[k], full[k])) for k in ("out", "pscores", "nout"))
    res["cyclic_single_buffer_equal"] = ok3
    ok2 = ok2 and ok3
    res["mean_persons"] = float(one["nout"].float().mean().item())
    eng.close()
    return res, ok and ok2


out = {"world": world}
ring8 = synth.ring_rig(8)
r1, ok1 = case("cfg3", ring8, 4, 133, 8, synth.MULTI_PARAMS, "mixed", 48 * world, 3, True)
z = np.load(os.path.join(ROOT, "tests", "golden", "floor_rig.npz"))
r2, ok2 = case("cfg2", synth.Rig(z["K"], z["R"], z["t"]), 1, 133, 1, synth.DEFAULT_PARAMS, "f32", 1024 * world, 4, False)
out["cfg3_geometry"], out["cfg2_geometry"] = r1, r2
flag = torch.tensor([1 if (ok1 and ok2) else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
out["byte_identical_to_one_gpu_on_every_rank"] = bool(flag.item())
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
