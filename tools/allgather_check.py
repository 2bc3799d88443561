#!/usr/bin/env python
"""Two-or-more-rank check of the C-ABI all-gather (run under torchrun on a multi-GPU box):
each rank triangulates its own frame block, the blocks are gathered with snowtri_allgather (handle-owned NCCL
communicator) and with torch.distributed; both must be bit-identical and in frame order."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snowmocap_b200 import synth  # noqa: E402
from snowmocap_b200.dist import all_gather_frames, all_gather_frames_native, init_native_comm  # noqa: E402
from snowmocap_b200.engine import TriangulationEngine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
z = np.load(os.path.join(ROOT, "tests", "golden", "floor_rig.npz"))
rig = synth.Rig(z["K"], z["R"], z["t"])
eng = TriangulationEngine(rig.K, rig.R, rig.t, device=local, precision="f32", **synth.DEFAULT_PARAMS)
F = 4096
kpts, scores = synth.make_frames_torch(rig, F, 1, 133, seed=1234 + rank, device=dev)
res = eng.run(kpts, scores, None, Pout=1)
init_native_comm(eng)
a = all_gather_frames_native(eng, res["out"], world)
b = all_gather_frames(res["out"], F * world)
torch.cuda.synchronize()
ok = bool(torch.equal(a, b)) and bool(torch.equal(a[rank * F:(rank + 1) * F], res["out"]))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"op": "snowtri_allgather", "world": world, "frames_per_rank": F, "bytes_per_rank": res["out"].numel() * 4,
                      "equal_to_torch_and_in_frame_order_on_every_rank": bool(flag.item())}))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
