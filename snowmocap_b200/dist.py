"""Frame sharding across the GPUs of one box: frames are independent, so each
rank triangulates a contiguous block of frames with no data-path collective; an optional final
all-gather collects the dense 3D-joint block on every rank (NCCL over NVLink on GPUs, gloo in
the CPU tests of the host logic)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(F, rank, world):
    """Contiguous frame block [lo, hi) of ``rank``; the first F % world ranks get one extra frame."""
    base, extra = divmod(F, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_cyclic(F, rank, world, nchunks):
    """Block-cyclic frame sharding for a gather that overlaps the compute: the clip is cut into ``nchunks * world``
    equal chunks and chunk ``m`` goes to rank ``m % world`` as its ``m // world``-th piece.  Returns the list of
    global frame ranges [(lo, hi), ...] of ``rank``, in processing order; ``F`` must be a multiple of
    ``nchunks * world``.  All-gathering piece k of every rank yields the global frames
    [k*world*c, (k+1)*world*c) in order (c = F / (nchunks*world)), so the gathered clip needs no reordering."""
    if F % (nchunks * world):
        raise ValueError(f"F={F} is not a multiple of nchunks*world={nchunks * world}")
    c = F // (nchunks * world)
    return [((k * world + rank) * c, (k * world + rank + 1) * c) for k in range(nchunks)]


def all_gather_frames(local, F, group=None):
    """All-gather per-rank frame blocks (uneven allowed) into the full (F, ...) tensor on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(F, r, world)[1] - shard_range(F, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    if len(set(sizes)) == 1:
        out = torch.empty((F,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def triangulate_sharded(engine, kpts_local, scores_local, counts_local, F, Pout=None, keypoint_num=None,
                        gather=True, group=None):
    """Run the fused path on this rank's frame block; optionally all-gather (out, pscores, nout)."""
    res = engine.run(kpts_local, scores_local, counts_local, Pout=Pout, keypoint_num=keypoint_num)
    if not gather or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return res
    return {k: all_gather_frames(v, F, group) for k, v in res.items()}


def bind_host_to_gpu(device_index):
    """Pin this process to the CPU cores closest to its GPU (NVML's ideal affinity) so the pinned host buffers it
    allocates afterwards are first-touched on the GPU's NUMA node -- with one process per GPU this keeps every
    rank's host<->device copies off the inter-socket link.  Returns the number of cores bound, or 0 if NVML or the
    affinity call is unavailable (the caller carries on unbound)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        if hasattr(props, "uuid"):
            handle = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{props.uuid}")
        else:
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(
                f"{getattr(props, 'pci_domain_id', 0):08x}:{props.pci_bus_id:02x}:{getattr(props, 'pci_device_id', 0):02x}.0")
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def init_native_comm(engine, group=None):
    """Give ``engine``'s handle its own NCCL communicator over the ranks of ``group`` (collective): rank 0 draws the
    unique id through the C ABI (``snowtri_comm_unique_id``), torch.distributed ships the 128 bytes,
    every rank calls ``snowtri_comm_init``."""
    import ctypes as ct
    from . import _lib
    lib = _lib.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    buf = ct.create_string_buffer(128)
    if rank == 0:
        _lib.check(lib.snowtri_comm_unique_id(buf))
    box = [bytes(buf.raw)]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ident = ct.create_string_buffer(box[0], 128)
    with torch.cuda.device(engine.device):
        _lib.check(lib.snowtri_comm_init(engine._h, ident, world, rank), engine._h)


def all_gather_frames_native(engine, local, world, stream=None):
    """``snowtri_allgather`` of equal-sized per-rank frame blocks (F/world frames each) through the handle's own
    communicator (``init_native_comm`` first): returns the (F, ...) tensor, rank r's frames at block r."""
    from . import _lib
    local = local.contiguous()
    out = torch.empty((local.shape[0] * world,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    st = torch.cuda.current_stream(local.device).cuda_stream if stream is None else stream
    with torch.cuda.device(engine.device):
        _lib.check(engine._lib.snowtri_allgather(engine._h, local.data_ptr(), out.data_ptr(),
                                                 local.numel() * local.element_size(), None, st), engine._h)
    return out


def triangulate_cyclic_overlapped(engine, pieces, local, full, world, comm_stream, Pout=None, keypoint_num=None):
    """Block-cyclic sharded run with the final all-gather of the 3D joints overlapped with the compute (north_star:
    "NCCL over NVLink appears only as a final all-gather of 3D joints"): piece k of this rank (``shard_cyclic``) is
    triangulated on the current stream, then its ``out`` block is all-gathered with ``snowtri_allgather`` on
    ``comm_stream`` while piece k+1 is being computed; the small ``pscores`` / ``nout`` arrays of all pieces travel in
    one gather each after the last piece.

      pieces  [(kpts_k, scores_k, counts_k or None), ...] device tensors of this rank, c frames each
      local   dict(out (n*c,Pout,Jout,4), pscores (n*c,Pout), nout (n*c,)) -- this rank's results, piece k at rows
              [k*c, (k+1)*c); or a list of n such dicts, one per piece (then everything is gathered piece by piece)
      full    dict(out (F,Pout,Jout,4), pscores (F,Pout), nout (F,)) -- the gathered clip, frames in global order

    ``init_native_comm(engine)`` first.  The current stream waits for the last gather before returning."""
    from . import _lib
    main = torch.cuda.current_stream(engine.device)
    n = len(pieces)
    c = pieces[0][1].shape[0] if pieces else 0
    per_piece = isinstance(local, (list, tuple))

    def gather(src, dst):
        _lib.check(engine._lib.snowtri_allgather(engine._h, src.data_ptr(), dst.data_ptr(), src.numel() * src.element_size(),
                                                 None, comm_stream.cuda_stream), engine._h)

    with torch.cuda.device(engine.device):
        for k, (kp, sc, cn) in enumerate(pieces):
            loc = local[k] if per_piece else {name: t[k * c:(k + 1) * c] for name, t in local.items()}
            engine.run(kp, sc, cn, Pout=Pout, keypoint_num=keypoint_num, out=loc)
            ev = torch.cuda.Event()
            ev.record(main)
            comm_stream.wait_event(ev)
            lo = k * world * c
            for name in (("out", "pscores", "nout") if per_piece else ("out",)):
                gather(loc[name], full[name][lo:lo + world * c])
        if not per_piece and n:
            # rank-major (world, n, c, ...) from one gather of every small array; the clip wants piece-major (n, world, c, ...)
            with torch.cuda.stream(comm_stream):
                for name in ("pscores", "nout"):
                    tmp = torch.empty((world,) + tuple(local[name].shape), dtype=local[name].dtype, device=engine.device)
                    gather(local[name], tmp)
                    shape = tuple(local[name].shape[1:])
                    full[name].view((n, world, c) + shape).copy_(tmp.view((world, n, c) + shape).transpose(0, 1))
    main.wait_stream(comm_stream)
