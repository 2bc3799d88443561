#!/usr/bin/env python
"""Benchmark of the fused triangulation path (BASELINE.json metric: 3D keypoints/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg1] [--impl reference]

A "step" is one pass of the fused path (one kernel launch) over one batch of synthetic frames
that is larger than L2, already resident in HBM.  `value` = F*P*J*K / device time (CUDA events,
max over ranks).  `e2e` is the same metric through the host-buffer C-ABI call
(snowtri_run_host: pinned host -> device, kernel, device -> host every step).
Multi-GPU: frames shard across ranks (weak scaling, no data-path collective); the final
all-gather of the 3D joints that north_star mentions is timed once, outside the steps
(`allgather`: torch.distributed and, under `c_abi`, snowtri_allgather checked against it).
`downstream` (N=1): ms per batch of the later per-frame stages of the reference's main.py on the
step's output, device-resident: snowtri_smooth_run, snowtri_blender_run, snowtri_blender_smooth_run.
`--impl reference` times the CPU restatement of the reference's own implementation
(oracle/loop_oracle.py: per-keypoint np.linalg.inv, per-pair 2x2 solve) on all host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (rig, C, P, J, frames/GPU, params, Pout, description)
    "cfg1": ("floor", 2, 1, 17, 1 << 20, "default", 1, "2 cameras, 1 person, 17 COCO keypoints (BASELINE configs[0])"),
    "cfg2": ("floor", 4, 1, 133, 1 << 17, "default", 1,
             "4 cameras (camera_group_floor.json calibration), 1 person, 133 Wholebody keypoints (BASELINE configs[1])"),
    "c8p1": ("ring", 8, 1, 133, 1 << 16, "default", 1, "8 cameras, 1 person, 133 keypoints (single performer in a larger rig)"),
    "cfg3": ("ring", 8, 4, 133, 10000, "multi", 8, "8 cameras, 4 persons, 133 keypoints, 10k frames (BASELINE configs[2])"),
    "cfg4": ("ring", 16, 8, 133, 2000, "multi", 16,
             "16 cameras, 8 persons, 133 keypoints (BASELINE configs[3] geometry; 2000 frames per GPU per step)"),
    "cfg5": ("ring", 32, 16, 133, 200, "multi", 32,
             "32 cameras, 16 persons, 133 keypoints (BASELINE configs[4] geometry; 200 frames per GPU per step)"),
}


def load_rig(kind, C):
    from snowmocap_b200 import synth
    if kind == "floor":
        z = np.load(os.path.join(ROOT, "tests", "golden", "floor_rig.npz"))
        return synth.Rig(z["K"], z["R"], z["t"]).subset(C)
    return synth.ring_rig(C)


def params_of(kind):
    from snowmocap_b200 import synth
    return dict(synth.DEFAULT_PARAMS if kind == "default" else synth.MULTI_PARAMS)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
def _loop_worker(args):
    from oracle import loop_oracle
    kpts, scores, counts, K, R, t, prm = args
    t0 = time.perf_counter()
    for f in range(kpts.shape[0]):
        loop_oracle.fused_frame(kpts[f], scores[f], counts[f], K, R, t, prm)
    return time.perf_counter() - t0


def cpu_reference_port(rig, P, J, prm, frames_per_core, cores, seed=77):
    """Wall-clock throughput (3D keypoints/s) of the Python restatement of the reference on `cores` processes."""
    import multiprocessing as mp
    from snowmocap_b200 import synth
    d = synth.make_frames(rig, frames_per_core * cores, P, J, seed=seed)
    jobs = [(d["kpts"][i::cores], d["scores"][i::cores], d["counts"][i::cores], rig.K, rig.R, rig.t, prm)
            for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        _loop_worker(jobs[0])
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_loop_worker, jobs)
    dt = time.perf_counter() - t0
    return frames_per_core * cores * P * J / dt, dt


def cpu_c_port(rig, P, J, prm, pout, frames, seed=78):
    from oracle import c_oracle
    from snowmocap_b200 import synth
    d = synth.make_frames(rig, frames, P, J, seed=seed)
    threads = c_oracle.max_threads()
    c_oracle.fused(d["kpts"][:8], d["scores"][:8], d["counts"][:8], rig.K, rig.R, rig.t, prm, Pout=pout)
    t0 = time.perf_counter()
    c_oracle.fused(d["kpts"], d["scores"], d["counts"], rig.K, rig.R, rig.t, prm, Pout=pout)
    dt = time.perf_counter() - t0
    return frames * P * J / dt, threads, dt


def loop_frames_per_core(C, P, J, budget_s=12.0):
    solves = C * (C - 1) // 2 * P * P * J
    per_frame = solves * 30e-6 + C * P * J * 14e-6 + 1e-3        # SURVEY section 6 probe figures
    return max(1, int(budget_s / per_frame))


def run_reference_arm(args, wl):
    rig_kind, C, P, J, _, pk, pout, desc = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rig, prm = load_rig(rig_kind, C), params_of(pk)
    cores = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    per_core = max(1, loop_frames_per_core(C, P, J, budget_s=60.0 / (steps + warmup)))
    vals = []
    for s in range(steps + warmup):
        v, dt = cpu_reference_port(rig, P, J, prm, per_core, cores, seed=500 + s)
        if s >= warmup:
            vals.append((v, dt))
    value = sum(per_core * cores * P * J for _ in vals) / sum(dt for _, dt in vals)
    sample = f"{per_core * cores} frames per step ({per_core} per process x {cores} processes), python loop port of the reference"
    line = {"impl": "reference", "metric": "3d_keypoints_per_sec", "value": value, "unit": "keypoints/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * sum(dt for _, dt in vals) / len(vals), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "description": desc, "C": C, "P": P, "J": J},
            "cpu_baseline": {"value": value, "unit": "keypoints/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "keypoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ---------------------------------------------------------------------------------------------
def _emit(line):
    """Print the one JSON line on the real stdout (fd saved before libraries could write banners to it)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def time_downstream(torch, eng, out, pout, J, reps=5):
    """Device-resident continuation of the step's output through the remaining per-frame stages of the reference's
    main.py (:72-87): Human_Triangulation_Smooth, Human_Triangulation_Blender, Human_Triangulation_Blender_Smooth.
    Times in ms per batch (CUDA events, mean of ``reps`` after one warm-up)."""
    from snowmocap_b200.blender import BlenderControl, BlenderSmoothState
    from snowmocap_b200.engine import SmoothState
    pts, nout = out["out"].clone(), out["nout"]
    F = pts.shape[0]
    sm = SmoothState(eng, pout, J, 2.5, 0.75, 0.0)          # configs/snowmocap_default_config.json:18-21
    bc = BlenderControl(eng)
    bs = BlenderSmoothState(eng, pout, [[2.5, 0.75, 0.0]] * 24)
    ctrl, valid = bc.run(pts, nout)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def smooth():
        sm.reset()
        sm.run(pts, nout, 1 / 30)

    def bsmooth():
        bs.reset()
        bs.run(ctrl, valid, nout, 1 / 30)

    res = {"frames": int(F), "persons_per_frame_slots": int(pout),
           "snowtri_smooth_run_ms": timed(smooth),
           "snowtri_blender_run_ms": timed(lambda: bc.run(pts, nout)),
           "snowtri_blender_smooth_run_ms": timed(bsmooth)}
    sm.close()
    bs.close()
    return res


def main():
    global _REAL_STDOUT
    # NCCL / torchrun print banners on stdout; the driver wants exactly one JSON line there
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (0 = workload default)")
    ap.add_argument("--precision", default="auto", choices=["auto", "f64", "f32", "mixed", "f32x"],
                    help="auto = f32 for one person per camera (single-person kernel), f64 otherwise")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="multi-GPU: do not bind each rank to its GPU's NUMA node")
    ap.add_argument("--no-others", action="store_true", help="skip the short runs of the other precision modes")
    ap.add_argument("--general-gen", type=int, default=2, choices=[1, 2],
                    help="several persons per camera: 2 = second-generation kernels (default), 1 = first generation")
    ap.add_argument("--jit", default="auto", choices=["off", "auto", "always"],
                    help="rig-specialised single-person kernel compiled at run time (NVRTC)")
    args = ap.parse_args()
    wl = args.workload
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist
    from snowmocap_b200 import synth
    from snowmocap_b200.engine import TriangulationEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        if not args.no_numa:
            from snowmocap_b200.dist import bind_host_to_gpu
            numa_cpus = bind_host_to_gpu(local)     # before the pinned buffers of the e2e leg are allocated

    rig_kind, C, P, J, F, pk, pout, desc = WORKLOADS[wl]
    F = args.frames or F
    rig, prm = load_rig(rig_kind, C), params_of(pk)
    if args.precision == "auto":
        args.precision = "f32" if P == 1 else "f64"
    eng = TriangulationEngine(rig.K, rig.R, rig.t, device=local, precision=args.precision, **prm)
    eng.set_jit(args.jit)
    eng.set_general_kernels(args.general_gen)
    kpts, scores = synth.make_frames_torch(rig, F, P, J, seed=1234 + rank, device=dev)
    out = {"out": torch.empty((F, pout, J, 4), dtype=torch.float32, device=dev),
           "pscores": torch.empty((F, pout), dtype=torch.float32, device=dev),
           "nout": torch.empty((F,), dtype=torch.int32, device=dev)}
    in_bytes = kpts.numel() * 4 + scores.numel() * 4
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity of this very batch (first frames) against the oracle, outside the timed region
    parity = None
    if rank == 0:
        from oracle import c_oracle
        nchk = min(F, 256 if C <= 4 else (32 if C <= 8 else (4 if C <= 16 else 1)))
        eng.run(kpts, scores, None, Pout=pout, out=out)
        torch.cuda.synchronize()
        ref = c_oracle.fused(kpts[:nchk].cpu().numpy(), scores[:nchk].cpu().numpy(), None, rig.K, rig.R, rig.t, prm, Pout=pout)
        got = out["out"][:nchk].cpu().numpy().astype(np.float64)
        m = np.arange(pout)[None, :] < np.minimum(ref["nout"], pout)[:, None]
        ks, kr = got[m][..., 3], ref["kscores"][m]
        nz = kr != 0
        parity = {"frames": nchk, "oracle": "oracle/snow_oracle.c (float64)",
                  "nout_equal": bool(np.array_equal(out["nout"][:nchk].cpu().numpy(), ref["nout"])),
                  "rel_l2_points": float(np.linalg.norm(got[m][..., :3] - ref["points"][m]) / np.linalg.norm(ref["points"][m])),
                  "tolerance_rel_l2_points": 1e-4,
                  "zero_pattern_equal": bool(np.array_equal(ks == 0, kr == 0)),
                  "median_rel_err_kscores": float(np.median(np.abs(ks[nz] - kr[nz]) / kr[nz])) if nz.any() else 0.0,
                  "mean_persons": float(ref["nout"].mean())}

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(warmup):
        eng.run(kpts, scores, None, Pout=pout, out=out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.nvtx.range_push("timed")      # lets `ncu --nvtx --nvtx-include timed/` list exactly these launches
    ev0.record()
    for _ in range(steps):
        eng.run(kpts, scores, None, Pout=pout, out=out)
    ev1.record()
    barrier()
    torch.cuda.nvtx.range_pop()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    launch_info = eng.last_launch_info()
    if rank == 0:
        time.sleep(0.1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    kp_per_step = F * P * J * world
    value = kp_per_step * steps / (ms_max * 1e-3)

    # ---- the other precision modes of the same kernel, same batch (short runs, reported beside the headline)
    others = None
    if rank == 0 and world == 1 and not args.no_others and P == 1:
        from oracle import c_oracle
        others = {}
        nchk = min(F, 256)
        ref = c_oracle.fused(kpts[:nchk].cpu().numpy(), scores[:nchk].cpu().numpy(), None, rig.K, rig.R, rig.t, prm, Pout=pout)
        mm = np.arange(pout)[None, :] < np.minimum(ref["nout"], pout)[:, None]
        for prec in ("f32", "mixed", "f64"):
            if prec == args.precision:
                continue
            eng.set_precision(prec)
            for _ in range(3):
                eng.run(kpts, scores, None, Pout=pout, out=out)
            torch.cuda.synchronize()
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0.record()
            for _ in range(10):
                eng.run(kpts, scores, None, Pout=pout, out=out)
            o1.record()
            torch.cuda.synchronize()
            oms = o0.elapsed_time(o1) / 10
            got = out["out"][:nchk].cpu().numpy().astype(np.float64)
            others[prec] = {"ms_per_step": oms, "value": F * P * J / (oms * 1e-3),
                            "roofline_frac": (12 * C + 16) * P * J * F / (oms * 1e-3) / 1e9 / measured_peak()[0],
                            "rel_l2_points": float(np.linalg.norm(got[mm][..., :3] - ref["points"][mm]) / np.linalg.norm(ref["points"][mm]))}
        eng.set_precision(args.precision)

    # ---- end to end through host buffers ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        hk = torch.empty(kpts.shape, dtype=torch.float32, pin_memory=True)
        hs = torch.empty(scores.shape, dtype=torch.float32, pin_memory=True)
        hk.copy_(kpts)
        hs.copy_(scores)
        ho = {"out": torch.empty(out["out"].shape, dtype=torch.float32, pin_memory=True).numpy(),
              "pscores": torch.empty(out["pscores"].shape, dtype=torch.float32, pin_memory=True).numpy(),
              "nout": torch.empty(out["nout"].shape, dtype=torch.int32, pin_memory=True).numpy()}
        hkn, hsn = hk.numpy(), hs.numpy()
        esteps = max(3, min(steps, 10))
        for _ in range(2):
            eng.run_host(hkn, hsn, None, Pout=pout, out=ho)
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            eng.run_host(hkn, hsn, None, Pout=pout, out=ho)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": kp_per_step * esteps / float(dt.item()), "unit": "keypoints/s",
               "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(sum(v.nbytes for v in ho.values())),
               "steps": esteps, "api": "snowtri_run_host (C ABI, pinned host buffers)",
               "host_cores_bound_to_gpu_numa_node": numa_cpus}

    # ---- final all-gather of the 3D joints (timed once, not part of a step) ---------------------
    gather = None
    if world > 1:
        from snowmocap_b200.dist import all_gather_frames
        all_gather_frames(out["out"], F * world)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = all_gather_frames(out["out"], F * world)
        g1.record()
        torch.cuda.synchronize()
        gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        gather = {"ms": float(gms.item()), "bytes_received_per_gpu": int(full.numel() * 4 * (world - 1) // world),
                  "api": "torch.distributed.all_gather_into_tensor (NCCL)"}
        # the same gather through the C ABI (snowtri_allgather on a communicator owned by the handle), checked
        # against the torch.distributed result
        try:
            from snowmocap_b200.dist import all_gather_frames_native, init_native_comm
            init_native_comm(eng)
            all_gather_frames_native(eng, out["out"], world)
            barrier()
            g0.record()
            full2 = all_gather_frames_native(eng, out["out"], world)
            g1.record()
            torch.cuda.synchronize()
            gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
            dist.all_reduce(gms, op=dist.ReduceOp.MAX)
            gather["c_abi"] = {"ms": float(gms.item()), "api": "snowtri_allgather (ncclAllGather, handle-owned communicator)",
                               "equal_to_torch": bool(torch.equal(full2, full))}
            del full2
        except Exception as e:   # reported, not fatal: the gather is not part of a step
            gather["c_abi"] = {"error": str(e)[:200]}
        del full

    # ---- the rest of main.py's per-frame body on the same batch, device-resident (not part of a step) ---------
    downstream = None
    if rank == 0 and world == 1 and J >= 130 and not args.no_others:
        try:
            downstream = time_downstream(torch, eng, out, pout, J)
        except Exception as e:   # reported, not fatal
            downstream = {"error": str(e)[:200]}

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = (12 * C + 16) * P * J * F                     # per output keypoint, per launch
        # average duration of the dominant kernel's launch; the streaming general path launches 4 kernels per
        # step (keep, cluster, fuse, person score), reported as one step
        per_step = launches / steps
        kernel_ms = ms / launches if per_step == 1 else ms / steps
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
                traffic = json.load(fh).get(f"{wl}_{args.precision}" + ("_jit" if launch_info["kernel"] == "p1-jit" else ""))
        except Exception:
            pass
        solves = C * (C - 1) // 2 * P * P * J * F
        line = {"metric": "3d_keypoints_per_sec", "value": value, "unit": "keypoints/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": {"workload": wl, "description": desc, "C": C, "P": P, "J": J, "frames_per_gpu": F,
                           "thresholds": prm, "Pout": pout, "timing": "inputs_larger_than_L2" if in_bytes > 126e6 else "inputs_fit_L2",
                           "input_bytes_per_gpu": int(in_bytes), "launch": launch_info},
                "gpu_launches": int(launches), "jit": eng.jit_status, "clocks": clocks, "parity": parity, "e2e": e2e,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src,
                             "kernel": {"p1": "snowtri::p1_kernel",
                                        "p1-jit": "p1_jit (snowtri::p1_body specialised for this rig and batch shape with NVRTC)",
                                        "general": "snowtri::gen_keep_kernel + gen_cluster + gen_fuse_kernel + gen_pscore (one step)"
                                        }.get(launch_info["kernel"], "snowtri::fused_kernel"),
                             "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": int(alg_bytes),
                             "frac_of_8TBs_spec": achieved / 8000.0,
                             "pair_solves_per_sec": solves / (kernel_ms * 1e-3)},
                "other_precisions": others, "allgather": gather, "downstream": downstream}
        if not args.no_cpu and world == 1:      # CPU baseline legs: rank 0 at N=1 only
            cores = os.cpu_count() or 1
            per_core = loop_frames_per_core(C, P, J)
            v, dt = cpu_reference_port(rig, P, J, prm, per_core, cores)
            line["cpu_baseline"] = {"value": v, "unit": "keypoints/s", "cores": cores, "kind": "port",
                                    "sample": f"{per_core * cores} frames of this workload in {dt:.1f} s, python loop port "
                                              f"of the reference (oracle/loop_oracle.py), {cores} processes"}
            cf = max(cores, int(2e8 / max(1, C * (C - 1) // 2 * P * P * J)))
            cv, threads, cdt = cpu_c_port(rig, P, J, prm, pout, min(cf, 200000))
            line["cpu_c_port"] = {"value": cv, "unit": "keypoints/s", "cores": threads, "kind": "port",
                                  "sample": f"{min(cf, 200000)} frames in {cdt:.2f} s, C/OpenMP restatement (oracle/snow_oracle.c)"}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
