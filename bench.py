#!/usr/bin/env python
"""Benchmark of the fused triangulation path (BASELINE.json metric: 3D keypoints/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|...] [--impl reference]

A "step" is one pass of the fused path over one batch of synthetic frames that is larger than L2 and already resident
in HBM.  `value` = F*P*J*K / device time (CUDA events on the launching stream, max over ranks).  The headline workload
is BASELINE configs[1] (4 cameras, 1 person, 133 keypoints: the configuration the metric is quoted on).

Keys beside the base contract:
  roofline        dominant kernel against the roof that binds the workload: HBM bandwidth for one person per camera,
                  FP32 issue rate for several (81 FLOP/B at configs[2]); both fractions are given.
  e2e             same metric through the C ABI with HOST buffers (snowtri_run_host: pinned host -> device, kernels,
                  device -> host every step), beside a copy-only ceiling (the same bytes over the same pinned buffers,
                  both directions at once, no kernel) and the drop-in per-frame latency through the reference's names.
  cpu_baseline    the REAL reference (oracle/_ref, vendored by build() from /root/reference) on the box's host cores,
                  single process (the reference is single-threaded) and one process per core; kind "port"
                  (oracle/loop_oracle.py) when the vendored copy is absent.
  secondary.cfg3  BASELINE configs[2] (8 cameras x 4 persons, 10k frames): the only single-GPU config that runs the
                  cross-camera person matching and the multi-cluster condense; own parity (C oracle + frames of the
                  real reference), cpu_baseline, e2e and roofline.
  secondary.cfg4_strong / cfg5_strong   BASELINE configs[3] / [4] geometry, a FIXED clip sharded over the ranks
                  (strong scaling) with the all-gather of the 3D joints inside the timed step (block-cyclic shards,
                  gather of piece k overlapped with the compute of piece k+1) and a checksum of the gathered clip that
                  must not depend on N.
  value_with_allgather (N > 1)   the headline workload with that gather inside the step.
`--impl reference` times the reference's own CPU implementation on all host cores on the same config.
"""
import argparse
import hashlib
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
# strong-scaling blocks: (workload geometry, frames of the whole clip, pieces per rank)
STRONG = {"cfg4_strong": ("cfg4", 100000, 5, "BASELINE configs[3]: 16 cameras x 8 persons x 133, 100k frames sharded over the ranks"),
          "cfg5_strong": ("cfg5", 20000, 5, "BASELINE configs[4] geometry: 32 cameras x 16 persons x 133, 20k frames sharded over the ranks")}
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 148 SMs x 128 FP32 lanes x 2 (FMA) x 1965 MHz = 74.4 (nominal; no measured figure)


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


def flop_per_keypoint(C, P):
    """17*C for the rays + 80 per ray-pair solve, C(C,2)*P solves per output keypoint."""
    return 17 * C + 80 * (C * (C - 1) // 2) * P


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
def _cpu_worker(args):
    """Frames of one process through the reference's main.py:55-71 body: the real reference when it is vendored
    (oracle/_ref), else the line-by-line port."""
    kind, kpts, scores, counts, K, R, t, prm = args
    t0 = time.perf_counter()
    if kind == "reference":
        from oracle import ref_runner
        ref_runner.run_frames(K, R, t, kpts, scores, counts, prm)
    else:
        from oracle import loop_oracle
        for f in range(kpts.shape[0]):
            loop_oracle.fused_frame(kpts[f], scores[f], counts[f], K, R, t, prm)
    return time.perf_counter() - t0


def cpu_kind():
    try:
        from oracle import ref_runner
        if ref_runner.available():
            ref_runner.load()          # cv2 / scipy importable on this box?
            return "reference"
    except Exception:
        pass
    return "port"


def cpu_reference(rig, P, J, prm, frames_per_core, cores, seed=77, kind=None):
    """Wall-clock throughput (3D keypoints/s) of the reference's CPU implementation on `cores` processes."""
    import multiprocessing as mp
    from snowmocap_b200 import synth
    kind = kind or cpu_kind()
    d = synth.make_frames(rig, frames_per_core * cores, P, J, seed=seed)
    jobs = [(kind, d["kpts"][i::cores], d["scores"][i::cores], d["counts"][i::cores], rig.K, rig.R, rig.t, prm)
            for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        _cpu_worker(jobs[0])
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_worker, jobs)
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


def frames_for_budget(C, P, J, budget_s):
    solves = C * (C - 1) // 2 * P * P * J
    per_frame = solves * 30e-6 + C * P * J * 14e-6 + 1e-3        # SURVEY section 6 probe figures
    return max(1, int(budget_s / per_frame))


def cpu_baseline_block(rig, C, P, J, prm, pout, budget_s=10.0):
    """`cpu_baseline` of one workload: single process (the reference is single-threaded), one process per core beside
    it, and the C/OpenMP restatement as a best-effort CPU line."""
    kind, cores = cpu_kind(), len(os.sched_getaffinity(0)) or 1
    what = ("the real reference (oracle/_ref/snowvision: add_human_2D_points + Human_Triangulation + "
            "Human_Triangulation_Condense + clear_2D_points per frame)" if kind == "reference"
            else "python loop port of the reference (oracle/loop_oracle.py; oracle/_ref not vendored on this box)")
    n1 = frames_for_budget(C, P, J, budget_s)
    v1, dt1 = cpu_reference(rig, P, J, prm, n1, 1, kind=kind)
    block = {"value": v1, "unit": "keypoints/s", "cores": 1, "kind": kind,
             "sample": f"{n1} frames of this workload in {dt1:.1f} s, single process, {what}",
             "host_cores": cores}
    if cores > 1:
        per = frames_for_budget(C, P, J, budget_s)
        va, dta = cpu_reference(rig, P, J, prm, per, cores, seed=79, kind=kind)
        block["all_cores"] = {"value": va, "unit": "keypoints/s", "cores": cores, "kind": kind,
                              "sample": f"{per * cores} frames in {dta:.1f} s, {cores} processes x {per} frames"}
    try:
        cf = max(cores, int(2e8 / max(1, C * (C - 1) // 2 * P * P * J)))
        cv, threads, cdt = cpu_c_port(rig, P, J, prm, pout, min(cf, 200000))
        block["c_port"] = {"value": cv, "unit": "keypoints/s", "cores": threads, "kind": "port",
                           "sample": f"{min(cf, 200000)} frames in {cdt:.2f} s, C/OpenMP restatement (oracle/snow_oracle.c)"}
    except Exception as e:
        block["c_port"] = {"error": str(e)[:200]}
    return block


def workload_config(wl, F):
    """The `config` object of the JSON line: the workload and nothing about how an arm ran it, so both arms print the
    same object (the reference arm times a bounded sample of it, described in its `cpu_baseline.sample`)."""
    _, C, P, J, _, pk, pout, desc = WORKLOADS[wl]
    in_bytes = F * C * P * J * 12
    return {"workload": wl, "description": desc, "C": C, "P": P, "J": J, "frames_per_gpu": int(F), "thresholds": params_of(pk),
            "Pout": pout, "timing": "inputs_larger_than_L2" if in_bytes > 126e6 else "inputs_fit_L2",
            "input_bytes_per_gpu": int(in_bytes)}


def run_reference_arm(args, wl):
    rig_kind, C, P, J, _, pk, pout, desc = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rig, prm = load_rig(rig_kind, C), params_of(pk)
    cores, kind = len(os.sched_getaffinity(0)) or 1, cpu_kind()
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    per_core = max(1, frames_for_budget(C, P, J, budget_s=60.0 / (steps + warmup)))
    vals = []
    for s in range(steps + warmup):
        v, dt = cpu_reference(rig, P, J, prm, per_core, cores, seed=500 + s, kind=kind)
        if s >= warmup:
            vals.append((v, dt))
    value = sum(per_core * cores * P * J for _ in vals) / sum(dt for _, dt in vals)
    what = "the real reference (oracle/_ref/snowvision)" if kind == "reference" else "python loop port of the reference"
    sample = f"{per_core * cores} frames per step ({per_core} per process x {cores} processes), {what}"
    line = {"impl": "reference", "metric": "3d_keypoints_per_sec", "value": value, "unit": "keypoints/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * sum(dt for _, dt in vals) / len(vals), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(wl, args.frames if args.frames > 0 else WORKLOADS[wl][4]),
            "cpu_baseline": {"value": value, "unit": "keypoints/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "keypoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ---------------------------------------------------------------------------------------------
def _emit(line):
    """Print the one JSON line on the real stdout (fd saved before libraries could write banners to it)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def new_out(torch, F, pout, J, dev):
    return {"out": torch.empty((F, pout, J, 4), dtype=torch.float32, device=dev),
            "pscores": torch.empty((F, pout), dtype=torch.float32, device=dev),
            "nout": torch.empty((F,), dtype=torch.int32, device=dev)}


def parity_block(torch, eng, rig, prm, kpts, scores, out, pout, C, precision):
    """This very batch (first frames) against the C oracle, outside any timed region.  Every output is bounded:
    points (rel-L2), keypoint scores (median, 99.9th percentile and maximum relative error), persons per frame."""
    from oracle import c_oracle
    F = kpts.shape[0]
    nchk = min(F, 256 if C <= 4 else (32 if C <= 8 else (4 if C <= 16 else 1)))
    eng.run(kpts, scores, None, Pout=pout, out=out)
    torch.cuda.synchronize()
    ref = c_oracle.fused(kpts[:nchk].cpu().numpy(), scores[:nchk].cpu().numpy(), None, rig.K, rig.R, rig.t, prm, Pout=pout)
    got = out["out"][:nchk].cpu().numpy().astype(np.float64)
    m = np.arange(pout)[None, :] < np.minimum(ref["nout"], pout)[:, None]
    ks, kr = got[m][..., 3], ref["kscores"][m]
    nz = kr != 0
    rel = np.abs(ks[nz] - kr[nz]) / kr[nz] if nz.any() else np.zeros(1)
    ps, pr = out["pscores"][:nchk].cpu().numpy().astype(np.float64)[m], ref["pscores"][m]
    # all-float32: a score is 1/distance of two nearly intersecting rays; the float32 ray distance is good to 1e-5 m,
    # which bounds every keypoint score's relative error by 1e-5 * 2000 * n_pairs * score + 1e-3 (tests/test_gpu_parity.py,
    # score_error_bound).  The float64-numerator modes hold 1e-4 flat.
    bound = 1e-5 * 2000.0 * (C * (C - 1) // 2) * kr[nz] + 1e-3 if precision == "f32" else np.full(int(nz.sum()), 1e-4)
    return {"frames": nchk, "kscores_bound": "rel err <= 1e-5 m * 2000 * pairs * score + 1e-3 (float32 ray distance)" if precision == "f32" else "rel err <= 1e-4",
            "kscores_within_bound": bool((rel <= bound).all()) if nz.any() else True, "oracle": "oracle/snow_oracle.c (float64)", "precision": precision,
            "nout_equal": bool(np.array_equal(out["nout"][:nchk].cpu().numpy(), ref["nout"])),
            "rel_l2_points": float(np.linalg.norm(got[m][..., :3] - ref["points"][m]) / np.linalg.norm(ref["points"][m])),
            "tolerance_rel_l2_points": 1e-4,
            "zero_pattern_equal": bool(np.array_equal(ks == 0, kr == 0)),
            "median_rel_err_kscores": float(np.median(rel)), "p999_rel_err_kscores": float(np.quantile(rel, 0.999)),
            "max_rel_err_kscores": float(rel.max()),
            "max_rel_err_pscores": float((np.abs(ps - pr) / np.maximum(np.abs(pr), 1e-30)).max()) if pr.size else 0.0,
            "mean_persons": float(ref["nout"].mean())}


def reference_frames_parity(torch, eng, name):
    """Frames the REAL reference emitted at this size (tests/golden/big_*.npz, made by tests/golden/make_golden_big.py)."""
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    F = z["kpts"].shape[0]
    con = [(z[f"con_pts_{f}"], z[f"con_ks_{f}"]) for f in range(F)]
    pout = min(32, max(c[0].shape[0] for c in con))
    dev = eng.device
    res = eng.run(torch.from_numpy(z["kpts"]).to(dev), torch.from_numpy(z["scores"]).to(dev),
                  torch.from_numpy(z["counts"]).to(dev), Pout=pout)
    torch.cuda.synchronize()
    got, nout = res["out"].cpu().numpy().astype(np.float64), res["nout"].cpu().numpy()
    num = den = 0.0
    ok = True
    for f in range(F):
        n = con[f][0].shape[0]
        ok = ok and int(nout[f]) == n
        k = min(n, pout, int(nout[f]))
        num += float(((got[f, :k, :, :3] - con[f][0][:k]) ** 2).sum())
        den += float((con[f][0][:k] ** 2).sum())
    return {"fixture": f"tests/golden/{name}.npz (outputs of the real reference, {float(z['seconds_per_frame']):.1f} s per frame there)",
            "frames": int(F), "nout_equal": bool(ok), "rel_l2_points": float((num / den) ** 0.5) if den else 0.0}


def timed_steps(torch, dist, eng, kpts, scores, pout, out, steps, warmup, world, dev, nvtx=None):
    """W untimed + K timed passes, barrier + synchronize on both sides, CUDA events, max over ranks (ms)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        eng.run(kpts, scores, None, Pout=pout, out=out)
    barrier()
    l0 = eng.launch_count
    ev0, ev1 = _events(torch)
    barrier()
    if nvtx:
        torch.cuda.nvtx.range_push(nvtx)      # lets `ncu --nvtx --nvtx-include timed/` list exactly these launches
    ev0.record()
    h0 = time.perf_counter()
    for _ in range(steps):
        eng.run(kpts, scores, None, Pout=pout, out=out)
    host_ms = (time.perf_counter() - h0) * 1e3   # host time to enqueue the K steps (diagnostic: must stay below the device time)
    ev1.record()
    barrier()
    if nvtx:
        torch.cuda.nvtx.range_pop()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(per, torch.tensor([ms, host_ms], dtype=torch.float64, device=dev))
        timed_steps.per_rank = {"device_ms": [round(float(p[0]), 4) for p in per], "host_enqueue_ms": [round(float(p[1]), 4) for p in per]}
    else:
        timed_steps.per_rank = {"device_ms": [round(ms, 4)], "host_enqueue_ms": [round(host_ms, 4)]}
    return float(t.item()), ms, eng.launch_count - l0


def e2e_block(torch, dist, eng, kpts, scores, out, pout, kp_per_step, steps, world, dev, numa_cpus):
    """End to end through snowtri_run_host (pinned host buffers in, pinned host buffers out, every step), and the
    copy-only ceiling of the same bytes: host->device and device->host at once on two streams, no kernel."""
    hk = torch.empty(kpts.shape, dtype=torch.float32, pin_memory=True)
    hs = torch.empty(scores.shape, dtype=torch.float32, pin_memory=True)
    hk.copy_(kpts)
    hs.copy_(scores)
    ho_t = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in out.items()}
    ho = {k: v.numpy() for k, v in ho_t.items()}
    hkn, hsn = hk.numpy(), hs.numpy()
    esteps = max(3, min(steps, 10))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxed(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    for _ in range(2):
        eng.run_host(hkn, hsn, None, Pout=pout, out=ho)
    barrier()
    t0 = time.perf_counter()
    for _ in range(esteps):
        eng.run_host(hkn, hsn, None, Pout=pout, out=ho)
    torch.cuda.synchronize()
    dt = maxed(time.perf_counter() - t0)
    # copy-only ceiling: same pinned buffers, same bytes, both directions concurrently
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    dk, ds = torch.empty_like(kpts), torch.empty_like(scores)

    def copies():
        with torch.cuda.stream(s_in):
            dk.copy_(hk, non_blocking=True)
            ds.copy_(hs, non_blocking=True)
        with torch.cuda.stream(s_out):
            for k in ho_t:
                ho_t[k].copy_(out[k], non_blocking=True)
    copies()
    copies()
    torch.cuda.synchronize()
    dtc = float("inf")
    for _ in range(2):   # best of two passes: the first one still pays for page-table and link warm-up at N > 1
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            copies()
        torch.cuda.synchronize()
        dtc = min(dtc, maxed(time.perf_counter() - t0))
    h2d = int(kpts.numel() * 4 + scores.numel() * 4)
    d2h = int(sum(v.nbytes for v in ho.values()))
    del dk, ds
    return {"value": kp_per_step * esteps / dt, "unit": "keypoints/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": esteps, "api": "snowtri_run_host (C ABI, pinned host buffers)",
            "copy_only_ceiling": {"value": kp_per_step * esteps / dtc, "unit": "keypoints/s",
                                  "h2d_GBps_per_gpu": h2d * esteps / dtc / 1e9, "d2h_GBps_per_gpu": d2h * esteps / dtc / 1e9,
                                  "what": "the step's bytes over the same pinned buffers, both directions at once, no kernel"},
            "frac_of_copy_ceiling": dtc / dt, "host_cores_bound_to_gpu_numa_node": numa_cpus}


def dropin_leg(rig, prm, P, J, frames=20):
    """ms per frame through the reference's own names (CameraGroup.add_human_2D_points, Human_Triangulation,
    Human_Triangulation_Condense, clear_2D_points: main.py:55-71 with F = 1), host arrays in, host lists out, beside the
    real reference on the same frames when it is vendored."""
    import snowmocap_b200 as sm
    from snowmocap_b200 import synth
    d = synth.make_frames(rig, frames + 2, P, J, seed=4242)
    group = sm.CameraGroup(cap_ids=list(range(rig.C)), resolutions=[(1280, 720)] * rig.C)
    for c in range(rig.C):
        group.cameras[c].K, group.cameras[c].R, group.cameras[c].t = rig.K[c], rig.R[c], rig.t[c].reshape(3, 1)

    def frame(f):
        for c in range(rig.C):
            for p in range(int(d["counts"][f, c])):
                group.add_human_2D_points(d["kpts"][f, c, p], d["scores"][f, c, p], c)
        tri = sm.Human_Triangulation(group, prm["kst"], prm["ast"], prm["dthr"])
        con = sm.Human_Triangulation_Condense(tri, prm["cond_tol"], prm["num_tol"], prm["score_tol"], prm["center"], J)
        group.clear_2D_points()
        return con
    frame(0)
    frame(1)
    t0 = time.perf_counter()
    for f in range(2, frames + 2):
        frame(f)
    ms = 1e3 * (time.perf_counter() - t0) / frames
    res = {"ms_per_frame": ms, "frames": frames, "api": "snowmocap_b200.CameraGroup / Human_Triangulation / Human_Triangulation_Condense (F = 1)"}
    try:
        from oracle import ref_runner
        if cpu_kind() == "reference":
            n = min(frames, 10)
            _, dt = ref_runner.run_frames(rig.K, rig.R, rig.t, d["kpts"][2:2 + n], d["scores"][2:2 + n], d["counts"][2:2 + n], prm)
            res["reference_ms_per_frame"] = 1e3 * dt / n
    except Exception as e:
        res["reference_error"] = str(e)[:200]
    return res


def time_downstream(torch, eng, out, pout, J, reps=5):
    """Device-resident continuation of the step's output through the remaining per-frame stages of the reference's
    main.py (:72-87): Human_Triangulation_Smooth, Human_Triangulation_Blender, Human_Triangulation_Blender_Smooth.
    Times in ms per batch (CUDA events, mean of ``reps`` after one warm-up) and fractions of the measured HBM peak on
    the algorithmic bytes of each stage."""
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
        e0, e1 = _events(torch)
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

    peak = measured_peak()[0]
    t_s, t_b, t_bs = timed(smooth), timed(lambda: bc.run(pts, nout)), timed(bsmooth)
    rows = F * pout
    res = {"frames": int(F), "persons_per_frame_slots": int(pout),
           "snowtri_smooth_run_ms": t_s, "snowtri_blender_run_ms": t_b, "snowtri_blender_smooth_run_ms": t_bs,
           # algorithmic bytes: smoothing reads and writes one float4 per joint (32 B); control points read 28 joints
           # x 12 B and write 24 x 16 B + 4 B per person row (724 B); their smoothing reads and writes 24 x 16 B
           "snowtri_smooth_run_frac_of_hbm_peak": rows * J * 32 / (t_s * 1e-3) / 1e9 / peak,
           "snowtri_blender_run_frac_of_hbm_peak": rows * 724 / (t_b * 1e-3) / 1e9 / peak,
           "snowtri_blender_smooth_run_frac_of_hbm_peak": rows * 24 * 32 / (t_bs * 1e-3) / 1e9 / peak}
    res["total_ms"] = t_s + t_b + t_bs
    sm.close()
    bs.close()
    return res


def downstream_cpu_baseline(pout, J, seconds=3.0):
    """CPU legs of the downstream stages on this box's host cores, one core each (the reference calls them once per
    frame from a single thread), on bounded samples: the real reference's Human_Triangulation_Smooth when the vendored
    copy is present (else the C restatement), and the NumPy restatement of Human_Triangulation_To_Blender's control
    points.  Test infrastructure timed as a baseline, never on the product path."""
    res = {}
    rng = np.random.default_rng(3)
    try:
        from oracle import ref_runner
        if ref_runner.available():
            ref = ref_runner.load()
            n, prev, t0 = 0, None, time.perf_counter()
            pts = rng.uniform(-1, 1, (64, pout, J, 3))
            while time.perf_counter() - t0 < seconds:
                frame = {"hrnet_triangulate_points": [pts[n % 64, k] for k in range(pout)],
                         "hrnet_triangulate_keypoint_scores": [np.ones(J)] * pout, "hrnet_triangulate_person_scores": [1.0] * pout}
                prev = ref.Human_Triangulation_Smooth(frame, prev, 2.5, 0.75, 0.0, 1 / 30)
                n += 1
            dt = time.perf_counter() - t0
            res["Human_Triangulation_Smooth"] = {"value": n * pout * J / dt, "unit": "keypoints/s", "cores": 1, "kind": "reference",
                                                 "sample": f"{n} frames x {pout} person(s) x {J} joints in {dt:.1f} s, the real reference frame by frame"}
    except Exception as e:
        res["Human_Triangulation_Smooth"] = {"error": str(e)[:160]}
    try:
        from oracle import c_oracle
        Fc = 20000
        pts = rng.uniform(-1, 1, (Fc, pout, J, 3))
        t0 = time.perf_counter()
        c_oracle.smooth(pts, np.full(Fc, pout, np.int32), 2.5, 0.75, 0.0, 1 / 30)
        dt = time.perf_counter() - t0
        res["smooth_c_port"] = {"value": Fc * pout * J / dt, "unit": "keypoints/s", "cores": 1, "kind": "port",
                                "sample": f"{Fc} frames in {dt:.2f} s, C restatement (oracle/snow_oracle.c)"}
    except Exception as e:
        res["smooth_c_port"] = {"error": str(e)[:160]}
    try:
        from oracle import blender_oracle as bo
        rows = (rng.random((256, 133, 3)) * 2.0).astype(np.float32).astype(np.float64)
        n, t0 = 0, time.perf_counter()
        with np.errstate(all="ignore"):
            while time.perf_counter() - t0 < seconds:
                bo.control_points(rows[n % 256])
                n += 1
        dt = time.perf_counter() - t0
        res["blender_control_points"] = {"value": n / dt, "unit": "persons/s", "cores": 1, "kind": "port",
                                         "sample": f"{n} person rows in {dt:.1f} s, NumPy restatement of blender.py:93-143"}
    except Exception as e:
        res["blender_control_points"] = {"error": str(e)[:160]}
    return res


def traffic_of(key):
    """Measured DRAM bytes per launch of the dominant kernel from the committed ncu capture, or None when the kernel's
    source changed since (the capture records the sha256 of the source file it profiled)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            entry = json.load(fh).get(key)
        if not isinstance(entry, dict):
            return None, "no capture with provenance for this kernel"
        h = hashlib.sha256()
        for src in entry["source"]:
            with open(os.path.join(ROOT, src), "rb") as fh:
                h.update(fh.read())
        if h.hexdigest()[:16] != entry["source_sha16"]:
            return None, f"stale: {', '.join(entry['source'])} changed since the capture at {entry['commit']}"
        return entry["bytes"], f"ncu --set full at commit {entry['commit']} ({entry['profile']})"
    except Exception as e:
        return None, f"unavailable: {e}"


def checksum(torch, t):
    """Order-sensitive 64-bit checksum of a tensor's bytes (int32 words, weighted by position mod a prime)."""
    w = t.contiguous().view(torch.int32).flatten().to(torch.int64)
    idx = torch.arange(w.numel(), device=w.device, dtype=torch.int64) % 1000003 + 1
    return int(((w * idx).sum() & 0x7FFFFFFFFFFFFFFF).item())


def strong_block(torch, dist, name, args, rank, world, local, dev, steps):
    """A fixed clip sharded over the ranks, the all-gather of the 3D joints inside the timed step."""
    from snowmocap_b200 import synth
    from snowmocap_b200.dist import init_native_comm, shard_cyclic, triangulate_cyclic_overlapped
    from snowmocap_b200.engine import TriangulationEngine
    wl, Ft, pieces_n, desc = STRONG[name]
    rig_kind, C, P, J, _, pk, pout, _ = WORKLOADS[wl]
    rig, prm = load_rig(rig_kind, C), params_of(pk)
    Ft -= Ft % (pieces_n * 8)                       # divisible for every N in 1, 2, 4, 8
    eng = TriangulationEngine(rig.K, rig.R, rig.t, device=local, precision="mixed", **prm)
    eng.set_general_kernels(args.general_gen)
    spans = shard_cyclic(Ft, rank, world, pieces_n)
    pieces = []
    for lo, hi in spans:
        kp, sc = synth.make_frames_torch_range(rig, lo, hi, P, J, seed=97, device=dev)
        pieces.append((kp, sc, None))
    c = spans[0][1] - spans[0][0]
    loc_all = new_out(torch, c * len(spans), pout, J, dev)
    loc = [{k2: v[k * c:(k + 1) * c] for k2, v in loc_all.items()} for k in range(len(spans))]
    full = new_out(torch, Ft, pout, J, dev) if world > 1 else None
    comm = torch.cuda.Stream() if world > 1 else None
    if world > 1:
        init_native_comm(eng)

    def step():
        if world > 1:
            triangulate_cyclic_overlapped(eng, pieces, loc_all, full, world, comm, Pout=pout)
        else:
            for k, (kp, sc, _) in enumerate(pieces):
                eng.run(kp, sc, None, Pout=pout, out=loc[k])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    step()
    barrier()
    ev0, ev1 = _events(torch)
    barrier()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    # compute-only time of the same pieces (no gather), to show what the collective costs
    ev0.record()
    for k, (kp, sc, _) in enumerate(pieces):
        eng.run(kp, sc, None, Pout=pout, out=loc[k])
    ev1.record()
    torch.cuda.synchronize()
    tc = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    if world > 1:
        cs = {k: checksum(torch, full[k]) for k in ("out", "nout")}
        persons = float(full["nout"].float().mean().item())
    else:
        cs = {k: checksum(torch, loc_all[k]) for k in ("out", "nout")}
        persons = float(loc_all["nout"].float().mean().item())
    kp_total = Ft * P * J
    tfl = kp_total * flop_per_keypoint(C, P) / (ms * 1e-3) / 1e12
    res = {"description": desc, "workload": wl, "frames_total": Ft, "frames_per_gpu": Ft // world, "pieces_per_gpu": pieces_n,
           "scaling": "strong", "value": kp_total / (ms * 1e-3), "unit": "keypoints/s", "ms_per_step": ms, "steps": steps,
           "compute_only_ms": float(tc.item()), "allgather_in_step": world > 1,
           "allgather_bytes_received_per_gpu": int(Ft * pout * J * 16 * (world - 1) // world) if world > 1 else 0,
           "checksum": cs, "checksum_of": "gathered out (x,y,z,score) and nout of the whole clip: must not depend on the number of GPUs",
           "mean_persons_per_frame": persons, "kernel": eng.last_launch_info()["kernel"],
           "fp32_tflops": tfl, "frac_of_fp32_peak_per_gpu": tfl / world / FP32_PEAK_TFLOPS}
    eng.close()
    del pieces, loc, loc_all, full
    torch.cuda.empty_cache()
    return res


def roofline_block(wl, precision, C, P, J, F, ms, steps, launches, launch_info, peak, peak_src):
    alg_bytes = (12 * C + 16) * P * J * F                     # per output keypoint, per launch
    per_step = launches / steps
    # average duration of the dominant kernel's launch; the streaming general path launches several kernels per
    # step, reported as one step
    kernel_ms = ms / launches if per_step == 1 else ms / steps
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    flops = flop_per_keypoint(C, P) * P * J * F
    tflops = flops / (kernel_ms * 1e-3) / 1e12
    kern = launch_info["kernel"]
    tkey = f"{wl}_{precision}" + ("_jit" if kern == "p1-jit" else "")
    traffic, traffic_src = traffic_of(tkey)
    solves = C * (C - 1) // 2 * P * P * J * F
    names = {"p1": "snowtri::p1_kernel",
             "p1-jit": "p1_jit (snowtri::p1_body specialised for this rig and batch shape with NVRTC)",
             "general": "snowtri::gen_keep + gen_centre + gen_cluster + gen_members + gen_fuse + gen_pscore (one step)",
             "general2": "snowtri::gen_match_smem_kernel + gen_cluster_*_kernel + mfuse_kernel (one step)",
             "general2m": "snowtri::gen_rays + gen_match_global + gen_cluster + gen_members + gen_fuse + gen_pscore (one step)"}
    hbm_bound = P == 1
    return {"bound": "hbm" if hbm_bound else "fp32_issue",
            "achieved": achieved if hbm_bound else tflops, "peak": peak if hbm_bound else FP32_PEAK_TFLOPS,
            "unit": "GB/s" if hbm_bound else "TFLOP/s",
            "frac": achieved / peak if hbm_bound else tflops / FP32_PEAK_TFLOPS,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src if hbm_bound else "nominal: 148 SMs x 128 lanes x 2 x 1965 MHz (no measured FP32 figure in MEASURED_PEAKS.json)",
            "kernel": names.get(kern, "snowtri::fused_kernel"), "kernel_ms": kernel_ms,
            "algorithmic_bytes_per_launch": int(alg_bytes), "algorithmic_flop_per_launch": int(flops),
            "hbm": {"achieved_GBps": achieved, "peak_GBps": peak, "frac": achieved / peak, "frac_of_8TBs_spec": achieved / 8000.0},
            "fp32": {"achieved_TFLOPs": tflops, "peak_TFLOPs": FP32_PEAK_TFLOPS, "frac": tflops / FP32_PEAK_TFLOPS,
                     "flop_per_keypoint": flop_per_keypoint(C, P)},
            "pair_solves_per_sec": solves / (kernel_ms * 1e-3)}


def secondary_cfg3(torch, dist, args, local, dev, numa_cpus):
    """BASELINE configs[2] on one GPU: 8 cameras x 4 persons x 133 keypoints, 10k frames, mixed precision."""
    from snowmocap_b200 import synth
    from snowmocap_b200.engine import TriangulationEngine
    rig_kind, C, P, J, F, pk, pout, desc = WORKLOADS["cfg3"]
    rig, prm = load_rig(rig_kind, C), params_of(pk)
    eng = TriangulationEngine(rig.K, rig.R, rig.t, device=local, precision="mixed", **prm)
    eng.set_general_kernels(args.general_gen)
    kpts, scores = synth.make_frames_torch(rig, F, P, J, seed=4321, device=dev)
    out = new_out(torch, F, pout, J, dev)
    parity = parity_block(torch, eng, rig, prm, kpts, scores, out, pout, C, "mixed")
    parity["reference_frames"] = reference_frames_parity(torch, eng, "big_cfg3_c8p4j133")
    steps = 10
    ms_max, ms, launches = timed_steps(torch, dist, eng, kpts, scores, pout, out, steps, 3, 1, dev)
    info = eng.last_launch_info()
    peak, peak_src = measured_peak()
    res = {"description": desc, "value": F * P * J * steps / (ms_max * 1e-3), "unit": "keypoints/s", "ms_per_step": ms_max / steps,
           "steps": steps, "warmup": 3, "dtype": "mixed", "frames": F, "Pout": pout, "thresholds": prm,
           "gpu_launches": int(launches), "launches_per_step": launches / steps, "launch": info, "parity": parity,
           "roofline": roofline_block("cfg3", "mixed", C, P, J, F, ms, steps, launches, info, peak, peak_src)}
    if not args.no_e2e:
        res["e2e"] = e2e_block(torch, dist, eng, kpts, scores, out, pout, F * P * J, 5, 1, dev, numa_cpus)
    if not args.no_cpu:
        res["cpu_baseline"] = cpu_baseline_block(rig, C, P, J, prm, pout, budget_s=8.0)
    eng.close()
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
                    help="auto = f32 for one person per camera (every output bounded in `parity`), mixed for several")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to its GPU's NUMA node")
    ap.add_argument("--no-others", action="store_true", help="skip the other precision modes and the downstream stages")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary blocks (cfg3, strong scaling)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling blocks only")
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
    all_cpus = sorted(os.sched_getaffinity(0))
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
        args.precision = "f32" if P == 1 else "mixed"
    eng = TriangulationEngine(rig.K, rig.R, rig.t, device=local, precision=args.precision, **prm)
    eng.set_jit(args.jit)
    eng.set_general_kernels(args.general_gen)
    kpts, scores = synth.make_frames_torch(rig, F, P, J, seed=1234 + rank, device=dev)
    out = new_out(torch, F, pout, J, dev)
    in_bytes = kpts.numel() * 4 + scores.numel() * 4
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    parity = parity_block(torch, eng, rig, prm, kpts, scores, out, pout, C, args.precision) if rank == 0 else None

    # ---- device-resident timing -------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    ms_max, ms, launches = timed_steps(torch, dist, eng, kpts, scores, pout, out, steps, warmup, world, dev, nvtx="timed")
    launch_info = eng.last_launch_info()
    per_rank_timing = dict(timed_steps.per_rank)
    if rank == 0:
        time.sleep(0.1)
    clocks = sampler.stop() if rank == 0 else None
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
            oms_max, _, _ = timed_steps(torch, dist, eng, kpts, scores, pout, out, 10, 3, 1, dev)
            oms = oms_max / 10
            got = out["out"][:nchk].cpu().numpy().astype(np.float64)
            ks, kr = got[mm][..., 3], ref["kscores"][mm]
            nz = kr != 0
            rel = np.abs(ks[nz] - kr[nz]) / kr[nz] if nz.any() else np.zeros(1)
            others[prec] = {"ms_per_step": oms, "value": F * P * J / (oms * 1e-3),
                            "roofline_frac": (12 * C + 16) * P * J * F / (oms * 1e-3) / 1e9 / measured_peak()[0],
                            "rel_l2_points": float(np.linalg.norm(got[mm][..., :3] - ref["points"][mm]) / np.linalg.norm(ref["points"][mm])),
                            "median_rel_err_kscores": float(np.median(rel)), "p999_rel_err_kscores": float(np.quantile(rel, 0.999)),
                            "max_rel_err_kscores": float(rel.max())}
        eng.set_precision(args.precision)

    # ---- end to end through host buffers ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = e2e_block(torch, dist, eng, kpts, scores, out, pout, kp_per_step, steps, world, dev, numa_cpus)
        if rank == 0 and world == 1 and not args.no_others:
            try:
                e2e["dropin"] = dropin_leg(rig, prm, P, J)
            except Exception as e:
                e2e["dropin"] = {"error": str(e)[:200]}

    # ---- N > 1: the all-gather of the 3D joints inside the step (pieces, gather k overlapped with compute k+1) -----
    gather = None
    value_with_gather = None
    if world > 1:
        from snowmocap_b200.dist import init_native_comm, triangulate_cyclic_overlapped
        try:
            init_native_comm(eng)
            npieces = 4
            c = F // npieces
            pieces = [(kpts[k * c:(k + 1) * c], scores[k * c:(k + 1) * c], None) for k in range(npieces)]
            loc = {k2: v[:c * npieces] for k2, v in out.items()}
            full = new_out(torch, c * npieces * world, pout, J, dev)
            comm = torch.cuda.Stream()
            for _ in range(2):
                triangulate_cyclic_overlapped(eng, pieces, loc, full, world, comm, Pout=pout)
            dist.barrier()
            torch.cuda.synchronize()
            g0, g1 = _events(torch)
            g0.record()
            for _ in range(steps):
                triangulate_cyclic_overlapped(eng, pieces, loc, full, world, comm, Pout=pout)
            g1.record()
            dist.barrier()
            torch.cuda.synchronize()
            gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
            dist.all_reduce(gms, op=dist.ReduceOp.MAX)
            gms = float(gms.item()) / steps
            value_with_gather = c * npieces * world * P * J / (gms * 1e-3)
            gather = {"ms_per_step_with_allgather": gms, "ms_per_step_kernel_only": ms_max / steps, "pieces": npieces,
                      "bytes_received_per_gpu_per_step": int(full["out"].numel() * 4 * (world - 1) // world),
                      "inbound_GBps_per_gpu": full["out"].numel() * 4 * (world - 1) / world / (gms * 1e-3) / 1e9,
                      "api": "snowtri_allgather (ncclAllGather on the handle's communicator, second stream)",
                      "note": "this workload writes 16 B per keypoint at HBM speed and every GPU must receive (N-1)/N of "
                              "the clip over one NVLink port: the gather binds, not the kernel"}
            del full, pieces, loc
        except Exception as e:   # reported, not fatal
            gather = {"error": str(e)[:300]}

    # ---- the rest of main.py's per-frame body on the same batch, device-resident (not part of a step) ---------
    downstream = None
    if rank == 0 and world == 1 and J >= 130 and not args.no_others:
        try:
            downstream = time_downstream(torch, eng, out, pout, J)
            if not args.no_cpu:
                downstream["cpu_baseline"] = downstream_cpu_baseline(pout, J)
        except Exception as e:   # reported, not fatal
            downstream = {"error": str(e)[:200]}

    peak, peak_src = measured_peak()
    roofline = roofline_block(wl, args.precision, C, P, J, F, ms, steps, launches, launch_info, peak, peak_src)
    jit_status = eng.jit_status
    cpu_main = None
    if rank == 0 and not args.no_cpu and world == 1:      # CPU baseline legs: rank 0 at N=1 only
        cpu_main = cpu_baseline_block(rig, C, P, J, prm, pout)

    # ---- secondary blocks ----------------------------------------------------------------------------------------
    secondary = {}
    if not args.no_secondary and wl == "cfg2":
        del kpts, scores, out
        eng.close()
        torch.cuda.empty_cache()
        if world == 1:
            try:
                secondary["cfg3"] = secondary_cfg3(torch, dist, args, local, dev, numa_cpus)
            except Exception as e:
                secondary["cfg3"] = {"error": str(e)[:300]}
        if not args.no_strong:
            for name in STRONG:
                try:
                    secondary[name] = strong_block(torch, dist, name, args, rank, world, local, dev, steps=2)
                except Exception as e:
                    secondary[name] = {"error": str(e)[:300]}

    if rank == 0:
        line = {"metric": "3d_keypoints_per_sec", "value": value, "unit": "keypoints/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": workload_config(wl, F), "launch": launch_info,
                "gpu_launches": int(launches), "timed_region_per_rank_ms": {**per_rank_timing, "what": f"the {steps} timed steps: CUDA-event time and host time to enqueue them, per rank"}, "jit": jit_status, "clocks": clocks, "parity": parity, "e2e": e2e,
                "roofline": roofline, "other_precisions": others, "allgather": gather, "downstream": downstream,
                "secondary": secondary or None, "host": {"cpus_visible": len(all_cpus), "cpu_count": os.cpu_count()}}
        if value_with_gather is not None:
            line["value_with_allgather"] = value_with_gather
        if cpu_main is not None:
            line["cpu_baseline"] = cpu_main
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
