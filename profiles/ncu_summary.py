#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) offline: key counters, SASS opcode mix, stall
reasons and the hottest CUDA source lines.  Usage: python profiles/ncu_summary.py rep.ncu-rep [topN]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__dynamic_shared_memory_per_block",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {vals[i]:>16s} {units[i]}")
    # SASS page: opcode mix + stalls
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    h = rows[1]
    ix = {n: i for i, n in enumerate(h)}
    ops, stall, tot = collections.Counter(), collections.Counter(), 0
    scols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    for r in rows[2:]:
        if len(r) < len(h):
            continue
        try:
            n = int(r[ix["Instructions Executed"]])
        except ValueError:
            continue
        toks = r[ix["Source"]].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops[op] += n
        tot += n
        for s in scols:
            try:
                stall[s] += int(r[ix[s]])
            except ValueError:
                pass
    print(f"== SASS opcode mix (total warp instructions {tot})")
    for o, n in ops.most_common(18):
        print(f"  {o:10s} {n:13d} {100 * n / tot:5.1f}%")
    st = sum(stall.values()) or 1
    print("== stall reasons (all samples)")
    for s, n in stall.most_common(8):
        print(f"  {s:25s} {100 * n / st:5.1f}%")
    # CUDA source lines (aggregates of the correlated cuda,sass view have Address == "-")
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    fname, h, lines = "?", None, []
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            fname = r[1]
            continue
        if r and r[0] == "Line No":
            h = {}
            for i, n in enumerate(r):
                h.setdefault(n, i)
            continue
        if h and len(r) > h["Instructions Executed"] and r[h["Address"]] == "-":
            try:
                samples = int(r[h["# Samples"]] or 0)
                inst = int(r[h["Instructions Executed"]] or 0)
            except ValueError:
                continue
            if samples or inst:
                lines.append((samples, inst, fname.split("/")[-1], r[0], r[1].strip()))
    ts = sum(x[0] for x in lines) or 1
    ti = sum(x[1] for x in lines) or 1
    print(f"== hottest CUDA lines by stall samples (total {ts}; warp instr {ti})")
    for s, i, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"  {100 * s / ts:5.1f}% smp {100 * i / ti:5.1f}% ins  {f}:{ln:>4s}  {src[:100]}")


if __name__ == "__main__":
    main()
