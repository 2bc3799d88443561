#!/usr/bin/env python
"""Record the measured DRAM traffic of a kernel (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
launch) in profiles/roofline_traffic.json, with the commit and the sha256 of the kernel sources it was captured from;
bench.py reports the figure as `roofline.traffic` only while those sources are unchanged.

Usage: python tools/traffic_from_ncu.py KEY PROFILE_NAME SOURCE[,SOURCE..] REP [REP ...]
  KEY           "<workload>_<precision>[_jit]" as bench.py looks it up (e.g. cfg2_f32_jit, cfg3_mixed)
  PROFILE_NAME  where the summaries of these captures are committed (e.g. profiles/r2f)
  SOURCE        kernel source files relative to the repo root (sha256 over their concatenation)
  REP           one .ncu-rep per kernel of the step (their traffic is added up)"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_sha16(sources):
    h = hashlib.sha256()
    for s in sources:
        with open(os.path.join(ROOT, s), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def main():
    key, profile, sources, reps = sys.argv[1], sys.argv[2], sys.argv[3].split(","), sys.argv[4:]
    total, kernels = 0.0, []
    for rep in reps:
        rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
            b += float(vals[i]) * scale
        total += b
        kernels.append({"kernel": vals[hdr.index("Kernel Name")], "bytes": int(b)})
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    data[key] = {"bytes": int(total), "kernels": kernels, "commit": commit, "profile": profile,
                 "source": sources, "source_sha16": source_sha16(sources)}
    with open(path, "w") as fh:
        json.dump(data, fh, indent=1)
    print(key, data[key])


if __name__ == "__main__":
    main()
