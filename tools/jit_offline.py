#!/usr/bin/env python
"""Build the rig-specialised single-person kernel OFFLINE (the same source snowtri_p1.cu hands to NVRTC at run time,
for the shipped 4-camera calibration and the cfg2 batch shape), compile it with nvcc for sm_100a and print the SASS
opcode histogram and the instruction count of the fuse loop.  A static proxy for an issue-bound kernel: use it to
screen variants (extra -D defines as arguments) before spending GPU time.
Usage: python tools/jit_offline.py [P1_NI=1] [OTHER_DEFINE=..]"""
import collections
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = np.load(os.path.join(ROOT, "tests", "golden", "floor_rig.npz"))
K, R, t = z["K"], z["R"], z["t"].reshape(-1, 3)
C, J, Pout, Gw = 4, 133, 1, 32
NP = C * (C - 1) // 2
camc = np.zeros(C * 12, np.float32)
for c in range(C):
    M = R[c] @ np.linalg.inv(K[c])
    for k in range(9):
        camc[12 * c + (k // 3) * 4 + k % 3] = M.reshape(-1)[k]
pdc = np.zeros(NP * 8, np.float32)
e = 0
for x in range(C - 1):
    for y in range(x + 1, C):
        pdc[e * 8:e * 8 + 3] = t[y] - t[x]
        pdc[e * 8 + 4:e * 8 + 7] = (t[x] + t[y]) / 2
        e += 1
inv_dthr = 1.0 / 0.05
defs = {"P1_NI": "2"}
TD = "float"
for arg in sys.argv[1:]:
    k, _, v = arg.partition("=")
    if k == "TD":          # TD=double: the mixed mode (float64 distance numerator)
        TD = v
        defs["P1_NI"] = "1"
        continue
    defs[k] = v or "1"
minb = defs.pop("MINB", "2")
src = "#define P1_JIT 1\n" + "".join(f"#define {k} {v}\n" for k, v in defs.items())
src += f"#define P1_JIT_J {J}\n#define P1_JIT_JOUT {J}\n#define P1_JIT_POUT {Pout}\n#define P1_JIT_GW {Gw}\n"
src += f"#define P1_JIT_KST {0.5:.9e}f\n#define P1_JIT_INV_DTHR {inv_dthr:.9e}f\n#define P1_JIT_GUARD_W {inv_dthr * 1e-3:.9e}f\n"
src += f"#define P1_JIT_KSCALE_FULL {0.0005 / NP:.9e}f\n"
src += "#define P1_JIT_CAMC {" + ",".join(f"{v:.9e}f" for v in camc) + "}\n"
src += "#define P1_JIT_PDC {" + ",".join(f"{v:.9e}f" for v in pdc) + "}\n"
src += ('#include "snowtri_p1.cuh"\nextern "C" __global__ void __launch_bounds__(256, ' + minb + ') p1_jit('
        f"const __grid_constant__ snowtri::P1Args<float, 4> a) {{ snowtri::p1_body<float, {TD}, 4, 256>(a); }}\n")
out = "/tmp/p1_jit_offline"
open(out + ".cu", "w").write(src)
cmd = ["nvcc", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
       "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "snowmocap_b200", "csrc"), "-Xptxas=-v", "-cubin",
       "-o", out + ".cubin", out + ".cu"]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    print(r.stdout + r.stderr)
    sys.exit(1)
print("\n".join(l for l in (r.stdout + r.stderr).splitlines() if "Used" in l or "spill" in l))
sass = subprocess.run(["cuobjdump", "-sass", out + ".cubin"], capture_output=True, text=True).stdout
ins = [m.group(2) for m in re.finditer(r"/\*([0-9a-f]{4})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", sass)]
ops = collections.Counter(i.split(".")[0] for i in ins)
print("total SASS instructions:", len(ins))
print("  ".join(f"{k} {v}" for k, v in ops.most_common(24)))
# the fuse loop = the region between the last two backward branches around the FFMA-dense block: report the densest
# window of 1000 instructions by FFMA count as a stable proxy
ff = [1 if i.startswith(("FFMA", "FMUL", "FADD")) else 0 for i in ins]
print("float arithmetic instructions:", sum(ff), " MUFU:", ops.get("MUFU", 0), " LDG:", ops.get("LDG", 0), " STG:", ops.get("STG", 0))

# loops: every backward branch, with its instruction count and opcode mix (the fuse loop is the big FFMA-dense one)
addr_ops = [(int(m.group(1), 16), m.group(2).split(".")[0], m.group(0)) for m in
            re.finditer(r"/\*([0-9a-f]{4})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)[^;]*;", sass)]
for a, op, text in addr_ops:
    if op == "BRA":
        t = re.search(r"0x([0-9a-f]+)", text.split("BRA")[1])
        if t and int(t.group(1), 16) < a:
            sub = [o for aa, o, _ in addr_ops if int(t.group(1), 16) <= aa <= a]
            if len(sub) > 150:
                print(f"loop {int(t.group(1), 16):#x}..{a:#x}: {len(sub)} instructions per step ({defs['P1_NI']} item(s) per lane):",
                      "  ".join(f"{k} {v}" for k, v in collections.Counter(sub).most_common(16)))
