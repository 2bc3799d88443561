#!/bin/bash
# cfg3 knob sweep: fuse tile size / CTA shape, matching warps per CTA
tag=${1:-r2g}; out=gpurun_out/$tag
mkdir -p $out
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/bench_cfg3_$name.json 2> $out/bench_cfg3_$name.err
  env "$@" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $out/launches_$name.csv \
      python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > /dev/null 2>&1
  python - <<PY
import json,csv
try:
    d=json.load(open("$out/bench_cfg3_$name.json"))
    ks=[r for r in csv.reader(l for l in open("$out/launches_$name.csv") if not l.startswith("=="))][1:]
    print("$name", "value=%.3e"%d["value"], "ms=%.4f"%d["ms_per_step"], d["parity"]["nout_equal"], " ".join("%s=%.0fus"%(r[6].split("(")[0].split("<")[0][-22:], float(r[-1])/1e3) for r in ks[:3]))
except Exception as e:
    print("$name failed", e); print(open("$out/bench_cfg3_$name.err").read()[-600:])
PY
}
run base A=1
run gw1 SNOWTRI_MF_GW=1
run gw4 SNOWTRI_MF_GW=4
run nt128 SNOWTRI_MF_NT=128
run nt128gw1 SNOWTRI_MF_NT=128 SNOWTRI_MF_GW=1
run mnw8 SNOWTRI_MATCH_NW=8
run mnw6 SNOWTRI_MATCH_NW=6
run mnw4 SNOWTRI_MATCH_NW=4
