#!/bin/bash
# Matching kernel, second form of the evaluation (sums + joint-major staging): general-path parity tests, then
# cfg3 / cfg4 / cfg5 bench lines of the library and of the variants under snowmocap_b200/variants/ (A/B on one box).
tag=${1:-r3b}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "general or large_rigs or full_size or random_config or fused or property or empty or golden" > $out/pytest_general.log 2>&1; echo "pytest rc=$?" >> $out/pytest_general.log
tail -4 $out/pytest_general.log
bench() {  # name
  for wl in cfg3 cfg4 cfg5; do
    timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/bench_${wl}_$1.json 2> $out/bench_${wl}_$1.err
    python - <<PY
import json
try:
    d=json.load(open("$out/bench_${wl}_$1.json"))
    print("$1 $wl", "value=%.4e"%d["value"], "ms=%.4f"%d["ms_per_step"], "relL2=%.2e"%d["parity"]["rel_l2_points"], "nout_equal", d["parity"]["nout_equal"])
except Exception as e:
    print("$1 $wl bench failed", e); print(open("$out/bench_${wl}_$1.err").read()[-1500:])
PY
  done
}
bench new
cp snowmocap_b200/libsnowtri.so /tmp/libsnowtri.orig.so
for so in snowmocap_b200/variants/*.so; do
  [ -f "$so" ] || continue
  cp $so snowmocap_b200/libsnowtri.so
  bench $(basename $so .so)
done
cp /tmp/libsnowtri.orig.so snowmocap_b200/libsnowtri.so
timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_launches_cfg3.log 2>&1
grep -v "^==" $out/launches_cfg3.csv | awk -F'","' 'NR>1{print $7, $NF}' | cut -c1-100 | tail -3
