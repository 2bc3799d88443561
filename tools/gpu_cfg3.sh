#!/bin/bash
# General-kernel iteration: tests that exercise fused_kernel, then cfg3 bench lines.
out=gpurun_out/${1:-c3}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "not smooth and not pack" > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
for prec in ${2:-f64 f32x}; do
  timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --precision $prec --no-cpu --no-e2e > $out/bench_cfg3_$prec.json 2> $out/bench_cfg3_$prec.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_cfg3_$prec.json"))
    print("$prec", "value=%.3e"%d["value"], "ms=%.4f"%d["ms_per_step"], d["parity"], d["launch"])
except Exception as e:
    print("$prec bench failed", e); print(open("$out/bench_cfg3_$prec.err").read()[-2000:])
PY
done
