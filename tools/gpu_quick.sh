#!/bin/bash
# Quick GPU iteration: parity tests then device-resident bench lines (no CPU legs).
tag=${1:-q}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
for prec in f32 mixed f64; do
  timeout 300 python bench.py --steps 20 --warmup 5 --precision $prec --no-cpu --no-e2e > $out/bench_cfg2_$prec.json 2> $out/bench_cfg2_$prec.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_cfg2_$prec.json"))
    print("$prec", "value=%.3e"%d["value"], "ms=%.4f"%d["ms_per_step"], "frac=%.3f"%d["roofline"]["frac"], d["parity"], d["config"]["launch"])
except Exception as e:
    print("$prec bench failed", e); print(open("$out/bench_cfg2_$prec.err").read()[-2000:])
PY
done
