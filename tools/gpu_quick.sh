#!/bin/bash
# Quick GPU iteration: parity tests then bench lines (no CPU legs).  Usage: tools/gpu_quick.sh tag [precisions] [extra bench flags]
tag=${1:-q}; precs=${2:-"f32 mixed f64"}; extra=${3:---no-e2e}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
for prec in $precs; do
  timeout 300 python bench.py --steps 20 --warmup 5 --precision $prec --no-cpu $extra > $out/bench_cfg2_$prec.json 2> $out/bench_cfg2_$prec.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_cfg2_$prec.json"))
    print("$prec", "value=%.3e"%d["value"], "ms=%.4f"%d["ms_per_step"], "frac=%.3f"%d["roofline"]["frac"], d["parity"], d["launch"], "e2e=", d.get("e2e") and "%.3e"%d["e2e"]["value"])
except Exception as e:
    print("$prec bench failed", e); print(open("$out/bench_cfg2_$prec.err").read()[-2000:])
PY
done
