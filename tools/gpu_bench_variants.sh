#!/bin/bash
# Bench several prebuilt library variants (snowmocap_b200/variants/<name>.so) on cfg2.  Usage: tools/gpu_bench_variants.sh tag "precisions"
tag=${1:-v}; precs=${2:-f32}
out=gpurun_out/$tag; mkdir -p $out
cp snowmocap_b200/libsnowtri.so /tmp/libsnowtri.orig.so
for so in snowmocap_b200/variants/*.so; do
  name=$(basename $so .so)
  cp $so snowmocap_b200/libsnowtri.so
  for prec in $precs; do
    timeout 300 python bench.py --steps 20 --warmup 5 --precision $prec --no-cpu --no-e2e > $out/bench_${name}_$prec.json 2> $out/bench_${name}_$prec.err
    python - <<PY
import json
try:
    d=json.load(open("$out/bench_${name}_$prec.json"))
    print("$name $prec", "value=%.3e"%d["value"], "ms=%.4f"%d["ms_per_step"], "frac=%.3f"%d["roofline"]["frac"], "relL2=%.2e"%d["parity"]["rel_l2_points"], d["launch"])
except Exception as e:
    print("$name $prec bench failed", e); print(open("$out/bench_${name}_$prec.err").read()[-1500:])
PY
  done
done
cp /tmp/libsnowtri.orig.so snowmocap_b200/libsnowtri.so
