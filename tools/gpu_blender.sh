#!/bin/bash
# Blender control-point stage on the GPU box: parity tests, timings, launch list and one full ncu capture.
# Usage (through gpurun): bash tools/gpu_blender.sh [tag]
tag=${1:-blender}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_blender.py -m gpu -x -q > $out/pytest_blender.log 2>&1; echo "pytest rc=$?" >> $out/pytest_blender.log
tail -15 $out/pytest_blender.log
timeout 300 python tools/blender_bench.py > $out/blender_bench.json 2> $out/blender_bench.err; cat $out/blender_bench.json; tail -3 $out/blender_bench.err
BLENDER_BENCH_P=8 timeout 300 python tools/blender_bench.py > $out/blender_bench_p8.json 2> $out/blender_bench_p8.err; cat $out/blender_bench_p8.json
# staging variants of blender_kernel (registers per thread cap x row loads in flight), rebuilt on the box
for v in "6 8" "6 16" "6 32" "8 16" "8 32"; do set -- $v
  SNOWTRI_NVCC_FLAGS="-DBLENDER_MINB=$1 -DBLENDER_UNROLL=$2" python -c "from snowmocap_b200 import build; build.build(only=['snowtri_blender.cu'])" > /dev/null 2>&1
  for p in 1 8; do
    SNOWTRI_NVCC_FLAGS="-DBLENDER_MINB=$1 -DBLENDER_UNROLL=$2" BLENDER_BENCH_NO_SMOOTH=1 BLENDER_BENCH_P=$p timeout 120 python tools/blender_bench.py >> $out/blender_variants.jsonl 2>> $out/blender_variants.err
  done
done
python -c "from snowmocap_b200 import build; build.build(only=['snowtri_blender.cu'])" > /dev/null 2>&1
python - <<PY
import json
for l in open("$out/blender_variants.jsonl"):
    d = json.loads(l); print(d["nvcc_flags"], "P=%d" % d["Pout"], "ms=%.4f" % d["ms"], "frac=%.3f" % d["frac_of_measured_hbm"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:blender -c 40 --csv --log-file $out/launches_blender.csv \
    python tools/blender_bench.py > $out/ncu_launches_blender.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blender_kernel -s 3 -c 1 -f -o $out/blender_kernel_f32 \
    python tools/blender_bench.py > $out/ncu_full_blender.log 2>&1
tail -2 $out/ncu_full_blender.log | cut -c1-200
