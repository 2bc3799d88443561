#!/bin/bash
# Round 2, first GPU pass: the new general-path kernels.  Usage: tools/gpu_r2a.sh [tag]
tag=${1:-r2a}; out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/nvidia-smi.txt 2>&1
# 1. the new tests first (fail fast), then everything
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "second_generation or full_size or large_rigs" > $out/pytest_new.log 2>&1
echo "pytest new rc=$?" >> $out/pytest_new.log; tail -25 $out/pytest_new.log
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -12 $out/pytest_gpu.log
# 2. cfg3/4/5 bench lines, both generations
for wl in cfg3 cfg4 cfg5; do
  for gen in 2 1; do
    timeout 300 python bench.py --workload $wl --precision mixed --general-gen $gen --steps 10 --warmup 3 --no-cpu --no-e2e --no-others \
        > $out/bench_${wl}_gen$gen.json 2> $out/bench_${wl}_gen$gen.err
    python - <<PY
import json
try:
    d=json.load(open("$out/bench_${wl}_gen$gen.json"))
    print("$wl gen$gen", "value=%.3e"%d["value"], "ms=%.4f"%d["ms_per_step"], d["parity"], d["launch"]["kernel"], "launches", d["gpu_launches"])
except Exception as e:
    print("$wl gen$gen bench failed", e); print(open("$out/bench_${wl}_gen$gen.err").read()[-1500:])
PY
  done
done
# 3. launch list of cfg3 (timed region only) and one full capture of each new kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $out/launches_cfg3_mixed.csv \
    python bench.py --workload cfg3 --precision mixed --steps 2 --warmup 3 --no-cpu --no-e2e --no-others > $out/launches_cfg3.log 2>&1
grep -v "^==" $out/launches_cfg3_mixed.csv | awk -F'","' 'NR>1{print $7, $NF}' | head -12
for kre in gen_match_smem_kernel mfuse_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -f -o $out/${kre}_cfg3_mixed \
    python bench.py --workload cfg3 --precision mixed --steps 2 --warmup 3 --no-cpu --no-e2e --no-others > $out/ncu_$kre.log 2>&1
  tail -2 $out/ncu_$kre.log | cut -c1-200
done
ls -la $out
