#!/bin/bash
# GPU-box suite: parity tests, bench lines, ncu launch list and one full capture of the fused kernel.
# Usage (from the repo root, through gpurun): bash tools/gpu_suite.sh [tag]
tag=${1:-r1}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi > $out/nvidia-smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_cfg2_f64.json 2> $out/bench_cfg2_f64.err
timeout 300 python bench.py --steps 20 --warmup 5 --precision f32 --no-cpu > $out/bench_cfg2_f32.json 2> $out/bench_cfg2_f32.err
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu > $out/bench_cfg3_f64.json 2> $out/bench_cfg3_f64.err
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --precision f32 > $out/bench_cfg3_f32.json 2> $out/bench_cfg3_f32.err
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu > $out/bench_cfg1_f64.json 2> $out/bench_cfg1_f64.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_cfg2.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > $out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 4 -c 1 -f -o $out/fused_cfg2_f64 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $out/ncu_full_cfg2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 4 -c 1 -f -o $out/fused_cfg3_f64 \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e > $out/ncu_full_cfg3.log 2>&1
tail -3 $out/pytest_gpu.log; cat $out/smoke.log | tail -2; cat $out/bench_cfg2_f64.json | cut -c1-600
