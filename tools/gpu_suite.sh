#!/bin/bash
# GPU-box suite: parity tests, smoke, bench lines, ncu launch list and full captures of the dominant kernels.
# Usage (from the repo root, through gpurun): bash tools/gpu_suite.sh [tag]
tag=${1:-r1}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi > $out/nvidia-smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_cfg2.json 2> $out/bench_cfg2.err
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu > $out/bench_cfg3_f64.json 2> $out/bench_cfg3_f64.err
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --precision mixed > $out/bench_cfg3_mixed.json 2> $out/bench_cfg3_mixed.err
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu --no-e2e --precision mixed > $out/bench_cfg4_mixed.json 2> $out/bench_cfg4_mixed.err
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu --no-e2e --precision mixed > $out/bench_cfg5_mixed.json 2> $out/bench_cfg5_mixed.err
timeout 120 python tools/smooth_bench.py > $out/smooth_bench.json 2> $out/smooth_bench.err
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu > $out/bench_cfg1.json 2> $out/bench_cfg1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_cfg2.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-others > $out/ncu_launches.log 2>&1
for prec in f32 mixed f64; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:p1_ -s 4 -c 1 -f -o $out/p1_cfg2_$prec \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --precision $prec > $out/ncu_full_cfg2_$prec.log 2>&1
done
for k in gen_keep gen_fuse; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $out/${k}_cfg3_mixed \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --precision mixed > $out/ncu_full_cfg3_$k.log 2>&1
done
timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_cfg3_mixed.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --precision mixed > $out/ncu_launches_cfg3.log 2>&1
tail -3 $out/pytest_gpu.log; cat $out/smoke.log | tail -2; cat $out/bench_cfg2.json | cut -c1-400
