#!/bin/bash
# register-resident warp clustering: general-path tests, cfg3 bench + launch list; chunk sweep of the one-pass Blender smoothing
tag=${1:-r3g}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "general or large_rigs or full_size or random_config or fused or property or empty or golden or pack" > $out/pytest_general.log 2>&1; echo "pytest rc=$?" >> $out/pytest_general.log
tail -4 $out/pytest_general.log
for i in 1 2; do
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/bench_cfg3_$i.json 2> $out/bench_cfg3_$i.err
python -c "
import json; d=json.load(open('$out/bench_cfg3_$i.json')); print('cfg3', 'value=%.4e'%d['value'], 'ms=%.4f'%d['ms_per_step'], d['parity']['nout_equal'], '%.2e'%d['parity']['rel_l2_points'], 'frac=%.4f'%d['roofline']['frac'])" || tail -3 $out/bench_cfg3_$i.err
done
timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_launches_cfg3.log 2>&1
grep -v "^==" $out/launches_cfg3.csv | awk -F'","' 'NR>1{print $7, $NF}' | cut -c1-100 | tail -3
for L in 32 64 128 256; do
  echo "bs chunk $L: $(SNOWTRI_BS_CHUNK=$L BLENDER_BENCH_NO_CPU=1 python tools/blender_bench.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: round(v['ms'],4) for k,v in d['smooth'].items() if isinstance(v, dict)})")"
done
