#!/bin/bash
# matching kernel with the split last chunk: general-path tests, cfg3 / cfg4 / cfg5 bench lines, launch lists; smoothing launch list
tag=${1:-r2n}; out=gpurun_out/$tag
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "general or multi or fused or golden or random or big or full" > $out/pytest_general.log 2>&1; echo "pytest rc=$?" >> $out/pytest_general.log
tail -3 $out/pytest_general.log
for wl in cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/bench_$wl.json 2> $out/bench_$wl.err
  python -c "
import json; d=json.load(open('$out/bench_$wl.json')); print('$wl value=%.3e ms=%.4f'%(d['value'], d['ms_per_step']), d['parity']['nout_equal'], d['parity']['rel_l2_points'], d['roofline']['frac'])" || tail -5 $out/bench_$wl.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $out/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > /dev/null 2>&1
grep -v "^==" $out/launches_cfg3.csv | awk -F'","' 'NR>1{print $7, $NF}' | cut -c1-120 | head -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_smooth.csv python tools/smooth_bench.py > $out/smooth_bench.json 2>&1
grep -v "^==" $out/launches_smooth.csv | awk -F'","' 'NR>1{print $7, $NF}' | cut -c1-120 | tail -8
python tools/smooth_bench.py | tail -1
