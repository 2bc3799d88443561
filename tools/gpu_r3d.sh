#!/bin/bash
# clustering by one warp per frame up to N candidates (SNOWTRI_CLUSTER_WARP_MAX): cfg3 bench lines
tag=${1:-r3d}; out=gpurun_out/$tag; mkdir -p $out
for m in 64 512 64 512; do
  SNOWTRI_CLUSTER_WARP_MAX=$m timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/bench_cfg3_w$m.json 2> $out/bench_cfg3_w$m.err
  python -c "
import json; d=json.load(open('$out/bench_cfg3_w$m.json')); print('warp_max $m', 'value=%.4e'%d['value'], 'ms=%.4f'%d['ms_per_step'], d['parity']['nout_equal'], '%.2e'%d['parity']['rel_l2_points'])" || tail -3 $out/bench_cfg3_w$m.err
done
SNOWTRI_CLUSTER_WARP_MAX=512 timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_cfg3_w512.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_launches_cfg3.log 2>&1
grep -v "^==" $out/launches_cfg3_w512.csv | awk -F'","' 'NR>1{print $7, $NF}' | cut -c1-100 | tail -3
