#!/bin/bash
# pipelined persistent matching kernel: general-path tests, cfg3 with and without it, launch list, ncu capture
tag=${1:-r2o}; out=gpurun_out/$tag
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "general or multi or fused or golden or random or big or full" > $out/pytest_general.log 2>&1; echo "pytest rc=$?" >> $out/pytest_general.log
tail -3 $out/pytest_general.log
for v in pipe; do
  if [ $v = nopipe ]; then export SNOWTRI_MATCH_PIPE=0; else unset SNOWTRI_MATCH_PIPE; fi
  timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/bench_cfg3_$v.json 2> $out/bench_cfg3_$v.err
  python -c "
import json; d=json.load(open('$out/bench_cfg3_$v.json')); print('cfg3 $v value=%.3e ms=%.4f'%(d['value'], d['ms_per_step']), d['parity']['nout_equal'], d['parity']['rel_l2_points'], d['roofline']['frac'])" || tail -5 $out/bench_cfg3_$v.err
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $out/launches_cfg3_$v.csv \
    python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > /dev/null 2>&1
  grep -v "^==" $out/launches_cfg3_$v.csv | awk -F'","' 'NR>1{print $7, $NF}' | cut -c1-120 | head -5
done
unset SNOWTRI_MATCH_PIPE
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gen_match_smem -s 3 -c 1 -f -o $out/gen_match_smem_kernel_cfg3_mixed \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_match.log 2>&1
tail -1 $out/ncu_match.log | cut -c1-150
