#!/bin/bash
# three-stream host pipeline: tests of the host entry points, e2e of cfg2 and cfg3
tag=${1:-r2q}; out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host or pipeline or ragged or dropin" > $out/pytest_host.log 2>&1; echo "pytest rc=$?" >> $out/pytest_host.log
tail -3 $out/pytest_host.log
for wl in cfg2 cfg3; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-others --no-secondary > $out/bench_$wl.json 2> $out/bench_$wl.err
  python -c "
import json; d=json.load(open('$out/bench_$wl.json')); e=d['e2e']; print('$wl value=%.3e e2e=%.3e ceiling=%.3e frac=%.3f'%(d['value'], e['value'], e['copy_only_ceiling']['value'], e['frac_of_copy_ceiling']))" || tail -5 $out/bench_$wl.err
done
