#!/bin/bash
# scan kernels for the chunk hand-over of both smoothing stages: tests, timings, launch lists
tag=${1:-r2p}; out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_blender.py tests/test_zz_clip.py -m gpu -x -q > $out/pytest_blender.log 2>&1; echo "pytest rc=$?" >> $out/pytest_blender.log
tail -3 $out/pytest_blender.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "smooth or ragged or pipeline or clip" > $out/pytest_smooth.log 2>&1; echo "pytest rc=$?" >> $out/pytest_smooth.log
tail -3 $out/pytest_smooth.log
python tools/smooth_bench.py > $out/smooth_bench.json; cat $out/smooth_bench.json
timeout 300 python tools/blender_bench.py > $out/blender_bench.json 2> $out/blender_bench.err; cat $out/blender_bench.json; tail -3 $out/blender_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_smooth.csv python tools/smooth_bench.py > /dev/null 2>&1
grep -v "^==" $out/launches_smooth.csv | awk -F'","' 'NR>1{print $5, $NF}' | cut -c1-120 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:blender -c 40 --csv --log-file $out/launches_blender.csv python tools/blender_bench.py > /dev/null 2>&1
grep -v "^==" $out/launches_blender.csv | awk -F'","' 'NR>1{print $5, $NF}' | cut -c1-120 | tail -6
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-secondary > $out/bench_cfg2.json 2> $out/bench_cfg2.err
python -c "
import json; d=json.load(open('$out/bench_cfg2.json')); print(d['ms_per_step'], d['downstream'])"
