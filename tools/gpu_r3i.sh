#!/bin/bash
# follower update with folded constants (FP64-pipe-bound smoothing kernels): all smoothing / Blender / clip tests, timings
tag=${1:-r3i}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_blender.py tests/test_zz_clip.py -m gpu -x -q -k "smooth or clip or pipeline or blender or Blender" > $out/pytest_smooth.log 2>&1; echo "pytest rc=$?" >> $out/pytest_smooth.log
tail -4 $out/pytest_smooth.log
for L in 160 192 256 384 512; do
  echo "chunk $L: $(SNOWTRI_SMOOTH_CHUNK=$L python tools/smooth_bench.py 2>&1 | tail -1 | cut -c1-100)"
done
python tools/smooth_bench.py > $out/smooth_bench.json
for L in 64 128; do
  echo "bs chunk $L: $(SNOWTRI_BS_CHUNK=$L BLENDER_BENCH_NO_CPU=1 python tools/blender_bench.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: round(v['ms'],4) for k,v in d['smooth'].items() if isinstance(v, dict)})")"
done
BLENDER_BENCH_NO_CPU=1 python tools/blender_bench.py > $out/blender_bench.json 2> $out/blender_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_smooth.csv python tools/smooth_bench.py > /dev/null 2>&1
grep -v "^==" $out/launches_smooth.csv | awk -F'","' 'NR>1{print $5, $NF}' | cut -c1-120 | tail -2
