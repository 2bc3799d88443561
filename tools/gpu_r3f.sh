#!/bin/bash
# one-pass Blender control-point smoothing: tests, timings for several chunk lengths
tag=${1:-r3f}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_blender.py tests/test_zz_clip.py -m gpu -x -q > $out/pytest_blender.log 2>&1; echo "pytest rc=$?" >> $out/pytest_blender.log
tail -5 $out/pytest_blender.log
for L in 32 64 128 256; do
  echo "chunk $L: $(SNOWTRI_BS_CHUNK=$L BLENDER_BENCH_NO_CPU=1 python tools/blender_bench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: round(v['ms'],4) for k,v in d['smooth'].items()})")"
done
BLENDER_BENCH_NO_CPU=1 python tools/blender_bench.py > $out/blender_bench.json 2> $out/blender_bench.err
