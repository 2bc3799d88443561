#!/bin/bash
# final check of the smoothing stages: all Blender / smoothing / clip tests (one-pass path with the shipped-like fast followers)
tag=${1:-r3k}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_blender.py tests/test_zz_clip.py -m gpu -x -q -k "smooth or clip or pipeline or blender or Blender" > $out/pytest_smooth.log 2>&1; echo "pytest rc=$?" >> $out/pytest_smooth.log
tail -4 $out/pytest_smooth.log
python tools/smooth_bench.py > $out/smooth_bench.json; cut -c1-110 $out/smooth_bench.json
