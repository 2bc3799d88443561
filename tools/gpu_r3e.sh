#!/bin/bash
# one-pass smoothing (cooperative launch, warm-up per chunk): smoothing / clip tests, timings for several chunk lengths
tag=${1:-r3e}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_clip.py -m gpu -x -q -k "smooth or clip or pipeline" > $out/pytest_smooth.log 2>&1; echo "pytest rc=$?" >> $out/pytest_smooth.log
tail -5 $out/pytest_smooth.log
for L in 256 384 512 1024; do
  echo "chunk $L: $(SNOWTRI_SMOOTH_CHUNK=$L python tools/smooth_bench.py 2>&1 | tail -1 | cut -c1-160)"
done
python tools/smooth_bench.py > $out/smooth_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_smooth.csv python tools/smooth_bench.py > /dev/null 2>&1
grep -v "^==" $out/launches_smooth.csv | awk -F'","' 'NR>1{print $5, $NF}' | cut -c1-120 | tail -3
