#!/bin/bash
# One full ncu capture of the dominant kernel per precision.  Usage: tools/gpu_ncu.sh tag "f32 mixed" [workload] [kernel-regex]
tag=${1:-n}; precs=${2:-f32}; wl=${3:-cfg2}; kre=${4:-p1_kernel}
out=gpurun_out/$tag
mkdir -p $out
for prec in $precs; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kre -s 4 -c 1 -f -o $out/${kre}_${wl}_$prec \
    python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --no-e2e --precision $prec > $out/ncu_${wl}_$prec.log 2>&1
  tail -2 $out/ncu_${wl}_$prec.log | cut -c1-200
done
