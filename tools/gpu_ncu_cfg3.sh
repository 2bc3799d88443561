#!/bin/bash
# full ncu captures of the three kernels of a cfg3 step.  Usage: tools/gpu_ncu_cfg3.sh tag [kernels]
tag=${1:-ncu3}; out=gpurun_out/$tag; mkdir -p $out
for k in ${2:-gen_match_smem_kernel mfuse_kernel}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $out/${k}_cfg3_mixed \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_full_cfg3_$k.log 2>&1
done
ls -la $out
