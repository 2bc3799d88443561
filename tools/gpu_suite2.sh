#!/bin/bash
# Round-2 GPU-box suite: parity tests, smoke, the default bench line (both arms), launch lists of the timed regions and
# full ncu captures of the dominant kernels (single-person kernel in both float modes; matching and fuse kernels of cfg3).
# Usage (from the repo root, through gpurun): bash tools/gpu_suite2.sh [tag]
tag=${1:-r3z}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi > $out/nvidia-smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
t0=$SECONDS
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "default bench wall $((SECONDS-t0)) s" | tee $out/bench_default_wall.txt
t0=$SECONDS
timeout 900 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference arm wall $((SECONDS-t0)) s" | tee -a $out/bench_default_wall.txt
timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_cfg2.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_launches_cfg2.log 2>&1
timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_launches_cfg3.log 2>&1
for prec in f32 mixed; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:p1_jit -s 4 -c 1 -f -o $out/p1_jit_cfg2_$prec \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary --precision $prec > $out/ncu_full_cfg2_$prec.log 2>&1
done
for k in gen_match_smem_kernel gen_cluster_warp_kernel mfuse_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $out/${k}_cfg3_mixed \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_full_cfg3_$k.log 2>&1
done
for k in smooth_overlap_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/${k}_f32 \
    python tools/smooth_bench.py > $out/ncu_full_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_smooth.csv python tools/smooth_bench.py > /dev/null 2>&1
timeout 120 python tools/smooth_bench.py > $out/smooth_bench.json 2> $out/smooth_bench.err
timeout 300 python tools/blender_bench.py > $out/blender_bench.json 2> $out/blender_bench.err
tail -3 $out/pytest_gpu.log; tail -4 $out/smoke.log; cut -c1-600 $out/bench_default.json; echo; cut -c1-400 $out/bench_reference.json
