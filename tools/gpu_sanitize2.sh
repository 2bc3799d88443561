#!/bin/bash
# compute-sanitizer over the round-2 kernels: matching (joint-major staging, predicated sums), clustering from registers,
# clique fuse, the one-pass smoothing kernels (cooperative launch).  memcheck, then racecheck on the shared-memory kernels.
out=gpurun_out/${1:-san2}; mkdir -p $out
SEL='second_generation_vs_oracle and (8-4-133-24 or 3-5-33 or 2-9-5) or edge_cases and prm_over0 or smooth_long_clip and batches1-True or batch_smooth_long_clip and fast-batches12'
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_blender.py -m gpu -x -q -k "$SEL" > $out/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $out/memcheck.log
SEL2='second_generation_vs_oracle and (3-5-33 or 2-9-5)'
timeout 700 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL2" > $out/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $out/racecheck.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|deselected" $out/memcheck.log $out/racecheck.log | tail -8
