#!/bin/bash
# compute-sanitizer over a representative slice of the GPU tests (memcheck, then racecheck on the shared-memory kernels).
out=gpurun_out/${1:-san}; mkdir -p $out
SEL='golden and (fused or dropin) or single_person_kernel_matches_c_oracle and (f32-0 or f64-5 or mixed-7) or smooth_long_clip or pack_detections or fused_matches_c_oracle and (tuning0 or tuning2)'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -x -q -k "$SEL" > $out/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $out/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests -m gpu -x -q -k "golden and fused or single_person_kernel_equals or smooth_batch" > $out/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $out/racecheck.log
grep -E "ERROR SUMMARY|passed|failed" $out/memcheck.log $out/racecheck.log | tail -8
