#!/bin/bash
# the two bench arms exactly as the driver runs them, on the committed code
tag=${1:-benchcheck}; out=gpurun_out/$tag; mkdir -p $out
t0=$SECONDS
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "default bench rc=$? wall $((SECONDS-t0)) s" | tee $out/wall.txt
t0=$SECONDS
timeout 900 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference arm rc=$? wall $((SECONDS-t0)) s" | tee -a $out/wall.txt
cut -c1-300 $out/bench_default.json; echo; cut -c1-200 $out/bench_reference.json; tail -3 $out/bench_default.err
