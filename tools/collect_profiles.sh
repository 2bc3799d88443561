#!/bin/bash
# Copy the evidence of a gpurun pass from gpurun_out/<tag>/ (scratch) to profiles/<tag>/ (tracked): bench lines, test
# logs, launch lists and a text summary of every ncu capture (profiles/ncu_summary.py); the .ncu-rep files stay out.
for tag in "$@"; do
  src=gpurun_out/$tag; dst=profiles/$tag
  [ -d $src ] || { echo "no $src"; continue; }
  mkdir -p $dst
  for f in $src/*.json $src/*.log $src/*.csv $src/*.txt; do
    [ -f "$f" ] || continue
    case "$f" in *ncu_*.log|*launches_*.log) continue;; esac
    [ $(stat -c %s "$f") -gt 400000 ] && continue
    cp "$f" $dst/
  done
  for rep in $src/*.ncu-rep; do
    [ -f "$rep" ] || continue
    out=$dst/$(basename ${rep%.ncu-rep})_ncu_summary.txt
    [ -f $out ] || python profiles/ncu_summary.py $rep 40 > $out 2>/dev/null
  done
  echo "$tag -> $dst: $(ls $dst | wc -l) files"
done
