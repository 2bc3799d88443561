#!/bin/bash
# quick pass: general-path tests, cfg3 bench + launch list + ncu of the matching kernel
tag=${1:-r2d}; out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "second_generation or full_size or large_rigs or random_configurations or float_modes or golden" > $out/pytest_general.log 2>&1; echo "pytest rc=$?" >> $out/pytest_general.log
tail -5 $out/pytest_general.log
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/bench_cfg3.json 2> $out/bench_cfg3.err
python - <<PY
import json
try:
    d=json.load(open("$out/bench_cfg3.json"))
    print("cfg3", "value=%.3e"%d["value"], "ms=%.4f"%d["ms_per_step"], d["launch"]["kernel"], "launches", d["gpu_launches"], d["roofline"]["frac"], d["parity"]["nout_equal"], d["parity"]["rel_l2_points"])
except Exception as e:
    print("cfg3 bench failed", e); print(open("$out/bench_cfg3.err").read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $out/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/launches_cfg3.log 2>&1
grep -v "^==" $out/launches_cfg3.csv | awk -F'","' 'NR>1{print $7, $NF}' | head -3
for kre in gen_match_smem_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -f -o $out/${kre}_cfg3_mixed \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_$kre.log 2>&1
  tail -1 $out/ncu_$kre.log | cut -c1-200
done
