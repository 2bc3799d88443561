#!/bin/bash
# Round 2, second GPU pass: full test suite, cfg3 A/B (rolled vs unrolled fuse), launch list, ncu of the two kernels.
tag=${1:-r2b}; out=gpurun_out/$tag
mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -12 $out/pytest_gpu.log
for rolled in 0; do
  SNOWTRI_MF_ROLLED=$rolled timeout 300 python bench.py --workload cfg3 --precision mixed --steps 10 --warmup 3 --no-cpu --no-e2e --no-others \
      > $out/bench_cfg3_rolled$rolled.json 2> $out/bench_cfg3_rolled$rolled.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_cfg3_rolled$rolled.json"))
    print("cfg3 rolled=$rolled", "value=%.3e"%d["value"], "ms=%.4f"%d["ms_per_step"], d["parity"], d["launch"]["kernel"], "launches", d["gpu_launches"])
except Exception as e:
    print("cfg3 bench failed", e); print(open("$out/bench_cfg3_rolled$rolled.err").read()[-1500:])
PY
  SNOWTRI_MF_ROLLED=$rolled timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $out/launches_cfg3_rolled$rolled.csv \
      python bench.py --workload cfg3 --precision mixed --steps 2 --warmup 3 --no-cpu --no-e2e --no-others > $out/launches_cfg3.log 2>&1
  grep -v "^==" $out/launches_cfg3_rolled$rolled.csv | awk -F'","' 'NR>1{print $7, $NF}' | head -3
done
for wl in cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl --precision mixed --steps 10 --warmup 3 --no-cpu --no-e2e --no-others > $out/bench_${wl}.json 2> $out/bench_${wl}.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_${wl}.json"))
    print("$wl", "value=%.3e"%d["value"], "ms=%.4f"%d["ms_per_step"], d["parity"], d["launch"]["kernel"], "launches", d["gpu_launches"])
except Exception as e:
    print("$wl bench failed", e); print(open("$out/bench_${wl}.err").read()[-1500:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $out/launches_cfg4.csv \
    python bench.py --workload cfg4 --precision mixed --steps 1 --warmup 3 --no-cpu --no-e2e --no-others > $out/launches_cfg4.log 2>&1
grep -v "^==" $out/launches_cfg4.csv | awk -F'","' 'NR>1{print $7, $NF}' | head -6
for kre in gen_match_smem_kernel mfuse_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -f -o $out/${kre}_cfg3_mixed \
    python bench.py --workload cfg3 --precision mixed --steps 2 --warmup 3 --no-cpu --no-e2e --no-others > $out/ncu_$kre.log 2>&1
  tail -1 $out/ncu_$kre.log | cut -c1-200
done
# 4. the full default bench line (new bench.py), N = 1
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("$out/bench_default.json"))
    print("cfg2", "value=%.3e"%d["value"], "frac=%.3f"%d["roofline"]["frac"], "e2e=%.3e"%d["e2e"]["value"], "ceil frac", d["e2e"].get("frac_of_copy_ceiling"), d["e2e"].get("dropin"))
    print(" parity", d["parity"])
    print(" cpu", json.dumps(d.get("cpu_baseline"))[:600])
    s=d["secondary"]
    for k,v in s.items():
        print(k, json.dumps(v)[:1500])
except Exception as e:
    print("default bench failed", e); print(open("$out/bench_default.err").read()[-3000:])
PY
