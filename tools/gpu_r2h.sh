#!/bin/bash
# lean p1 loop (magic division, per-tile clique vote, person score read back): tests, cfg2 in all modes, ncu, cfg3 with the new defaults
tag=${1:-r2h}; out=gpurun_out/$tag
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> $out/pytest_parity.log
tail -4 $out/pytest_parity.log
run() {  # name, precision, env...
  name=$1; prec=$2; shift 2
  env "$@" timeout 300 python bench.py --precision $prec --steps 20 --warmup 5 --no-cpu --no-e2e --no-others --no-secondary > $out/bench_cfg2_$name.json 2> $out/bench_cfg2_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_cfg2_$name.json"))
    p=d["parity"]
    print("$name", "value=%.3e"%d["value"], "ms=%.4f"%d["ms_per_step"], "frac=%.3f"%d["roofline"]["frac"], d["launch"]["kernel"], d["jit"][:40], "relL2=%.2e"%p["rel_l2_points"], "ks med/p999/max=%.1e/%.1e/%.1e"%(p["median_rel_err_kscores"],p["p999_rel_err_kscores"],p["max_rel_err_kscores"]))
except Exception as e:
    print("$name failed", e); print(open("$out/bench_cfg2_$name.err").read()[-800:])
PY
}
run f32 f32 A=1
run f32_ni1 f32 SNOWTRI_JIT_DEFINES="P1_NI=1"
run mixed mixed A=1
run mixed_ni2 mixed SNOWTRI_JIT_DEFINES="P1_NI=2"
run f64 f64 A=1
env timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/bench_cfg3.json 2> $out/bench_cfg3.err
python -c "
import json; d=json.load(open('$out/bench_cfg3.json')); print('cfg3 value=%.3e ms=%.4f'%(d['value'], d['ms_per_step']), d['parity']['nout_equal'], d['parity']['rel_l2_points'])"
for prec in f32 mixed; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:p1_jit -s 4 -c 1 -f -o $out/p1_jit_cfg2_$prec \
    python bench.py --precision $prec --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_p1_$prec.log 2>&1
  tail -1 $out/ncu_p1_$prec.log | cut -c1-160
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $out/launches_cfg2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/launches_cfg2.log 2>&1
grep -v "^==" $out/launches_cfg2.csv | awk -F'","' 'NR>1{print $7, $NF}' | head -4
