#!/bin/bash
# p1 with cp.async staging: items per lane x resident CTAs per SM, both float modes; tests first
tag=${1:-r2j}; out=gpurun_out/$tag
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
    print("$name", "value=%.3e"%d["value"], "ms=%.4f"%d["ms_per_step"], "frac=%.3f"%d["roofline"]["frac"], d["launch"], d["jit"][:40], "relL2=%.2e"%p["rel_l2_points"], "ks max=%.1e"%(p["max_rel_err_kscores"]))
except Exception as e:
    print("$name failed", e); print(open("$out/bench_cfg2_$name.err").read()[-800:])
PY
}
run f32_ni2_b2 f32 A=1
run f32_ni1_b2 f32 SNOWTRI_JIT_DEFINES="P1_NI=1"
run f32_ni1_b3 f32 SNOWTRI_JIT_DEFINES="P1_NI=1" SNOWTRI_JIT_MINB=3
run f32_ni2_b3 f32 SNOWTRI_JIT_MINB=3
run f32_ni1_b4 f32 SNOWTRI_JIT_DEFINES="P1_NI=1" SNOWTRI_JIT_MINB=4
run mixed_ni1_b2 mixed A=1
run mixed_ni1_b3 mixed SNOWTRI_JIT_MINB=3
run f64 f64 A=1
for v in "f32_ni1_b3 f32 3" ; do
  set -- $v
  SNOWTRI_JIT_DEFINES="P1_NI=1" SNOWTRI_JIT_MINB=$3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:p1_jit -s 4 -c 1 -f -o $out/p1_jit_cfg2_$1 \
    python bench.py --precision $2 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_p1_$1.log 2>&1
  tail -1 $out/ncu_p1_$1.log | cut -c1-160
done
