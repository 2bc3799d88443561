#!/bin/bash
# full GPU suite, then the p1 modes with the shared-memory float64 constants of the mixed mode
tag=${1:-r2m}; out=gpurun_out/$tag
mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
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
run f32 f32 A=1
run mixed_smem mixed A=1
run mixed_smem_ni2 mixed SNOWTRI_JIT_DEFINES="P1_NI=2"
run mixed_cbank mixed SNOWTRI_JIT_DEFINES="P1_E64_SMEM=0"
run mixed_smem_b1 mixed SNOWTRI_JIT_MINB=1
run f64 f64 A=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:p1_jit -s 4 -c 1 -f -o $out/p1_jit_cfg2_mixed \
    python bench.py --precision mixed --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_p1_mixed.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:p1_jit -s 4 -c 1 -f -o $out/p1_jit_cfg2_f32 \
    python bench.py --precision f32 --steps 2 --warmup 3 --no-cpu --no-e2e --no-others --no-secondary > $out/ncu_p1_f32.log 2>&1
