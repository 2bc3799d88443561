#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): byte-identity test, then the bench line at N (torchrun, one rank per GPU).
N=${1:-2}; tag=${2:-r2m$N}; out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > $out/pytest_multi_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_multi_gpu.log
tail -5 $out/pytest_multi_gpu.log
for n in ${NLIST:-$(seq 1 $N)}; do
  case $n in 1|2|4|8) ;; *) continue;; esac
  t0=$SECONDS
  if [ $n -eq 1 ]; then
    timeout 1200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $out/bench_${n}gpu.json 2> $out/bench_${n}gpu.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 10 --warmup 3 > $out/bench_${n}gpu.json 2> $out/bench_${n}gpu.err
  fi
  echo "N=$n bench wall $((SECONDS-t0)) s"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/bench_${n}gpu.json") if l.startswith("{")][-1])
    print("N=$n value=%.3e"%d["value"], "with_gather=%s"%d.get("value_with_allgather"), "e2e=%.3e"%d["e2e"]["value"], "ceil", d["e2e"]["frac_of_copy_ceiling"], d.get("allgather"))
    for k,v in (d.get("secondary") or {}).items():
        if "value" in v: print("   ", k, "value=%.3e"%v["value"], "ms=%.1f"%v["ms_per_step"], "compute_only_ms=%s"%v.get("compute_only_ms"), v.get("checksum"))
        else: print("   ", k, v)
except Exception as e:
    print("N=$n bench failed", e); print(open("$out/bench_${n}gpu.err").read()[-2500:])
PY
done
