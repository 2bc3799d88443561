#!/bin/bash
# per-rank timing of the headline workload at N ranks (diagnostic)
N=${1:-8}; out=gpurun_out/${2:-n8diag}; mkdir -p $out
for steps in 10 50; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps $steps --warmup 5 --no-cpu --no-e2e --no-others --no-secondary > $out/bench_${N}gpu_s$steps.json 2> $out/bench_${N}gpu_s$steps.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/bench_${N}gpu_s$steps.json") if l.startswith("{")][-1])
    print("steps=$steps ms=%.4f"%d["ms_per_step"], d["timed_region_per_rank_ms"], d["clocks"])
except Exception as e:
    print("failed", e); print(open("$out/bench_${N}gpu_s$steps.err").read()[-1500:])
PY
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv > $out/smi.txt; cat $out/smi.txt
