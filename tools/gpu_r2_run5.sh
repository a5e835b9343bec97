#!/bin/bash
# 2-GPU checks: the training step (broadcast at construction, graph + all-reduce), configs[2] and configs[3] at N = 2
mkdir -p gpurun_out
for wl in c2 c3 c4; do
  steps=10; [ $wl = c3 ] && steps=5; [ $wl = c4 ] && steps=20
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${wl}_n2.json 2> gpurun_out/r2_bench_${wl}_n2.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_${wl}_n2.json").read().strip().splitlines()[-1])
    print("${wl}", "N=2", d["value"], d["unit"], d["ms_per_step"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("${wl} failed", e); print(open("gpurun_out/r2_bench_${wl}_n2.err").read()[-1500:])
PY
done
