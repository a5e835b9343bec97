#!/bin/bash
# 8-GPU evidence for configs[2] and configs[3] (the driver's own scaling run covers configs[1])
mkdir -p gpurun_out
for wl in c3 c4; do
  steps=5; [ $wl = c4 ] && steps=20
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload $wl --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${wl}_n8.json 2> gpurun_out/r2_bench_${wl}_n8.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_${wl}_n8.json").read().strip().splitlines()[-1])
    print("${wl}", "N=8", d["value"], d["unit"], d["ms_per_step"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("${wl} failed", e); print(open("gpurun_out/r2_bench_${wl}_n8.err").read()[-1500:])
PY
done
