#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_2.log; tail -n 12 gpurun_out/r2_pytest_gpu_2.log
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4.err; cat gpurun_out/r2_bench_c4_n1.json; tail -n 3 gpurun_out/r2_bench_c4.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_c.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])"; tail -n 5 gpurun_out/r2_bench_c.err
