#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "group or grouping" > gpurun_out/r2_pytest_group.log 2>&1; tail -5 gpurun_out/r2_pytest_group.log
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4.err; cat gpurun_out/r2_bench_c4_n1.json; tail -3 gpurun_out/r2_bench_c4.err
timeout 300 ncu --set full --clock-control none -k regex:"gemm_tf32x3" -c 3 -o gpurun_out/r2_traffic_einsum -f python tools/prof_msda.py einsum > gpurun_out/ncu_t3.log 2>&1
timeout 120 ncu -i gpurun_out/r2_traffic_einsum.ncu-rep --page raw --csv > gpurun_out/r2_traffic_einsum_raw.csv 2>/dev/null
tail -n 2 gpurun_out/ncu_t3.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 2500 gpurun_out/r2_bench_b.json; tail -n 5 gpurun_out/r2_bench_b.err
