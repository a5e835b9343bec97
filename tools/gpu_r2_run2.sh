#!/bin/bash
# Round 2: bench line (both arms), C4 line, ncu captures for the per-kernel DRAM traffic of this round's kernels.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4.err; cat gpurun_out/r2_bench_c4_n1.json; tail -3 gpurun_out/r2_bench_c4.err
timeout 300 ncu --set full --clock-control none -k regex:"msda_fwd|msda_bwd|gemm_tf32x3" -c 12 -o gpurun_out/r2_traffic_c2 -f python tools/prof_msda.py all > gpurun_out/ncu_t1.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"msda_fwd|msda_bwd" -c 8 -o gpurun_out/r2_traffic_c5 -f python tools/prof_msda.py all --c5 > gpurun_out/ncu_t2.log 2>&1
for f in r2_traffic_c2 r2_traffic_c5; do timeout 120 ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
tail -2 gpurun_out/ncu_t1.log gpurun_out/ncu_t2.log
