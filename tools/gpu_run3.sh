#!/bin/bash
# profile artefacts for profiles/: ncu --set full of the named kernels, the launch list of one eager step, bench lines
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 3 -c 1 -o gpurun_out/prof_einsum_fwd -f python tools/prof_gemm.py > gpurun_out/ncu_einsum.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 2 -c 2 -o gpurun_out/prof_msda -f python tools/prof_msda.py > gpurun_out/ncu_msda.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph > gpurun_out/bench_under_ncu.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-300; tail -1 gpurun_out/bench_reference.log | cut -c1-300; wc -l gpurun_out/launches_r01.csv
