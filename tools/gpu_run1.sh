#!/bin/bash
# round-1 measurement pass: parity tests, kernel microbench, step profile, ncu captures, bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python tools/microbench.py > gpurun_out/microbench.log 2>&1
python tools/step_profile.py gpurun_out/step_profile.txt > gpurun_out/step_profile.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'msda_fwd_d32|mask_einsum_tc' -c 6 -o gpurun_out/prof_r01_kernels -f python tools/microbench.py > gpurun_out/ncu_full.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/microbench.log | head -60; tail -2 gpurun_out/bench.log
