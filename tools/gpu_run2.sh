#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python tools/microbench.py > gpurun_out/microbench.log 2>&1
python tools/step_profile.py gpurun_out/step_profile.txt > gpurun_out/step_profile.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; grep -A5 mask_einsum gpurun_out/microbench.log; head -30 gpurun_out/step_profile.txt | cut -c1-140; tail -1 gpurun_out/bench.log | cut -c1-400
