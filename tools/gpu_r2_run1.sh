#!/bin/bash
# Round 2, first GPU call: the whole GPU suite WITHOUT -x (log kept), then the same-box reference-PyTorch operator bars.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_1.log
tail -40 gpurun_out/r2_pytest_gpu_1.log
timeout 400 python tools/bench_reference_ops.py > gpurun_out/reference_ops.log 2>&1; tail -50 gpurun_out/reference_ops.log
timeout 300 python tools/bench_postprocess.py > gpurun_out/pp_bench.json 2> gpurun_out/pp_bench.err; cat gpurun_out/pp_bench.json | tail -30
