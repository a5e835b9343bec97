#!/bin/bash
# profile artefacts for profiles/: ncu --set full of the mask-einsum GEMM, the launch list of one eager step, the torch.profiler step table
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 3 -c 1 -o gpurun_out/prof_einsum_fwd_v3 -f python tools/prof_gemm.py > gpurun_out/ncu_einsum.log 2>&1
timeout 120 ncu -i gpurun_out/prof_einsum_fwd_v3.ncu-rep --page raw --csv > gpurun_out/ncu_einsum_raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r01_v3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph > gpurun_out/bench_under_ncu.log 2>&1
timeout 400 python tools/step_profile.py gpurun_out/step_profile_v10.txt > gpurun_out/step_profile.log 2>&1
wc -l gpurun_out/launches_r01_v3.csv gpurun_out/ncu_einsum_raw.csv; head -12 gpurun_out/step_profile_v10.txt | cut -c1-140
