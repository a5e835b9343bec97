set -x
mkdir -p gpurun_out
PDB_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/r2_step_launches_v2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph > /dev/null 2>&1
wc -l gpurun_out/r2_step_launches_v2.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_small_kernel" -c 4 -o gpurun_out/r2_ncu_gemm_small -f python tools/bench_small_gemm.py > gpurun_out/ncu_small.log 2>&1
timeout 120 ncu -i gpurun_out/r2_ncu_gemm_small.ncu-rep --page raw --csv > gpurun_out/r2_ncu_gemm_small_raw.csv 2>/dev/null
tail -2 gpurun_out/ncu_small.log
