set -x
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py tests/test_head_gpu.py tests/test_zz_engine_state_gpu.py tests/test_zz_graph_bucket_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/step_profile.py gpurun_out/step_profile_c3_v8.txt c3 > /dev/null 2>&1
head -1 gpurun_out/step_profile_c3_v8.txt
python bench.py --workload c3 --steps 8 --warmup 3 2>&1 | tail -1 | cut -c1-300
