set -x
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py tests/test_head_gpu.py tests/test_zz_graph_bucket_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/step_profile.py gpurun_out/step_profile_c3_v9.txt c3 > /dev/null 2>&1
head -1 gpurun_out/step_profile_c3_v9.txt | cut -c150-300; grep "matcher_cost" gpurun_out/step_profile_c3_v9.txt | cut -c1-100
python tools/step_profile.py gpurun_out/step_profile_c2_v11.txt c2 > /dev/null 2>&1
head -1 gpurun_out/step_profile_c2_v11.txt | cut -c100-250; grep "matcher_cost" gpurun_out/step_profile_c2_v11.txt | cut -c1-100
