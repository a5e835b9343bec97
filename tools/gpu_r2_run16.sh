set -x
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py tests/test_head_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/step_profile.py gpurun_out/step_profile_c2_v10.txt c2 > /dev/null 2>&1
head -1 gpurun_out/step_profile_c2_v10.txt | cut -c100-250; grep "threshold\|gemm_tf32x3_kernel<128" gpurun_out/step_profile_c2_v10.txt | cut -c1-140
sed -i 's/^ffn_fused_rows = 2048 /ffn_fused_rows = 1 << 30 /' partdistillation_b200/functional.py
python tools/step_profile.py gpurun_out/step_profile_c2_v10_nofuse.txt c2 > /dev/null 2>&1
head -1 gpurun_out/step_profile_c2_v10_nofuse.txt | cut -c100-250; grep "threshold\|gemm_tf32x3_kernel<128" gpurun_out/step_profile_c2_v10_nofuse.txt | cut -c1-140
