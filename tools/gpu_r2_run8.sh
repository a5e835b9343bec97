set -x
mkdir -p gpurun_out
python tools/t_colsum.py
python -m pytest tests/test_ops_gpu.py tests/test_head_gpu.py -x -q -m gpu 2>&1 | tail -5
PROFILE_STACK=1 python tools/step_profile.py gpurun_out/step_profile_c2_v5.txt c2 > /dev/null 2>&1
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400
