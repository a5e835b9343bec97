set -x
mkdir -p gpurun_out
python -m pytest tests/test_zz_graph_bucket_gpu.py tests/test_head_gpu.py -x -q -m gpu 2>&1 | tail -25
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-300
