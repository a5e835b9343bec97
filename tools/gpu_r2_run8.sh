set -x
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "gemm_small or small_rows" 2>&1 | tail -5
python tools/bench_small_gemm.py 2>&1 | tee gpurun_out/small_gemm.txt
