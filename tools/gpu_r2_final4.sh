set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_gpu_v7.log; cat gpurun_out/pytest_gpu_v7.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py 2>&1 | tail -1 > gpurun_out/bench_v5.json; cut -c1-300 gpurun_out/bench_v5.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200
