set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260
python bench.py --workload c3 --steps 8 --warmup 3 2>&1 | tail -1 | cut -c1-260
