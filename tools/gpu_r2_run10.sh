set -x
mkdir -p gpurun_out
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --distinct-batches 4 2>&1 | tail -1 | cut -c1-1100 > gpurun_out/bench_c2_vary_nobucket.json
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --distinct-batches 4 --target-bucket 8 2>&1 | tail -1 | cut -c1-1100 > gpurun_out/bench_c2_vary_bucket8.json
python bench.py --workload c3 --steps 8 --warmup 3 --distinct-batches 4 2>&1 | tail -1 | cut -c1-1100 > gpurun_out/bench_c3_vary_nobucket.json
python bench.py --workload c3 --steps 8 --warmup 3 --distinct-batches 4 --target-bucket 8 2>&1 | tail -1 | cut -c1-1100 > gpurun_out/bench_c3_vary_bucket8.json
python bench.py --workload c3 --steps 8 --warmup 3 2>&1 | tail -1 | cut -c1-1100 > gpurun_out/bench_c3_fixed.json
head -c 1100 gpurun_out/bench_c*_vary*.json gpurun_out/bench_c3_fixed.json
