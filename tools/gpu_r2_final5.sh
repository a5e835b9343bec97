set -x
mkdir -p gpurun_out
python bench.py 2>&1 | tail -1 > gpurun_out/bench_v6.json; cut -c1-260 gpurun_out/bench_v6.json
python bench.py --workload c3 --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_c3_n1_v5.json; cut -c1-260 gpurun_out/bench_c3_n1_v5.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --packed-masks 2>&1 | tail -1 > gpurun_out/bench_v6_packed.json
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --distinct-batches 4 --target-bucket 8 2>&1 | tail -1 > gpurun_out/bench_c2_vary_bucket8_v2.json; cut -c1-200 gpurun_out/bench_c2_vary_bucket8_v2.json
