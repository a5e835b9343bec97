set -x
mkdir -p gpurun_out
python bench.py 2>&1 | tail -1 > gpurun_out/bench_v7.json; cut -c1-260 gpurun_out/bench_v7.json
python bench.py --workload c3 --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_c3_n1_v6.json; cut -c1-260 gpurun_out/bench_c3_n1_v6.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
