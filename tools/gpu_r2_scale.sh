set -x
mkdir -p gpurun_out
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}" 2>&1 | tail -1; }
run 29511 --steps 10 --warmup 3 > gpurun_out/bench_c2_n${N}_v3.json
run 29512 --workload c3 --steps 8 --warmup 3 > gpurun_out/bench_c3_n${N}_v3.json
for f in c2 c3; do cut -c1-330 gpurun_out/bench_${f}_n${N}_v3.json; done
