set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_gpu_v6.log
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_v4.json
python bench.py --workload c3 --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_c3_n1_v4.json
python tools/step_profile.py gpurun_out/step_profile_c2_v9.txt c2 > /dev/null 2>&1
cat gpurun_out/pytest_gpu_v6.log
for f in bench_v4 bench_c3_n1_v4; do cut -c1-330 gpurun_out/$f.json; done
head -1 gpurun_out/step_profile_c2_v9.txt
