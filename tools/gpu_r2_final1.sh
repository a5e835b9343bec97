set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu_v5.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference_arm_v2.json
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_v3.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --packed-masks 2>&1 | tail -1 > gpurun_out/bench_v3_packed.json
python bench.py --workload c3 --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_c3_n1_v3.json
python bench.py --workload c4 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_c4_n1_v4.json
python tools/step_profile.py gpurun_out/step_profile_c2_v8.txt c2 > /dev/null 2>&1
cat gpurun_out/pytest_gpu_v5.log
for f in bench_reference_arm_v2 bench_v3 bench_v3_packed bench_c3_n1_v3 bench_c4_n1_v4; do cut -c1-420 gpurun_out/$f.json; done
head -1 gpurun_out/step_profile_c2_v8.txt
