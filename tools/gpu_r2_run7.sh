#!/bin/bash
# Final round-2 measurements on one GPU: full GPU suite, the driver's bench command (both arms), C3, C4, launch list + step share.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_3.log; tail -n 3 gpurun_out/r2_pytest_gpu_3.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -n 2 gpurun_out/r2_bench_final.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
timeout 500 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_final.json 2>/dev/null
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r2_bench_c4_final.json 2>/dev/null
timeout 500 python tools/step_profile.py gpurun_out/r2_step_profile_c2_v3.txt > /dev/null 2>&1
timeout 600 python tools/step_profile.py gpurun_out/r2_step_profile_c3_v3.txt c3 > /dev/null 2>&1
PDB_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/r2_step_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph > /dev/null 2>&1
python - <<'PY'
import json
for f in ("r2_bench_final", "r2_bench_reference_arm", "r2_bench_c3_final", "r2_bench_c4_final"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], d.get("ms_per_step"), "e2e", d["e2e"]["value"], d.get("cpu_baseline", {}).get("kind"))
    except Exception as e:
        print(f, "failed", e)
PY
