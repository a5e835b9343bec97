#!/bin/bash
# First GPU call for the f4 eval-branch kernels (written after round 1's GPU budget was spent):
# parity tests, the per-image microbenchmark against the dense PyTorch expression, launch list + one ncu --set full capture.
#   gpurun --timeout 900 -- 'bash tools/gpu_run6.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_postprocess_gpu.py tests/test_zz_engine_lr_gpu.py -q > gpurun_out/pp_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pp_tests.log
timeout 300 python tools/bench_postprocess.py > gpurun_out/pp_bench.json 2> gpurun_out/pp_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/pp_launches.csv python tools/bench_postprocess.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:postprocess_masks_kernel -s 3 -c 1 -o gpurun_out/prof_postprocess_v1 -f python tools/bench_postprocess.py > gpurun_out/ncu_pp.log 2>&1
timeout 120 ncu -i gpurun_out/prof_postprocess_v1.ncu-rep --page raw --csv > gpurun_out/ncu_pp_raw.csv 2>/dev/null
tail -3 gpurun_out/pp_tests.log; cat gpurun_out/pp_bench.json
timeout 600 python tools/bench_reference_ops.py > gpurun_out/reference_ops.log 2>&1; tail -40 gpurun_out/reference_ops.log
# f3: the e2e leg fed with bit-packed masks next to the default (bool) one
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bool_masks.json 2> gpurun_out/bench_bool_masks.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --packed-masks > gpurun_out/bench_packed_masks.json 2> gpurun_out/bench_packed_masks.err
python - <<'EOF'
import json
for f in ("bench_bool_masks", "bench_packed_masks"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"])
    except Exception as e:
        print(f, "failed", e)
EOF
