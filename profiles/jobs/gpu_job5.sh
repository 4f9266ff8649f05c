#!/bin/bash
# round 2, GPU job 5 (1 GPU): thermal-invariant split A/B, C3 bench, C5 slab at N = 1, ncu of the final C3 step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scenarios.py tests/test_golden.py -m gpu -q -k "heat" > gpurun_out/r2_gpu_tests_5.txt 2>&1; tail -3 gpurun_out/r2_gpu_tests_5.txt
for v in t8u2 t6u2 t8u1 t6u5 t5u10; do
  timeout 200 python build/ab/run.py build/ab/libsf3d_$v.so --heat --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ab_heat_$v.json 2> gpurun_out/r2_ab_heat_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_ab_heat_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step %.2f"%d["ms_per_step"], {k:x for k,x in d["kernel_ms"].items() if x})
except Exception as e: print("$v", "failed", e)
PY
done
timeout 600 python tests/run_config5.py --steps-per-phase 5 > gpurun_out/r2_c5_n1.log 2>&1; tail -c 1500 gpurun_out/r2_c5_n1.log
