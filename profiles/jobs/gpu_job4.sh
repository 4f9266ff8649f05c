#!/bin/bash
# round 2, GPU job 4 (1 GPU): heat kernels after the per-node operand rewrite: tests, A/B of the assembly variants, C3 bench, slab tests on a shared device
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scenarios.py tests/test_golden.py tests/test_gpu_slabs.py -m gpu -q -s --durations=8 > gpurun_out/r2_gpu_tests_4.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_4.txt; tail -12 gpurun_out/r2_gpu_tests_4.txt
for v in g1b4 g2b4 g1b5 g2b5 g2b6h6 g2b4u2; do
  timeout 200 python build/ab/run.py build/ab/libsf3d_$v.so --heat --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ab_heat_$v.json 2> gpurun_out/r2_ab_heat_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_ab_heat_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step %.2f"%d["ms_per_step"], {k:x for k,x in d["kernel_ms"].items() if x})
except Exception as e: print("$v", "failed", e)
PY
done
timeout 300 python bench.py --heat --steps 4 --warmup 3 > gpurun_out/r2_bench_heat_b.json 2> gpurun_out/r2_bench_heat_b.err
