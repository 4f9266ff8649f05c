#!/bin/bash
# round 2, GPU job 23 (1 GPU): heat rows with pattern-derived link nodes and the write-only first flux save: parity tests, then A/B of the link-loop unrolls
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scenarios.py tests/test_golden.py -m gpu -q -k "heat" > gpurun_out/r2_gpu_tests_23.txt 2>&1; tail -3 gpurun_out/r2_gpu_tests_23.txt
for v in a1h1 a2h1 a5h1 a1h2 a2h5; do
  timeout 200 python build/ab/run.py build/ab/libsf3d_$v.so --heat --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ab_heat2_$v.json 2> gpurun_out/r2_ab_heat2_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_ab_heat2_$v.json").read().strip().splitlines()[-1])
    k=d["kernel_ms"]; n=d["heat_steps"]
    print("$v", "ms/step %.2f"%d["ms_per_step"], "heat_assemble %.3f heat_accept %.3f heat_coeffs(+begin) %.3f ms per heat sub-step"%(k["heat_assemble"]/n, k["heat_accept"]/n, k["heat_coeffs"]/n), "asm+thermal %.3f"%(k["assemble"]/d["approximations"]), d["clocks"]["sm_mhz"])
except Exception as e: print("$v", "failed", e)
PY
done
