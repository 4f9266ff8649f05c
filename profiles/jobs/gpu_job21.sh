#!/bin/bash
# round 2, GPU job 21 (1 GPU): A/B of the two-pass thermal-invariant kernel (resident blocks x link unroll)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in t8u2 t8u5 t8u10 t8u1 t6u5; do
  timeout 200 python build/ab/run.py build/ab/libsf3d_$v.so --heat --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ab_thermal_$v.json 2> gpurun_out/r2_ab_thermal_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_ab_thermal_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step %.2f"%d["ms_per_step"], "assemble+thermal per approximation %.3f ms"%(d["kernel_ms"]["assemble"]/d["approximations"]), d["clocks"]["sm_mhz"])
except Exception as e: print("$v", "failed", e)
PY
done
