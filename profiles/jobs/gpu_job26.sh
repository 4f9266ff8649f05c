#!/bin/bash
# round 2, GPU job 26 (1 GPU): post pass enqueued behind the sweeps (one control read per approximation): parity tests, config 1, short bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_persistent_solve.py tests/test_gpu_scenarios.py tests/test_gpu_large_window.py "tests/test_gpu_slabs.py::test_slabs_match_oracle[2-False]" -m gpu -q > gpurun_out/r2_gpu_tests_26.txt 2>&1; tail -3 gpurun_out/r2_gpu_tests_26.txt
timeout 300 python tests/run_config1.py product 24 > gpurun_out/r2_c1_d.log 2>&1; tail -1 gpurun_out/r2_c1_d.log | cut -c1-420
SF3D_POST_FOLLOWS_SOLVE=0 timeout 300 python tests/run_config1.py product 24 > gpurun_out/r2_c1_d_off.log 2>&1; tail -1 gpurun_out/r2_c1_d_off.log | cut -c1-200
timeout 600 python bench.py --steps 40 --warmup 3 --no-c4 --no-cpu-baseline > gpurun_out/r2_bench_n1_g.json 2> gpurun_out/r2_bench_n1_g.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1_g.json").read().strip().splitlines()[-1])
print("ms/step %.3f value %.4g e2e ms %.3f"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d["e2e"]["same_steps_as_value"], d["gpu_launches"], d["clocks"])
PY
