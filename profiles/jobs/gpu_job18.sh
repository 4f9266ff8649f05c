#!/bin/bash
# round 2, GPU job 18 (1 GPU): heat tests after the thermal-invariant change, C3 bench, smoke(), slab heat check
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scenarios.py tests/test_golden.py -m gpu -q -k "heat" > gpurun_out/r2_gpu_tests_18.txt 2>&1; tail -3 gpurun_out/r2_gpu_tests_18.txt
timeout 300 python -m pytest "tests/test_gpu_slabs.py::test_slabs_match_oracle[2-True]" -m gpu -q >> gpurun_out/r2_gpu_tests_18.txt 2>&1; tail -2 gpurun_out/r2_gpu_tests_18.txt
timeout 300 python bench.py --heat --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_heat_d.json 2> gpurun_out/r2_bench_heat_d.err; echo "heat bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_heat_d.json").read().strip().splitlines()[-1])
print("C3 ms/step %.3f"%d["ms_per_step"], {k:x for k,x in d["kernel_ms"].items() if x}, d["clocks"])
PY
