#!/bin/bash
# round 2, GPU job 14 (1 GPU): mapped forcing maps: tests + bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_async_rasters.py tests/test_gpu_scenarios.py -m gpu -q -k "async or page_locked or raster_forcing" > gpurun_out/r2_gpu_tests_14.txt 2>&1; tail -3 gpurun_out/r2_gpu_tests_14.txt
timeout 600 python bench.py --steps 40 --warmup 3 --no-c4 --no-cpu-baseline > gpurun_out/r2_bench_n1_f.json 2> gpurun_out/r2_bench_n1_f.err; echo "bench rc=$?"
SF3D_NO_MAPPED_FORCING=1 timeout 600 python bench.py --steps 40 --warmup 3 --no-c4 --no-cpu-baseline > gpurun_out/r2_bench_n1_f_staged.json 2> gpurun_out/r2_bench_n1_f_staged.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_n1_f","r2_bench_n1_f_staged"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e ms %.3f value %.4g"%(d["e2e"]["ms_per_step"], d["e2e"]["value"]), d["e2e"]["host_ms_per_step"], d["clocks"])
PY
