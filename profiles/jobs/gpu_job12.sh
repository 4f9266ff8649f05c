#!/bin/bash
# round 2, GPU job 12 (1 GPU): overlapped output-map download, value region without per-launch events: tests + bench (40 and default 20 steps)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_async_rasters.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2_gpu_tests_12.txt 2>&1; tail -3 gpurun_out/r2_gpu_tests_12.txt
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1_d.err
timeout 600 python bench.py > gpurun_out/r2_bench_n1_default.json 2> gpurun_out/r2_bench_n1_default.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_n1_d","r2_bench_n1_default"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e ms %.3f value %.4g"%(d["e2e"]["ms_per_step"], d["e2e"]["value"]), d["e2e"]["host_ms_per_step"], d["e2e"]["same_steps_as_value"], d["kernel_times_from"][:60], {k:x for k,x in d["kernel_ms"].items() if x}, "frac", d["roofline"]["frac"], d["roofline"]["frac_traffic"], "c4", (d.get("c4") or {}).get("ms_per_step"), d["clocks"])
PY
