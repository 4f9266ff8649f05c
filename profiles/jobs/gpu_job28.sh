#!/bin/bash
# round 2, GPU job 28 (1 GPU): the driver's default bench invocation on the final code of the round
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/r2_bench_final2_n1.json 2> gpurun_out/r2_bench_final2_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_final2_n1.json").read().strip().splitlines()[-1])
print("ms/step %.3f value %.4g e2e %.4g c4 %.2f ms"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["c4"]["ms_per_step"]), "frac", round(d["roofline"]["frac"],3), round(d["roofline"]["frac_traffic"],3), "asm", round(d["roofline_assembly"]["frac"],3), "cpu", d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"])
PY
