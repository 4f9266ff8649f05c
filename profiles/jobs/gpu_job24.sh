#!/bin/bash
# round 2, GPU job 24 (1 GPU): final validation of the round: whole GPU suite, smoke(), default bench (both arms)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --durations=6 > gpurun_out/r2_gpu_tests_24.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_24.txt; tail -12 gpurun_out/r2_gpu_tests_24.txt | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout 600 python bench.py > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_final_ref.json 2> gpurun_out/r2_bench_final_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_final_n1.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2_bench_final_ref.json").read().strip().splitlines()[-1])
print("product ms/step %.3f value %.4g e2e %.4g  c4 %.2f ms"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["c4"]["ms_per_step"]), "frac", round(d["roofline"]["frac"],3), round(d["roofline"]["frac_traffic"],3), "asm frac", round(d["roofline_assembly"]["frac"],3), d["clocks"])
print("reference value %.4g cores %s same_config %s"%(r["value"], r["cpu_baseline"]["cores"], r["config"]["same_config"]), "ratio value %.1f e2e %.1f"%(d["value"]/r["value"], d["e2e"]["value"]/r["value"]))
PY
