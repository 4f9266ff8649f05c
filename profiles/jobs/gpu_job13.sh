#!/bin/bash
# round 2, GPU job 13 (1 GPU): heat tests after dropping the unobserved flux snapshot, C3 bench, ncu of one C3 step, 200-step C2 bench, 40-step bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scenarios.py tests/test_gpu_async_rasters.py -m gpu -q > gpurun_out/r2_gpu_tests_13.txt 2>&1; tail -3 gpurun_out/r2_gpu_tests_13.txt
timeout 300 python bench.py --heat --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_heat_c.json 2> gpurun_out/r2_bench_heat_c.err; echo "heat bench rc=$?"
timeout 600 python bench.py --steps 40 --warmup 3 --no-c4 > gpurun_out/r2_bench_n1_e.json 2> gpurun_out/r2_bench_n1_e.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 200 --warmup 3 --no-c4 > gpurun_out/r2_bench_n1_200.json 2> gpurun_out/r2_bench_n1_200.err; echo "bench200 rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_heat_c","r2_bench_n1_e","r2_bench_n1_200"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e ms %.3f value %.4g"%(d["e2e"]["ms_per_step"], d["e2e"]["value"]), d["e2e"]["host_ms_per_step"], {k:x for k,x in d["kernel_ms"].items() if x}, (d.get("cpu_baseline") or {}).get("value"), d["clocks"])
PY
NCU="ncu --set full --clock-control none --profile-from-start off"
timeout 600 $NCU -k regex:'kern_(heat_assemble|heat_coeffs|save_water_fluxes|boundary_heat|heat_begin|heat_post|heat_accept|update_conductance|assemble|node_phase|post|accept|begin_try|heat_copy|thermal_invariant)' \
    -o gpurun_out/r2_c3_step_b python profiles/capture_step.py --heat > gpurun_out/r2_ncu_c3_step_b.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_c3_step_b.ncu-rep gpurun_out/r2_c3_step_b_summary.json > gpurun_out/r2_c3_step_b_summary.txt 2>&1
rm -f gpurun_out/r2_c3_step_b.ncu-rep
cat gpurun_out/r2_c3_step_b_summary.txt | sort | uniq -c | sort -rn | head -5 > /dev/null
awk '{print $1,$2}' gpurun_out/r2_c3_step_b_summary.txt | head -0
python - <<'PY'
import json, collections
recs=json.load(open("gpurun_out/r2_c3_step_b_summary.json"))
agg=collections.defaultdict(lambda:[0,0.0,0.0])
for r in recs:
    a=agg[r["kernel"]]; a[0]+=1; a[1]+=r.get("duration_us",0); a[2]+=r.get("dram_bytes",0)
for k,(n,us,b) in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{k:32s} n={n:3d} avg {us/n:8.1f} us  dram {b/n/1e9:6.3f} GB")
PY
