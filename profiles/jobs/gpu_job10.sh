#!/bin/bash
# round 2, GPU job 10 (1 GPU): whole GPU suite with the prepared-try accept pass, bench (40 steps) + A/B without it, ncu launch list and
# ncu --set full of one water step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --durations=12 > gpurun_out/r2_gpu_tests_10.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_10.txt; tail -14 gpurun_out/r2_gpu_tests_10.txt
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; echo "bench rc=$?"
SF3D_NO_PREPARED_TRY=1 timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_c_noprep.json 2> gpurun_out/r2_bench_n1_c_noprep.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_n1_c","r2_bench_n1_c_noprep"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step %.3f"%d["ms_per_step"], "e2e %.3f"%d["e2e"]["ms_per_step"], "value %.4g"%d["value"], {k:x for k,x in d["kernel_ms"].items() if x}, d["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1; echo "ncu list rc=$?"
NCU="ncu --set full --clock-control none --profile-from-start off --import-source on"
timeout 600 $NCU -k regex:'kern_(jacobi|assemble|node_phase|post|accept|begin_try)' \
    -o gpurun_out/r2_c2_step python profiles/capture_step.py > gpurun_out/r2_ncu_c2_step.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/r2_c2_step.ncu-rep gpurun_out/r2_c2_step_summary.json > gpurun_out/r2_c2_step_summary.txt 2>&1
SZ=$(stat -c %s gpurun_out/r2_c2_step.ncu-rep); echo "rep size $SZ"; if [ "$SZ" -gt 45000000 ]; then rm gpurun_out/r2_c2_step.ncu-rep; fi
cat gpurun_out/r2_c2_step_summary.txt | head -40
