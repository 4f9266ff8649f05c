#!/bin/bash
# round 2, GPU job 1: GPU test suite, baseline bench (water, heat), config 1 full day, ncu captures of the heat kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_job1_gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q -s --durations=12 > gpurun_out/r2_gpu_tests_1.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_1.txt
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
timeout 300 python bench.py --heat --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_heat_a.json 2> gpurun_out/r2_bench_heat_a.err
timeout 600 python tests/run_config1.py product 24 > gpurun_out/r2_c1_a.log 2>&1
NCU="ncu --set full --clock-control none"
timeout 400 $NCU -k regex:'kern_(heat_assemble|heat_coeffs|save_water_fluxes|boundary_heat|heat_begin|heat_post|heat_accept|update_conductance)' -s 10 -c 10 \
    -o gpurun_out/r2_heat_rows_a python bench.py --heat --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_heat_rows_a.log 2>&1
timeout 400 $NCU -k regex:'kern_(assemble|node_phase|heat_jacobi)' -s 12 -c 6 \
    -o gpurun_out/r2_heat_water_a python bench.py --heat --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_heat_water_a.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r2_gpu_tests_1.txt
cat gpurun_out/r2_bench_a.json | head -c 1500
