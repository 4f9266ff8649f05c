#!/bin/bash
# round 2, GPU job 16 (1 GPU): persistent small-graph solve: whole GPU suite, config 1 (24 h) with and without it
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --durations=6 > gpurun_out/r2_gpu_tests_16.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_16.txt; tail -12 gpurun_out/r2_gpu_tests_16.txt
timeout 300 python tests/run_config1.py product 24 > gpurun_out/r2_c1_c.log 2>&1; tail -1 gpurun_out/r2_c1_c.log | cut -c1-700
SF3D_PERSISTENT_SOLVE=0 timeout 300 python tests/run_config1.py product 24 > gpurun_out/r2_c1_c_launch_per_sweep.log 2>&1; tail -1 gpurun_out/r2_c1_c_launch_per_sweep.log | cut -c1-300
