#!/bin/bash
# round 2, GPU job 27 (1 GPU): final whole GPU suite of the round
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_gpu_tests_27.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_27.txt; tail -10 gpurun_out/r2_gpu_tests_27.txt | cut -c1-160
