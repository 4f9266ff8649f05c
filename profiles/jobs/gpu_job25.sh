#!/bin/bash
# round 2, GPU job 25 (1 GPU): 2 ranks sharing the device, coupled heat in save mode All (runs the water-flux snapshot pass under several ranks)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_gpu_slabs.py::test_slabs_match_oracle[2-heat-all]" "tests/test_gpu_slabs.py::test_slabs_match_oracle[2-True]" -m gpu -q -s > gpurun_out/r2_slab_tests_25.txt 2>&1; echo "rc=$?"; grep -E "mgpu_slab_check|passed|failed|Error" gpurun_out/r2_slab_tests_25.txt | cut -c1-700
