#!/bin/bash
# round 2, GPU job 17 (1 GPU): full C2 6 h run, compute-sanitizer memcheck / racecheck on the round-2 kernels, ncu of the persistent solve and accept<true>
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tests/run_full_config.py 2 > gpurun_out/r2_full_c2.log 2>&1; echo "full c2 rc=$?"; tail -2 gpurun_out/r2_full_c2.log | cut -c1-600
cp gpurun_out/full_config2.json gpurun_out/r2_full_config2_c2_6h.json 2>/dev/null
SEL="tests/test_gpu_persistent_solve.py::test_launch_counts tests/test_gpu_async_rasters.py tests/test_gpu_raster_prep.py::test_edge_rasters tests/test_gpu_raster_prep.py::test_bundled_dem_matches_the_reference_golden tests/test_gpu_parity.py"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $SEL -m gpu -q > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_persistent_solve.py::test_launch_counts tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2_sanitizer_racecheck.log
NCU="ncu --set full --clock-control none --profile-from-start off"
timeout 300 $NCU -k regex:'kern_(jacobi_persistent|accept|assemble|node_phase|post)' -o gpurun_out/r2_c1_step python profiles/capture_step.py --rows 139 --cols 150 --soil-layers 5 > gpurun_out/r2_ncu_c1_step.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/r2_c1_step.ncu-rep gpurun_out/r2_c1_step_summary.json > gpurun_out/r2_c1_step_summary.txt 2>&1; rm -f gpurun_out/r2_c1_step.ncu-rep; cat gpurun_out/r2_c1_step_summary.txt | head
