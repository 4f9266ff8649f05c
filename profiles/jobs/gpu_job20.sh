#!/bin/bash
# round 2, GPU job 20 (1 GPU): ncu --set full with source of the thermal-invariant pass and the heat assembly (one launch each)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --profile-from-start off --import-source on"
timeout 400 $NCU -k regex:'kern_(thermal_invariant|heat_assemble)' -c 2 -o gpurun_out/r2_heat_two python profiles/capture_step.py --heat > gpurun_out/r2_ncu_heat_two.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_heat_two.ncu-rep
