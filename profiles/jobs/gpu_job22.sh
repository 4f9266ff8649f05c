#!/bin/bash
# round 2, GPU job 22 (1 GPU): ncu --set full with source of kern_assemble<0> / kern_node_phase<0> (C2) and kern_heat_assemble / kern_heat_accept (C3), one launch each
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --profile-from-start off --import-source on"
timeout 300 $NCU -k regex:'kern_(assemble|node_phase)' -c 2 -o gpurun_out/r2_src_water python profiles/capture_step.py > gpurun_out/r2_ncu_src_water.log 2>&1; echo "ncu rc=$?"
timeout 300 $NCU -k regex:'kern_heat_(assemble|accept)' -c 2 -o gpurun_out/r2_src_heat python profiles/capture_step.py --heat > gpurun_out/r2_ncu_src_heat.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_src_*.ncu-rep
