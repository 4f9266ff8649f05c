#!/bin/bash
# round 2, GPU job 3 (1 GPU): whole GPU suite, shared-device multi-rank debug, bench (both arms), config 1, ncu of one C3 step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SF3D_SHARE_DEVICE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29801 \
    tests/mgpu_slab_check.py > gpurun_out/r2_share_debug.txt 2>&1
echo "share rc=$?"; grep -n "sf3d\|rror" gpurun_out/r2_share_debug.txt | head -20
timeout 1500 python -m pytest tests -m gpu -q -s --durations=12 > gpurun_out/r2_gpu_tests_3.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_3.txt; tail -12 gpurun_out/r2_gpu_tests_3.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref_b.json 2> gpurun_out/r2_bench_ref_b.err; echo "ref rc=$?"
timeout 300 python tests/run_config1.py product 24 > gpurun_out/r2_c1_b.log 2>&1
NCU="ncu --set full --clock-control none --profile-from-start off"
timeout 600 $NCU -k regex:'kern_(heat_assemble|heat_coeffs|save_water_fluxes|boundary_heat|heat_begin|heat_post|heat_accept|update_conductance|assemble|node_phase|post|accept|begin_try|heat_copy)' \
    -o gpurun_out/r2_c3_step python profiles/capture_step.py --heat > gpurun_out/r2_ncu_c3_step.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_c3_step.ncu-rep gpurun_out/r2_c3_step_summary.json > gpurun_out/r2_c3_step_summary.txt 2>&1
ls -la gpurun_out/r2_c3_step.ncu-rep; SZ=$(stat -c %s gpurun_out/r2_c3_step.ncu-rep); if [ "$SZ" -gt 45000000 ]; then rm gpurun_out/r2_c3_step.ncu-rep; fi
cat gpurun_out/r2_c3_step_summary.txt
