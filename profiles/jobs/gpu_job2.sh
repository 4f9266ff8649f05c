#!/bin/bash
# round 2, GPU job 2 (2 GPUs): slab parity tests (real NVLink pair, 4 ranks on a shared device, heat, time-out) and the 2-rank bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_job2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_slabs.py -q -s --durations=8 > gpurun_out/r2_slab_tests_2gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_slab_tests_2gpu.txt
tail -15 gpurun_out/r2_slab_tests_2gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err
echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n2_a.err
SF3D_FUSED_EXCHANGE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus 2 --steps 20 --warmup 3 --no-c4 --no-parity-check > gpurun_out/r2_bench_n2_separate.json 2> gpurun_out/r2_bench_n2_separate.err
head -c 600 gpurun_out/r2_bench_n2_a.json
