#!/bin/bash
# round 2, multi-GPU job: N ranks (N = number of GPUs of the box): slab parity tests on real devices, bench (C2 slabs + parity_check + C4 block), C5 slabs
N=${1:-4}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_job_n${N}_gpus.txt
if [ "$N" -le 4 ]; then
  timeout 600 python -m pytest tests/test_gpu_slabs.py -q -s > gpurun_out/r2_slab_tests_n${N}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_slab_tests_n${N}.txt; tail -4 gpurun_out/r2_slab_tests_n${N}.txt
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n${N}.err; head -c 400 gpurun_out/r2_bench_n${N}.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29722 \
    tests/run_config5.py --steps-per-phase 5 > gpurun_out/r2_c5_n${N}.log 2>&1
echo "c5 rc=$?"; tail -c 600 gpurun_out/r2_c5_n${N}.log
