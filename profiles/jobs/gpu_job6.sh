#!/bin/bash
# round 2, GPU job 6 (2 GPUs): conditional system fence in the fused sweep: slab tests + 2-rank bench; 1-GPU bench in the same call for a like-for-like efficiency
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slabs.py -q > gpurun_out/r2_slab_tests_6.txt 2>&1; tail -3 gpurun_out/r2_slab_tests_6.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 \
    bench.py --gpus 2 --steps 40 --warmup 3 --no-c4 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err; echo "rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 40 --warmup 3 --no-c4 --no-cpu-baseline > gpurun_out/r2_bench_n1_same_box.json 2> gpurun_out/r2_bench_n1_same_box.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_n2_b","r2_bench_n1_same_box"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], {k:x for k,x in d["kernel_ms"].items() if x}, d["clocks"])
PY
