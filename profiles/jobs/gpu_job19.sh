#!/bin/bash
# round 2, GPU job 19 (2 GPUs): slab parity with the C4 recipe and 4-rank coupled heat; bench at N = 2 with the C4-recipe parity check in its c4 block
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slabs.py -m gpu -q -s -k "c4 or 4-True or 2-False" > gpurun_out/r2_slab_tests_19.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_slab_tests_19.txt; grep -E "mgpu_slab_check|passed|failed|rc=" gpurun_out/r2_slab_tests_19.txt | cut -c1-330
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2_d.json 2> gpurun_out/r2_bench_n2_d.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n2_d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n2_d.json").read().strip().splitlines()[-1])
print("ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], "c4 ms %.2f"%d["c4"]["ms_per_step"], d["parity_check"]["ok"], d["c4"]["parity_check"])
PY
