#!/bin/bash
# round 2, GPU job 15 (2 GPUs): C5 slabs at N = 2 (the missing point of the 1/2/4/8 sweep), bench at N = 2 in the final format
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 \
    tests/run_config5.py --steps-per-phase 5 > gpurun_out/r2_c5_n2.log 2>&1
echo "c5 rc=$?"; tail -c 400 gpurun_out/r2_c5_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2_c.json 2> gpurun_out/r2_bench_n2_c.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n2_c.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_same_box_c.json 2> gpurun_out/r2_bench_n1_same_box_c.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_n2_c","r2_bench_n1_same_box_c"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e ms %.3f"%d["e2e"]["ms_per_step"], "c4 ms %.2f value %.4g"%(d["c4"]["ms_per_step"], d["c4"]["value"]), (d.get("parity_check") or {}).get("ok"), d["clocks"])
PY
