#!/bin/bash
# round 2, GPU job 8 (2 GPUs): second round of timing variants of the fused sweep (SF3D_MULTI_DEBUG: 4 = no boundary-first loop,
# 8 = no cross-rank all-reduce at all; results invalid, only kernel_ms is read)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 \
      bench.py --gpus 2 --steps 20 --warmup 3 --no-c4 --no-cpu-baseline --no-parity-check > gpurun_out/r2_exp8_$name.json 2> gpurun_out/r2_exp8_$name.err
  echo "$name rc=$?"
}
run base   SF3D_MULTI_DEBUG=0
run noboundary SF3D_MULTI_DEBUG=4
run noreduce SF3D_MULTI_DEBUG=8
run nothing SF3D_MULTI_DEBUG=15
python - <<'PY'
import json
for f in ("base","noboundary","noreduce","nothing"):
    try:
        d=json.loads(open(f"gpurun_out/r2_exp8_{f}.json").read().strip().splitlines()[-1])
        k=d["kernel_ms"]; n=d["sweeps"]
        print(f"{f:10s} ms/step {d['ms_per_step']:.3f} sweeps {n} jacobi/sweep {1e3*k['jacobi']/n:.1f} us comm {k['comm']:.2f} asm {k['assemble']:.1f} post {k['post']:.1f} launches {d['gpu_launches']}")
    except Exception as e: print(f, "failed", e)
PY
