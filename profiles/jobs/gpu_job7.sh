#!/bin/bash
# round 2, GPU job 7 (2 GPUs): where do the ~22 us per sweep of the 2-rank fused sweep go?  timing variants (SF3D_MULTI_DEBUG:
# 1 = no peer stores, 2 = no early system fence; results are invalid for those, only kernel_ms is read)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 \
      bench.py --gpus 2 --steps 20 --warmup 3 --no-c4 --no-cpu-baseline --no-parity-check > gpurun_out/r2_exp7_$name.json 2> gpurun_out/r2_exp7_$name.err
  echo "$name rc=$?"
}
run base   SF3D_MULTI_DEBUG=0
run nostore SF3D_MULTI_DEBUG=1
run nofence SF3D_MULTI_DEBUG=2
run neither SF3D_MULTI_DEBUG=3
run unfused SF3D_FUSED_EXCHANGE=0
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-c4 --no-cpu-baseline > gpurun_out/r2_exp7_n1.json 2> gpurun_out/r2_exp7_n1.err
python - <<'PY'
import json
for f in ("n1","base","nostore","nofence","neither","unfused"):
    try:
        d=json.loads(open(f"gpurun_out/r2_exp7_{f}.json").read().strip().splitlines()[-1])
        k=d["kernel_ms"]; n=d["sweeps"]
        print(f"{f:8s} ms/step {d['ms_per_step']:.3f} sweeps {n} jacobi/sweep {1e3*k['jacobi']/n:.1f} us comm {k['comm']:.2f} asm {k['assemble']:.1f} post {k['post']:.1f} launches {d['gpu_launches']}")
    except Exception as e: print(f, "failed", e)
PY
