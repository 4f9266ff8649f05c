#!/bin/bash
# round 2, GPU job 11 (1 GPU): whole GPU suite (device raster preparation, reference-rounding build), 8 ranks sharing the device
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --durations=8 > gpurun_out/r2_gpu_tests_11.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_11.txt; tail -14 gpurun_out/r2_gpu_tests_11.txt
SF3D_SHARE_DEVICE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 \
    tests/mgpu_slab_check.py > gpurun_out/r2_share8.txt 2>&1
echo "share8 rc=$?"; grep "mgpu_slab_check" gpurun_out/r2_share8.txt | tail -3
python - <<'PY'
import time, numpy as np
from criteria3d_b200.raster import prepare_on_device
for n in (1024, 4096, 8192):
    y, x = np.mgrid[0:n, 0:n].astype(np.float32)
    dem = (100 + 0.01 * x + 0.02 * y).astype(np.float32)
    prepare_on_device(dem[:64, :64], 10.0)
    t = time.time(); prepare_on_device(dem, 10.0); print(f"device raster prep {n}x{n}: {time.time() - t:.3f} s (host buffers in and out)")
PY
