"""GPU, >= 2 devices: the row-slab path (NCCL halo + all-reduces) against the oracle on the whole
catchment.  Skipped on a single-GPU box; the host-side partition logic is covered on CPU with gloo
(tests/test_partition.py)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world,heat", [(2, False), (4, False), (2, True)])
def test_slabs_match_oracle(world, heat):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world + (10 if heat else 0)),
           str(ROOT / "tests" / "mgpu_slab_check.py")] + (["--heat"] if heat else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "[mgpu_slab_check] ok" in r.stdout
