"""GPU: the row-slab path (peer-memory halo inside the sweep kernel + in-kernel all-reduces; NCCL fallback) against
the oracle on the whole catchment.  With fewer devices than ranks the ranks SHARE cuda:0 (SF3D_SHARE_DEVICE=1: gloo
process group, CUDA-IPC peer memory of the same device, no NCCL), so the multi-rank control flow, the halo lists,
the mailbox all-reduce and its time-out are exercised on a one-GPU box too; the host-side partition logic is
covered on CPU with gloo (tests/test_partition.py)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _torchrun(world, port, *args, env_extra=None, timeout=900):
    import torch
    env = dict(os.environ)
    if torch.cuda.device_count() < world:
        env["SF3D_SHARE_DEVICE"] = "1"
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "mgpu_slab_check.py"), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.parametrize("world,heat", [(2, False), (4, False), (8, False), (2, True), (4, True), (4, "c4"), (2, "heat-all")])
def test_slabs_match_oracle(world, heat):
    """water storm, coupled heat, and the C4 recipe (20 soil layers, saturated lower third, free drainage) at small size"""
    args = {"c4": ["--c4"], "heat-all": ["--heat", "--save-all"]}.get(heat, ["--heat"] if heat else [])
    r = _torchrun(world, 29600 + world + {"c4": 20, "heat-all": 30}.get(heat, 10 if heat else 0), *args)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "[mgpu_slab_check] ok" in r.stdout
    print(r.stdout.strip().splitlines()[-1])


def test_silent_peer_is_an_error_not_a_numerical_event():
    """VERDICT r1 / ADVICE r1: a mailbox time-out must come back as SolverError from computeStep"""
    r = _torchrun(2, 29631, "--timeout-test", env_extra={"SF3D_MAILBOX_TIMEOUT_S": "1.5"}, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "[mgpu_slab_check] timeout ok" in r.stdout
