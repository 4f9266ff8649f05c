"""Run under torchrun (one rank per GPU): the row-slab product against the oracle on the WHOLE
catchment.  Every rank runs the CPU oracle on the full (small) catchment and checks its owned nodes.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29611 tests/mgpu_slab_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from criteria3d_b200 import BoundaryType, Field, SoilFluxes3D, load_product  # noqa: E402
from oracle import ORACLE_LIB, REFERENCE_LIB  # noqa: E402
from criteria3d_b200.mgpu import setup_slab, wire_ranks  # noqa: E402
from criteria3d_b200.synth import Catchment, run_hours, set_heat_forcing, setup  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    heat = "--heat" in sys.argv
    R, C, L = (24, 16, 4) if heat else (48, 40, 5)
    hours, max_steps = ([0.0, 10.0], 10) if heat else ([20.0, 40.0], 50)

    gpu = load_product()
    assert gpu.set_device(local) == 0
    wire_ranks(gpu, rank, world, dev)
    slab, lc = setup_slab(gpu, R, C, L, rank, world, heat=heat)

    def run(sf, cat):
        dts = []
        for h, mm in enumerate(hours):
            if heat:
                set_heat_forcing(sf, cat, 10 + h)
            dts += run_hours(sf, cat, [mm], max_steps=max_steps)
        return dts
    dts = run(gpu, lc)

    chk = SoilFluxes3D(REFERENCE_LIB if REFERENCE_LIB.exists() else ORACLE_LIB)
    cat = Catchment(R, C, L, heat=heat)
    setup(chk, cat, threads=1)
    dts_ref = run(chk, cat)

    assert dts == dts_ref, f"rank {rank}: accepted steps differ\n{dts}\n{dts_ref}"
    own, l2g = slab.owned_mask(), slab.local_to_global()
    fields = [(Field.TOTAL_POTENTIAL, 1e-6), (Field.WATER_CONTENT, 1e-7), (Field.DEGREE_OF_SATURATION, 1e-6)]
    if heat:
        fields.append((Field.TEMPERATURE, 1e-6))
    for f, tol in fields:
        a = gpu.get_field(f, 0, lc.n_nodes)[own]
        b = chk.get_field(f, 0, cat.n_nodes)[l2g[own]]
        if f == Field.TEMPERATURE:                      # surface nodes carry the TopographyError sentinel on both sides
            keep = l2g[own] >= cat.n_surface
            a, b = a[keep], b[keep]
        err = np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))
        assert err <= tol, f"rank {rank}: {f.name} err {err}"
    tw, tw_ref = gpu.getTotalWaterContent(), chk.getTotalWaterContent()
    assert abs(tw - tw_ref) <= 1e-9 * abs(tw_ref), (tw, tw_ref)
    for bt in (BoundaryType.Runoff, BoundaryType.FreeDrainage, BoundaryType.FreeLateralDrainage):
        a, b = gpu.getTotalBoundaryWaterFlow(int(bt)), chk.getTotalBoundaryWaterFlow(int(bt))
        assert abs(a - b) <= 1e-6 * abs(b) + 1e-12, (bt, a, b)
    cg, cr = gpu.counters(), chk.counters()
    assert cg["approximations"] == cr["approximations"], (cg["approximations"], cr["approximations"])
    dist.barrier()
    if rank == 0:
        print(f"[mgpu_slab_check] ok: world={world}, heat={heat}, {len(dts)} steps, {cg['sweeps']} sweeps, total water {tw:.6f}")
    gpu.comm_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
