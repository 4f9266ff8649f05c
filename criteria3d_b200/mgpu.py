"""Harness plumbing for the row-slab multi-GPU path: one process per GPU (torchrun), NCCL unique id
broadcast through torch.distributed, slab catchment set up through the ordinary C ABI."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .capi import SoilFluxes3D
from .partition import Slab, make_slab, slab_catchment
from .synth import Catchment, _ok, setup


def wire_ranks(sf: SoilFluxes3D, rank: int, world: int, device: torch.device) -> None:
    """Create the library's NCCL communicator over the already initialised torch process group."""
    uid = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(sf.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0)
    _ok(sf.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes())), "sf3d_ext_comm_init")


def setup_slab(sf: SoilFluxes3D, rows: int, cols: int, n_soil_layers: int, rank: int, world: int,
               numerics=None, **cat_kw) -> tuple[Slab, Catchment]:
    """initialize3DModel's sequence on this rank's slab (owned rows + ghost rows), then the halo lists.
    The balance is initialised after the ghosts are known so that storage counts owned nodes only."""
    slab = make_slab(rows, cols, n_soil_layers, world, rank)
    cat = slab_catchment(slab, **cat_kw)
    setup(sf, cat, numerics=numerics)
    if world > 1:
        peers, send, recv = slab.halo()
        _ok(sf.set_halo(peers, send, recv, slab.n_global), "sf3d_ext_set_halo")
        _ok(sf.initializeBalance(), "initializeBalance")
    return slab, cat
