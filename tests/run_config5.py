"""BASELINE config 5 (8192 x 8192 x 20 layers, 72 h multi-event storm) as a weak-scaling sweep: every GPU owns a row slab of
1024 DEM rows x 8192 columns x (1+20) layers = 176 M nodes (~101 GB of the 180 GB HBM), so N = 1 / 2 / 4 / 8 GPUs simulate a
1024N x 8192 catchment and N = 8 is the named 8192 x 8192 grid.  Not collected by pytest; run under torchrun for N > 1:

    python tests/run_config5.py [--steps-per-phase K] [--rows 1024 --cols 8192 --soil-layers 20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/run_config5.py

The 72 h schedule is three storms (the C2 hyetograph 5-20-40-25-10-2 mm/h at hours 0, 24 and 48) separated by dry gaps with a
surface evaporation sink of 0.1 mm/h.  A full run is ~20 000 accepted steps of ~0.25 s: hours of box time.  What is
measured here is a COMPRESSED PREFIX of that schedule, labelled as such: the forcing phases in order (storm 1 at 5, 20, 40 mm/h;
dry gap with evaporation; storm 2 at 5, 20 mm/h), each cut after K accepted steps, the state carried from phase to phase.  It
exercises everything the full run does (rain, the evaporation clamp water.cpp:645-652, re-wetting) at the full per-GPU size;
the 72 h estimate extrapolates the measured time per accepted step with the steps-per-simulated-hour of the C2 run
(profiles/r01_full_config2_c2_6h.json: 7 060 steps in 6 storm hours) and is an ESTIMATE.
Writes gpurun_out/config5_n<N>.json (rank 0)."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

PHASES = [("storm 1, 5 mm/h", 5.0), ("storm 1, 20 mm/h", 20.0), ("storm 1, 40 mm/h", 40.0), ("dry gap, evaporation 0.1 mm/h", -0.1),
          ("storm 2, 5 mm/h", 5.0), ("storm 2, 20 mm/h", 20.0)]


def main():
    import torch
    import torch.distributed as dist
    from criteria3d_b200 import BoundaryType, Field, load_product
    from criteria3d_b200.synth import Catchment, setup

    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1024)
    ap.add_argument("--cols", type=int, default=8192)
    ap.add_argument("--soil-layers", type=int, default=20)
    ap.add_argument("--steps-per-phase", type=int, default=6)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sf = load_product()
    assert sf.set_device(local) == 0
    t_setup = time.perf_counter()
    if world > 1:
        from criteria3d_b200.mgpu import setup_slab, wire_ranks
        wire_ranks(sf, rank, world)
        slab, cat = setup_slab(sf, a.rows * world, a.cols, a.soil_layers, rank, world)
        n_owned = slab.n_owned
    else:
        cat = Catchment(a.rows, a.cols, a.soil_layers)
        setup(sf, cat)
        n_owned = cat.n_nodes
    t_setup = time.perf_counter() - t_setup
    stream = torch.cuda.ExternalStream(sf.stream(), device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    phases = []
    sim_total = 0.0
    for name, mm in PHASES:
        # rain as the hourly precipitation map; the dry gap as a surface sink map (layer 0 of the per-layer sink maps, mm/h removed)
        if mm >= 0:
            assert sf.set_forcing_rasters(precipitation=cat.rain_raster(mm)) == 0
        else:
            evap = np.full((1, cat.rows, cat.cols), -mm, np.float32)
            assert sf.set_forcing_rasters(precipitation=np.zeros((cat.rows, cat.cols), np.float32), layer_sink=evap) == 0
        c0 = sf.counters()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        sim = 0.0
        for _ in range(a.steps_per_phase):
            sim += sf.computeStep(3600.0)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        c1 = sf.counters()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        sim_total += sim
        phases.append({"phase": name, "mm_per_h": mm, "accepted_steps": a.steps_per_phase, "simulated_s": sim, "ms": ms,
                       "ms_per_step": ms / a.steps_per_phase, "sweeps": int(c1["sweeps"] - c0["sweeps"]),
                       "approximations": int(c1["approximations"] - c0["approximations"]), "tries": int(c1["tries"] - c0["tries"]),
                       "last_mbr": float(c1["last_mbr"])})
    owned = torch.tensor([float(n_owned)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(owned)
    tw = sf.getTotalWaterContent()
    runoff = sf.getTotalBoundaryWaterFlow(int(BoundaryType.Runoff))
    if rank == 0:
        tot_ms = sum(p["ms"] for p in phases)
        tot_sweeps = sum(p["sweeps"] for p in phases)
        steps = sum(p["accepted_steps"] for p in phases)
        ms_step = tot_ms / steps
        out = {
            "config": f"C5 weak scaling: {world} slab(s) of {a.rows}x{a.cols} DEM x (1+{a.soil_layers}) layers = {a.rows * world}x{a.cols} catchment"
                      + (" = BASELINE configs[4] grid" if (a.rows * world, a.cols, a.soil_layers) == (8192, 8192, 20) else ""),
            "n_gpus": world, "nodes_per_gpu": cat.n_nodes, "owned_nodes_total": float(owned[0]), "setup_s": t_setup,
            "what": "compressed prefix of the 72 h three-storm schedule: every forcing phase cut after K accepted steps, state carried over",
            "phases": phases, "accepted_steps": steps, "sweeps": tot_sweeps, "ms_per_step": ms_step,
            "node_iterations_per_s": float(owned[0]) * tot_sweeps / (tot_ms * 1e-3), "simulated_s": sim_total,
            "total_water_m3": tw, "runoff_m3": runoff, "halo": getattr(sf, "halo_mode", "single GPU"),
            "estimate_72h": {"assumption": "3 storms x 7 060 accepted steps (steps of the C2 6 h storm run) + 54 dry hours x ~4 steps, at the measured mean time per accepted step",
                             "accepted_steps": 3 * 7060 + 54 * 4, "wall_hours": (3 * 7060 + 54 * 4) * ms_step * 1e-3 / 3600.0, "kind": "ESTIMATE, not measured"},
        }
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / f"config5_n{world}.json").write_text(json.dumps(out, indent=1))
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
