#!/usr/bin/env python
"""bench.py -- node-iterations/s of the soilFluxes3D water time step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA product (default N=1)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU/OpenMP path

A "step" is one computeStep() (soilFluxes3D.cpp:1785) = one accepted water time step of the
catchment: Picard approximations, each an assembly pass and a batch of Jacobi sweeps.  The
headline unit is node-iterations/s = nodes x Jacobi sweeps executed / time; simulated
hours per wall second are reported beside it.

  value : inputs resident in HBM, device time (CUDA events on the library's stream, max over ranks)
  e2e   : the same metric through the C ABI with HOST buffers: every step uploads the forcing
          as the hourly precipitation map (sf3d_ext_set_forcing_rasters = assignPrecipitation +
          setSinkSource, pinned host memory) and reads back the matric potential maps of all
          layers (sf3d_ext_get_layer_rasters = computeCriteria3DMap per layer, as saveModelsState);
          host<->device copies are inside the timed region
  roofline : Jacobi sweep kernel, algorithmic bytes (12 B per link + 32 B per node) / measured
          kernel time (CUDA events around every launch, same timed region) vs the measured HBM
          copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline : the reference (oracle/_ref, unmodified sources) on the box's host cores, bounded
          sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = dict(rows=1024, cols=1024, soil_layers=10)          # BASELINE.json configs[1]
RAIN_MM_H = 40.0                                                # peak hour of the C2 hyetograph
# bounded samples of the same generator for the CPU legs, large enough (1.9 / 4.2 GB of reference state) to be
# DRAM-resident like the full 7.5 GB workload rather than cache-resident, sized for 10-30 s of CPU work on 16 cores
CPU_SAMPLE = dict(rows=512, cols=512, soil_layers=10)
REF_ARM_SAMPLE = dict(rows=768, cols=768, soil_layers=10)


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.device)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in Path(self.path).read_text().strip().splitlines() if r.strip()]
            sm = [float(r[1]) for r in rows]
            out["sm_mhz"] = float(np.median(sm)) if sm else None
            out["sm_max_mhz"] = float(rows[0][2]) if rows else None
            out["power_w_max"] = max(float(r[3]) for r in rows) if rows else None
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            seen = set()
            for r in rows:
                for k, nme in enumerate(names):
                    if r[4 + k].strip().lower().startswith("active"):
                        seen.add(nme)
            out["reasons"] = sorted(seen)
            out["samples"] = len(rows)
        except Exception as e:  # noqa: BLE001
            out["error"] = str(e)
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


def cpu_run(sample: dict, steps: int, warmup: int, budget_s: float):
    """Time the reference CPU implementation (all host threads) on a bounded sample."""
    from criteria3d_b200 import Field, SoilFluxes3D
    from oracle import ORACLE_LIB, REFERENCE_LIB
    from criteria3d_b200.synth import Catchment, setup
    if REFERENCE_LIB.exists():
        sf, kind = SoilFluxes3D(REFERENCE_LIB), "reference"
    else:
        sf, kind = SoilFluxes3D(ORACLE_LIB), "port"
    cat = Catchment(sample["rows"], sample["cols"], sample["soil_layers"])
    setup(sf, cat, threads=0)
    cores = sf.setThreadsNumber(0)
    sink = np.zeros(cat.n_nodes)
    sink[: cat.n_surface] = cat.rain_sink_source(RAIN_MM_H)
    sf.set_field(Field.WATER_SINK_SOURCE, 0, sink)
    for _ in range(warmup):
        sf.computeStep(3600.0)
    c0 = sf.counters()
    t0 = time.perf_counter()
    sim = 0.0
    done = 0
    for _ in range(steps):
        sim += sf.computeStep(3600.0)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    wall = time.perf_counter() - t0
    c1 = sf.counters()
    sweeps = c1["sweeps"] - c0["sweeps"]
    return {
        "value": cat.n_nodes * sweeps / wall, "unit": "node-iterations/s", "cores": int(cores), "kind": kind,
        "sample": f"{sample['rows']}x{sample['cols']}x(1+{sample['soil_layers']}) window of the same generator, "
                  f"{done} computeStep calls, {sweeps} sweeps, {wall:.2f} s wall",
        "sim_hours_per_wall_s": sim / 3600.0 / wall, "ms_per_step": 1e3 * wall / max(done, 1),
        "steps": done, "sweeps": int(sweeps), "approximations": int(c1["approximations"] - c0["approximations"]),
        "n_nodes": cat.n_nodes,
    }


def config_name(args, world) -> str:
    """BASELINE.json configuration the shape corresponds to (C2 is the headline single-GPU workload)"""
    shape = (args.rows * world, args.cols, args.soil_layers)
    if args.heat and shape == (1024, 1024, 10):
        return "C3"
    return {(1024, 1024, 10): "C2", (4096, 4096, 20): "C4", (8192, 8192, 20): "C5"}.get(shape, "C2-like slab" if world > 1 else "custom")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=WORKLOAD["rows"])
    ap.add_argument("--cols", type=int, default=WORKLOAD["cols"])
    ap.add_argument("--soil-layers", type=int, default=WORKLOAD["soil_layers"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--saturated-bottom", action="store_true", help="config 4: lower third of the layers start saturated")
    ap.add_argument("--heat", action="store_true", help="config 3: coupled heat transport (not the headline workload)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_run(REF_ARM_SAMPLE, args.steps, args.warmup, budget_s=150.0)
        line = {
            "impl": "reference", "metric": "node-iterations/s", "value": r["value"], "unit": "node-iterations/s",
            "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2 synthetic 1024x1024 DEM x (1+10) layers, 40 mm/h storm hour, water only "
                                   "(reference timed on a bounded sample: " + r["sample"] + ")"},
            "sim_hours_per_wall_s": r["sim_hours_per_wall_s"],
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "node-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    from criteria3d_b200 import Field, load_product
    from criteria3d_b200.synth import Catchment, setup

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sf = load_product()
    assert sf.set_device(local_rank) == 0
    if world > 1:
        # weak scaling: every GPU owns a rows x cols slab of a (world*rows) x cols catchment
        from criteria3d_b200.mgpu import setup_slab, wire_ranks
        wire_ranks(sf, rank, world, torch.device("cuda", local_rank))
        slab, cat = setup_slab(sf, args.rows * world, args.cols, args.soil_layers, rank, world,
                               saturated_bottom=args.saturated_bottom)
        n_owned = slab.n_owned
    else:
        cat = Catchment(args.rows, args.cols, args.soil_layers, heat=args.heat, saturated_bottom=args.saturated_bottom)
        setup(sf, cat)
        n_owned = cat.n_nodes
    N = cat.n_nodes
    stream = torch.cuda.ExternalStream(sf.stream(), device=local_rank)

    sink_host = torch.zeros(N, dtype=torch.float64).pin_memory()
    sink_np = sink_host.numpy()
    sink_np[: cat.n_surface] = cat.rain_sink_source(RAIN_MM_H)
    # end-to-end output: the matric potential maps of all layers, float32 rasters as saveModelsState /
    # computeCriteria3DMap produce them
    out_host = torch.empty((cat.layers, cat.rows, cat.cols), dtype=torch.float32).pin_memory()
    out_np = out_host.numpy()
    assert sf.set_field(Field.WATER_SINK_SOURCE, 0, sink_np) == 0
    # end-to-end input: the hourly precipitation map [mm h-1], float32 like the reference's meteo maps
    rain_host = torch.from_numpy(cat.rain_raster(RAIN_MM_H)).pin_memory()
    rain_np = rain_host.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        sf.computeStep(3600.0)

    # ---------------- timed region 1: inputs resident in HBM --------------------------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    sf.profile(True)
    c0 = sf.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    sim = 0.0
    for _ in range(args.steps):
        sim += sf.computeStep(3600.0)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    c1 = sf.counters()
    ktimes = sf.kernel_times()
    sf.profile(False)

    # ---------------- timed region 2: end to end through the C ABI, host buffers ---------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sf.set_forcing_rasters(precipitation=rain_np)                       # untimed warm-up of the host-facing calls
    sf.get_layer_rasters(Field.MATRIC_POTENTIAL, 0, cat.layers, (cat.rows, cat.cols), out=out_np)
    c1e = sf.counters()
    tcall = [0.0, 0.0, 0.0]
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    sim_e2e = 0.0
    for _ in range(args.steps):
        ta = time.perf_counter()
        sf.set_forcing_rasters(precipitation=rain_np)                   # H2D, rows x cols floats -> sink/source on the device
        tb = time.perf_counter()
        sim_e2e += sf.computeStep(3600.0)
        tc = time.perf_counter()
        sf.get_layer_rasters(Field.MATRIC_POTENTIAL, 0, cat.layers, (cat.rows, cat.cols), out=out_np)   # D2H, layers x rows x cols floats
        td = time.perf_counter()
        tcall[0] += tb - ta; tcall[1] += tc - tb; tcall[2] += td - tc
    e1.record(stream)
    barrier()
    wall_e2e = time.perf_counter() - t0
    ms_e2e = max(e0.elapsed_time(e1), wall_e2e * 1e3)
    c2 = sf.counters()
    clk = clocks.stop()

    sweeps = c1["sweeps"] - c0["sweeps"]
    sweeps_e2e = c2["sweeps"] - c1e["sweeps"]
    t = torch.tensor([ms, ms_e2e, float(sweeps), float(sweeps_e2e), sim, sim_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ms_e2e = float(tmax[0]), float(tmax[1])
        owned = torch.tensor([float(n_owned)], dtype=torch.float64, device="cuda")
        dist.all_reduce(owned)
        # every rank executes the same sweeps on its slab: node-iterations = global owned nodes x sweeps
        tot_iter = float(owned[0]) * float(sweeps)
        tot_iter_e2e = float(owned[0]) * float(sweeps_e2e)
    else:
        tot_iter, tot_iter_e2e = N * float(sweeps), N * float(sweeps_e2e)

    if rank == 0:
        peak, peak_src = measured_peak()
        links = c1["links"]
        bytes_sweep = 12.0 * links + 32.0 * N
        jac = ktimes["jacobi"]
        jac_gbs = bytes_sweep * sweeps / (jac["ms"] * 1e-3) / 1e9 if jac["ms"] > 0 else None
        # assembly (node phase + link phase): SURVEY 8d algorithmic bytes 144 N + 12 Lk + 8 nnz per approximation
        approx = int(c1["approximations"] - c0["approximations"])
        bytes_asm = 144.0 * N + 12.0 * links + 8.0 * links
        asm_ms = ktimes["assemble"]["ms"] + ktimes["node_phase"]["ms"]
        asm_gbs = bytes_asm * approx / (asm_ms * 1e-3) / 1e9 if asm_ms > 0 and approx else None
        traffic = None
        tfile = ROOT / "profiles" / "jacobi_traffic.json"
        if tfile.exists():
            try:
                tj = json.loads(tfile.read_text())
                # the capture is of the C2 single-GPU shape; other shapes have no capture
                traffic = tj.get("dram_bytes_per_launch") if int(tj.get("nodes", 0)) == int(N) else None
            except Exception:
                traffic = None
        line = {
            "metric": "node-iterations/s", "value": tot_iter / (ms * 1e-3), "unit": "node-iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{config_name(args, world)} synthetic {args.rows}x{args.cols} DEM per GPU x (1+{args.soil_layers}) layers, "
                            f"{RAIN_MM_H:g} mm/h storm hour, " + ("coupled heat (diffusive + latent)" if args.heat else "water only") + ", Richards + Manning runoff",
                "nodes_per_gpu": N, "links_per_gpu": int(links),
                "parallelism": "single GPU" if world == 1 else
                               f"{world} row slabs of {args.rows} DEM rows each (+1 ghost row per side); per sweep: "
                               + ("boundary rows stored into the neighbours' ghost rows over NVLink peer memory + mailbox all-reduce of the "
                                  "residual + stopping rule, one fused kernel" if getattr(sf, "halo_mode", "nccl") == "peer-memory"
                                  else "ncclSend/ncclRecv halo of x + ncclAllReduce of the residual")
                               + f"; all-reduce of Courant / balance sums per approximation; global catchment {args.rows * world}x{args.cols}",
                "l2": "working set per sweep (12 B/link + 32 B/node = %.2f GB) >> 126 MB L2; no explicit flush" % (bytes_sweep / 1e9),
                "numerics": "setNumericalParameters(0.5, 3600, 150, 10, 10, 3)",
            },
            "sim_hours_per_wall_s": sim / 3600.0 / (ms * 1e-3),
            "sweeps": int(sweeps), "approximations": int(c1["approximations"] - c0["approximations"]),
            "heat_steps": int(c1["heat_steps"] - c0["heat_steps"]), "heat_sweeps": int(c1["heat_sweeps"] - c0["heat_sweeps"]),
            "tries": int(c1["tries"] - c0["tries"]),
            "e2e": {"value": tot_iter_e2e / (ms_e2e * 1e-3), "unit": "node-iterations/s",
                    "h2d_bytes_per_step": int(rain_np.nbytes), "d2h_bytes_per_step": int(out_np.nbytes),
                    "ms_per_step": ms_e2e / args.steps,
                    "host_ms_per_step": {"forcing_upload": 1e3 * tcall[0] / args.steps, "compute_step": 1e3 * tcall[1] / args.steps,
                                         "result_download": 1e3 * tcall[2] / args.steps}, "sim_hours_per_wall_s": sim_e2e / 3600.0 / (ms_e2e * 1e-3)},
            "gpu_launches": int(c1["kernel_launches"] - c0["kernel_launches"]),
            "roofline": {"kernel": "kern_jacobi", "bound": "hbm", "achieved": jac_gbs, "peak": peak, "unit": "GB/s",
                         "frac": (jac_gbs / peak) if jac_gbs else None, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_launch": bytes_sweep, "avg_launch_ms": jac["ms"] / max(sweeps, 1),
                         "launches": jac["launches"], "executed_sweeps": int(sweeps),
                         "frac_of_nominal_8000": (jac_gbs / 8000.0) if jac_gbs else None,
                         "frac_traffic": (traffic / (jac["ms"] / max(sweeps, 1) * 1e-3) / 1e9 / peak) if (traffic and jac["ms"] > 0) else None,
                         "note": "achieved counts ALGORITHMIC bytes; the kernel moves fewer (traffic) because column "
                                 "indices are pattern-compressed, so frac can exceed 1; frac_traffic = measured DRAM bytes (ncu) / time / peak"},
            "roofline_assembly": {"kernel": "kern_node_phase + kern_assemble", "bound": "fp64 issue (HBM reported)",
                                  "achieved": asm_gbs, "peak": peak, "unit": "GB/s", "frac": (asm_gbs / peak) if asm_gbs else None,
                                  "bytes_per_approximation": bytes_asm, "avg_ms_per_approximation": asm_ms / max(approx, 1)},
            "kernel_ms": {k: round(v["ms"], 3) for k, v in ktimes.items()},
            "kernel_share": {k: round(v["ms"] / max(ms, 1e-9), 4) for k, v in ktimes.items()},
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = cpu_run(CPU_SAMPLE, steps=20, warmup=1, budget_s=25.0)
                line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
                line["cpu_baseline"]["sim_hours_per_wall_s"] = r["sim_hours_per_wall_s"]
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "error": str(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
