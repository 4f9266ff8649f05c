"""BASELINE config 1 end to end: the bundled STH catchment (34 x 139 cells of 2 m, 5 soil layers, 17 412 nodes), 24 h of
2 mm/h rain, water only (SURVEY 8d "C1").  Not collected by pytest.

    python tests/run_config1.py reference [threads] [hours]   # oracle/_ref on the host; writes the golden final state
    python tests/run_config1.py product [hours]               # the CUDA product; compares with the golden final state

The reference run writes tests/golden/config1_24h_final.npz (accepted-dt sequence, final potentials / water contents,
boundary totals, counters, wall time); the product run reports wall time, steps per second and parity against it and
writes gpurun_out/config1_24h_product.json."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from criteria3d_b200 import BoundaryType, Field, SoilFluxes3D, load_product  # noqa: E402
from criteria3d_b200.synth import Catchment, run_hours, setup  # noqa: E402

GOLDEN = ROOT / "tests" / "golden" / "config1_24h_final.npz"
BTS = (BoundaryType.Runoff, BoundaryType.FreeDrainage, BoundaryType.FreeLateralDrainage)


def catchment():
    with np.load(ROOT / "tests" / "golden" / "config1_sth_inputs.npz") as z:
        dem, soil, cell = z["dem"], z["soil"], float(z["cell"])
    valid = dem != np.float32(-9999)
    return Catchment(dem.shape[0], dem.shape[1], 5, cell=cell, valid=valid, dem_override=np.where(valid, dem, 0).astype(np.float32),
                     soil_override=np.where(valid, soil, 1).astype(np.uint16))


def run(sf, hours, threads):
    cat = catchment()
    setup(sf, cat, threads=threads)
    t0 = time.perf_counter()
    dts = run_hours(sf, cat, [2.0] * hours)
    H = sf.get_field(Field.TOTAL_POTENTIAL, 0, cat.n_nodes)        # includes the final device sync
    wall = time.perf_counter() - t0
    W = sf.get_field(Field.WATER_CONTENT, 0, cat.n_nodes)
    c = sf.counters()
    return dict(dts=np.asarray(dts), H=H, W=W, boundary=np.array([sf.getTotalBoundaryWaterFlow(int(b)) for b in BTS]),
                total_water=np.float64(sf.getTotalWaterContent()), counters=np.array([c["steps"], c["approximations"], c["sweeps"]], np.float64),
                wall_s=np.float64(wall), nodes=np.int64(cat.n_nodes))


def main():
    who = sys.argv[1] if len(sys.argv) > 1 else "product"
    if who == "reference":
        from oracle import REFERENCE_LIB
        threads = int(sys.argv[2]) if len(sys.argv) > 2 else 1
        hours = int(sys.argv[3]) if len(sys.argv) > 3 else 24
        r = run(SoilFluxes3D(REFERENCE_LIB), hours, threads)
        r["threads"], r["hours"] = np.int64(threads), np.int64(hours)
        if hours == 24 and threads == 1:
            np.savez_compressed(GOLDEN, **r)
        print(json.dumps({"who": "reference", "threads": threads, "hours": hours, "wall_s": float(r["wall_s"]), "steps": len(r["dts"]),
                          "approximations": int(r["counters"][1]), "sweeps": int(r["counters"][2]),
                          "ms_per_step": 1e3 * float(r["wall_s"]) / len(r["dts"])}))
        return
    hours = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    r = run(load_product(), hours, 0)
    out = {"who": "product", "hours": hours, "nodes": int(r["nodes"]), "wall_s": float(r["wall_s"]), "steps": len(r["dts"]),
           "approximations": int(r["counters"][1]), "sweeps": int(r["counters"][2]), "ms_per_step": 1e3 * float(r["wall_s"]) / len(r["dts"]),
           "sim_hours_per_wall_s": hours / float(r["wall_s"])}
    if GOLDEN.exists() and hours == 24:
        with np.load(GOLDEN) as g:
            n = min(len(g["dts"]), len(r["dts"]))
            first = next((k for k in range(n) if g["dts"][k] != r["dts"][k]), None)
            out.update({"reference_wall_s_1_thread_build_container": float(g["wall_s"]), "reference_steps": int(len(g["dts"])),
                        "dt_sequence_equal": bool(first is None and len(g["dts"]) == len(r["dts"])), "first_divergent_step": first,
                        "max_abs_dH": float(np.max(np.abs(g["H"] - r["H"]))), "max_abs_dtheta": float(np.max(np.abs(g["W"] - r["W"]))),
                        "boundary_totals": [r["boundary"].tolist(), g["boundary"].tolist()],
                        "total_water": [float(r["total_water"]), float(g["total_water"])],
                        "approximations_sweeps": [r["counters"][1:].tolist(), g["counters"][1:].tolist()]})
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "config1_24h_product.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
