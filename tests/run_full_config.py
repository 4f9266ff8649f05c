"""Full-length run of a BASELINE.json configuration on the GPU (not collected by pytest):

    python tests/run_full_config.py 2        # C2: 1024x1024x(1+10), 6 h storm, water only
    python tests/run_full_config.py 3        # C3: the same with coupled heat (diffusive + latent), 6 h

Reports the second metric of BASELINE.json, simulated hours per wall-second over the computeStep loops
(forcing uploads as hourly rasters included), with steps, approximations, sweeps and node-iterations/s.
Writes gpurun_out/full_config<N>.json.  The product only: the reference needs hours for the same run
(its node-iterations/s on a bounded sample is the bench's reference arm)."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from criteria3d_b200 import BoundaryType, load_product  # noqa: E402
from criteria3d_b200.synth import STORM_MM_H, Catchment, set_heat_forcing, setup, setup_heat  # noqa: E402


def main():
    config = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    hours = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    heat = config == 3
    rows, cols = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1024, 1024)
    cat = Catchment(rows, cols, 10, heat=heat)
    sf = load_product()
    setup(sf, cat)
    if heat:
        setup_heat(sf, cat)
    mm = [STORM_MM_H[h % len(STORM_MM_H)] for h in range(hours)]
    sf.reset_counters()
    per_hour = []
    t_all = time.perf_counter()
    for h, rain in enumerate(mm):
        t0 = time.perf_counter()
        if heat:
            set_heat_forcing(sf, cat, h)
        assert sf.set_forcing_rasters(precipitation=cat.rain_raster(rain)) == 0
        t, n = 0.0, 0
        while t < 3600.0:
            t += sf.computeStep(3600.0 - t)
            n += 1
        sf.getTotalWaterContent()                      # one device sync per hour
        per_hour.append({"mm": rain, "steps": n, "wall_s": time.perf_counter() - t0})
    wall = time.perf_counter() - t_all
    c = sf.counters()
    out = {
        "config": f"C{config}: {rows}x{cols}x(1+10), {hours} h storm {mm} mm/h" + (", coupled heat (diffusive + latent)" if heat else ", water only"),
        "nodes": cat.n_nodes, "hours": hours, "wall_s": wall, "sim_hours_per_wall_s": hours / wall,
        "accepted_steps": int(c["steps"]), "tries": int(c["tries"]), "approximations": int(c["approximations"]),
        "sweeps": int(c["sweeps"]), "heat_steps": int(c["heat_steps"]), "heat_sweeps": int(c["heat_sweeps"]),
        "node_iterations_per_s": cat.n_nodes * float(c["sweeps"]) / wall, "ms_per_step": 1e3 * wall / max(1, int(c["steps"])),
        "last_step_mbr": float(c["last_mbr"]), "total_water_m3": sf.getTotalWaterContent(),
        "runoff_m3": sf.getTotalBoundaryWaterFlow(int(BoundaryType.Runoff)), "per_hour": per_hour,
    }
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"full_config{config}.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
