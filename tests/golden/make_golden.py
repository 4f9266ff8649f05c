"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsf3d_ref.so, built from
/root/reference by oracle/Makefile).  Run in the build container, where the reference exists:

    python tests/golden/make_golden.py

Each file holds the outputs of one scenario of tests/scenarios.py (1 OpenMP thread): potentials,
water contents, conductivities, flow sums, accepted time steps, counters, balance scalars.  The
inputs are the seeded generator itself (criteria3d_b200/synth.py), so the fixtures stay small."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from criteria3d_b200 import REFERENCE_LIB, SoilFluxes3D  # noqa: E402
from scenarios import HEAT_SCENARIOS, SCENARIOS  # noqa: E402


def config1_inputs():
    """rasters of the bundled STH sample project -> tests/golden/config1_sth_inputs.npz"""
    from criteria3d_b200.raster import read_flt
    base = Path("/root/reference/DATA/PROJECT/STH")
    dem = read_flt(base / "MAPS" / "DEM_STH.flt")
    soil = read_flt(base / "SOIL" / "soilMap_STH.flt")
    np.savez_compressed(Path(__file__).parent / "config1_sth_inputs.npz", dem=dem.values, soil=soil.values,
                        cell=np.float64(dem.cell), xll=np.float64(dem.xll), yll=np.float64(dem.yll))


def main():
    config1_inputs()
    ref = SoilFluxes3D(REFERENCE_LIB)
    assert ref.backend == "reference"
    for name, fn in sorted({**SCENARIOS, **HEAT_SCENARIOS}.items()):
        out = fn(ref)
        path = Path(__file__).parent / f"{name}.npz"
        np.savez_compressed(path, **out)
        print(f"{name}: {path.stat().st_size} bytes, {len(out['dts'])} steps")


if __name__ == "__main__":
    main()
