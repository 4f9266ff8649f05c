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


def main():
    ref = SoilFluxes3D(REFERENCE_LIB)
    assert ref.backend == "reference"
    for name, fn in sorted({**SCENARIOS, **HEAT_SCENARIOS}.items()):
        out = fn(ref)
        path = Path(__file__).parent / f"{name}.npz"
        np.savez_compressed(path, **out)
        print(f"{name}: {path.stat().st_size} bytes, {len(out['dts'])} steps")


if __name__ == "__main__":
    main()
