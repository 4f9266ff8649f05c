"""GPU: the -DSF3D_REFERENCE_ROUNDING build of the product (criteria3d_b200/libsf3d_b200_refround.so: pow() instead of
exp(y log x), area / distance and Math::computeMean evaluated per link as the reference writes them).  The DEFAULT build
is tolerance-parity: its few-ulp arithmetic differences could in principle flip one of the control loop's threshold tests
(Courant < 1.01, MBR < threshold, norm > 10 x best) on a long run and change the accepted-dt sequence from there on; no
such flip has been observed (tests/test_gpu_large_window.py, bench.py's accepted_dt against the reference arm, the 24 h
config-1 run), and this variant is the one to use when a caller needs the decision sequence itself pinned.  It is held
to the same lock-step test with TIGHTER state tolerances."""
from pathlib import Path

import numpy as np
import pytest

from criteria3d_b200 import BoundaryType, Field, SoilFluxes3D
from criteria3d_b200.synth import STORM_MM_H, Catchment, run_hours, setup

pytestmark = pytest.mark.gpu
LIB = Path(__file__).resolve().parent.parent / "criteria3d_b200" / "libsf3d_b200_refround.so"


def test_reference_rounding_build_lockstep_over_two_storm_hours(checker):
    if not LIB.exists():
        pytest.fail(f"{LIB} missing: run __graft_entry__.build()")
    sf = SoilFluxes3D(LIB)
    assert sf.backend == "b200"
    cat = Catchment(96, 96, 10)
    mm = STORM_MM_H[1:3]
    res = {}
    for name, lib in (("gpu", sf), ("ref", checker)):
        setup(lib, cat, threads=0)
        dts = run_hours(lib, cat, mm)
        res[name] = (dts, lib.get_field(Field.TOTAL_POTENTIAL, 0, cat.n_nodes), lib.get_field(Field.WATER_CONTENT, 0, cat.n_nodes),
                     lib.getTotalBoundaryWaterFlow(int(BoundaryType.Runoff)), lib.counters())
    (dg, Hg, Wg, rg, cg), (dr, Hr, Wr, rr, cr) = res["gpu"], res["ref"]
    assert dg == dr, "accepted time-step sequence"
    assert (cg["approximations"], cg["sweeps"]) == (cr["approximations"], cr["sweeps"])
    assert np.max(np.abs(Hg - Hr)) <= 1e-7                     # metres, absolute, heads ~230 m (default build: 1e-6 relative = 2e-4 m); measured 2.4e-8
    assert np.max(np.abs(Wg - Wr)) <= 1e-9
    assert rg == pytest.approx(rr, rel=1e-9, abs=1e-12)
    print(f"[refround] {len(dg)} steps, {cg['sweeps']} sweeps on both sides; max |dH| {np.max(np.abs(Hg - Hr)):.2e} m")
