"""GPU: parity on a 128 x 128 x (1+10) window of the C2 generator (180 224 nodes) over two storm hours against the
unmodified reference on all host cores: identical accepted-dt sequence, approximation and sweep counts, potentials,
water contents, boundary totals and the step's mass-balance error.  (Round 1 ran this as a script on a 256 x 256
window -- profiles/r01_parity_large.json; `python tests/validate_large.py 256 256 10 3` still does.)"""
import numpy as np
import pytest

from criteria3d_b200 import BoundaryType, Field
from criteria3d_b200.synth import STORM_MM_H, Catchment, run_hours, setup

pytestmark = pytest.mark.gpu


def test_c2_window_two_storm_hours(product, checker):
    cat = Catchment(128, 128, 10)
    mm = STORM_MM_H[1:3]                   # 20 and 40 mm/h
    res = {}
    for name, sf in (("gpu", product), ("ref", checker)):
        setup(sf, cat, threads=0)          # reference: all host cores
        dts = run_hours(sf, cat, mm)
        res[name] = (dts, sf.get_field(Field.TOTAL_POTENTIAL, 0, cat.n_nodes), sf.get_field(Field.WATER_CONTENT, 0, cat.n_nodes),
                     [sf.getTotalBoundaryWaterFlow(int(b)) for b in (BoundaryType.Runoff, BoundaryType.FreeDrainage, BoundaryType.FreeLateralDrainage)],
                     sf.getTotalWaterContent(), sf.counters())
    (dg, Hg, Wg, bg, tg, cg), (dr, Hr, Wr, br, tr, cr) = res["gpu"], res["ref"]
    assert len(dg) > 150, "the window is expected to need a few hundred accepted steps"
    first = next((k for k, (x, y) in enumerate(zip(dg, dr)) if x != y), None)
    assert dg == dr, f"accepted time-step sequences differ from step {first}"
    assert cg["approximations"] == cr["approximations"] and cg["sweeps"] == cr["sweeps"]
    assert np.max(np.abs(Hg - Hr) / np.maximum(1.0, np.abs(Hr))) <= 1e-6
    assert np.max(np.abs(Wg - Wr)) <= 1e-7
    for a, b in zip(bg, br):
        assert a == pytest.approx(b, rel=1e-6, abs=1e-9)
    assert tg == pytest.approx(tr, rel=1e-9)
    sink_total = sum(np.sum(np.abs(cat.rain_sink_source(m))) * 3600.0 for m in mm)
    assert abs(cg["last_mbe"] - cr["last_mbe"]) <= 1e-6 * sink_total
    print(f"[c2 window] {len(dg)} accepted steps, {cg['approximations']} approximations, {cg['sweeps']} sweeps on both sides; "
          f"max |dH| {np.max(np.abs(Hg - Hr)):.2e} m, max |dtheta| {np.max(np.abs(Wg - Wr)):.2e}")
