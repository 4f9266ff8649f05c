"""Parity of the CUDA path (through the C ABI) with the oracle on identical inputs.

Tolerances (fp64; the only arithmetic differences are CUDA libm vs glibc and reduction order):
  lock-step single computeStep : |dH| <= 1e-9 * max(1, |psi|), |dtheta| <= 1e-10, same accepted dt,
                                 same approximation count
  trajectory (hours of storm)  : |dH| <= 1e-6 * max(1, |H|) (abs floor 1e-8 m), |dtheta| <= 1e-7,
                                 boundary flow sums rel 1e-6, mass-balance error abs <= 1e-6 * sum|sink|
  integer maps                 : bit-exact
"""
import numpy as np
import pytest

from criteria3d_b200 import BoundaryType, Field
from criteria3d_b200.synth import Catchment, run_hours, setup

pytestmark = pytest.mark.gpu


def _both(product, checker, cat, **kw):
    for sf in (product, checker):
        setup(sf, cat, **kw)


def test_index_maps_bit_exact(product, checker):
    cat = Catchment(37, 29, 5)
    _both(product, checker, cat)
    for slot in range(10):
        a, b = product.link_table(slot, 0, cat.n_nodes), checker.link_table(slot, 0, cat.n_nodes)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), f"slot {slot}"
        assert np.array_equal(a[2], b[2]), f"slot {slot} areas"
    for x, y in zip(product.node_meta(0, cat.n_nodes), checker.node_meta(0, cat.n_nodes)):
        assert np.array_equal(x, y)


def test_initial_state_and_storage(product, checker):
    cat = Catchment(32, 32, 6)
    _both(product, checker, cat)
    for f in (Field.TOTAL_POTENTIAL, Field.DEGREE_OF_SATURATION, Field.WATER_CONDUCTIVITY, Field.WATER_CONTENT):
        a, b = product.get_field(f, 0, cat.n_nodes), checker.get_field(f, 0, cat.n_nodes)
        assert np.allclose(a, b, rtol=1e-12, atol=1e-14), f.name
    assert product.getTotalWaterContent() == pytest.approx(checker.getTotalWaterContent(), rel=1e-12)


def test_lockstep_single_step(product, checker):
    cat = Catchment(40, 32, 6)
    _both(product, checker, cat)
    sink = np.zeros(cat.n_nodes)
    sink[: cat.n_surface] = cat.rain_sink_source(20.0)
    for sf in (product, checker):
        assert sf.set_field(Field.WATER_SINK_SOURCE, 0, sink) == 0
    dt_g, dt_o = product.computeStep(3600.0), checker.computeStep(3600.0)
    assert dt_g == dt_o
    cg, co = product.counters(), checker.counters()
    assert cg["approximations"] == co["approximations"]
    Hg, Ho = (s.get_field(Field.TOTAL_POTENTIAL, 0, cat.n_nodes) for s in (product, checker))
    psi = checker.get_field(Field.MATRIC_POTENTIAL, 0, cat.n_nodes)
    assert np.max(np.abs(Hg - Ho) / np.maximum(1.0, np.abs(psi))) <= 1e-9
    Wg, Wo = (s.get_field(Field.WATER_CONTENT, 0, cat.n_nodes) for s in (product, checker))
    assert np.max(np.abs(Wg - Wo)[cat.n_surface:]) <= 1e-10


@pytest.mark.parametrize("shape,hours", [((40, 32, 6), [20.0, 40.0, 5.0]), ((64, 48, 10), [40.0, 25.0])])
def test_trajectory(product, checker, shape, hours):
    cat = Catchment(*shape)
    _both(product, checker, cat)
    dg, do = run_hours(product, cat, hours), run_hours(checker, cat, hours)
    assert dg == do, "accepted time-step sequences differ"
    Hg, Ho = (s.get_field(Field.TOTAL_POTENTIAL, 0, cat.n_nodes) for s in (product, checker))
    assert np.max(np.abs(Hg - Ho) / np.maximum(1.0, np.abs(Ho))) <= 1e-6
    Wg, Wo = (s.get_field(Field.WATER_CONTENT, 0, cat.n_nodes) for s in (product, checker))
    assert np.max(np.abs(Wg - Wo)) <= 1e-7
    for bt in (BoundaryType.Runoff, BoundaryType.FreeDrainage, BoundaryType.FreeLateralDrainage):
        a, b = product.getTotalBoundaryWaterFlow(int(bt)), checker.getTotalBoundaryWaterFlow(int(bt))
        assert a == pytest.approx(b, rel=1e-6, abs=1e-9), bt.name
    assert product.getTotalWaterContent() == pytest.approx(checker.getTotalWaterContent(), rel=1e-9)
    cg, co = product.counters(), checker.counters()
    assert cg["steps"] == co["steps"] and cg["approximations"] == co["approximations"]
    sink_total = sum(np.sum(np.abs(cat.rain_sink_source(mm))) * 3600.0 for mm in hours)
    assert abs(cg["last_mbe"] - co["last_mbe"]) <= 1e-6 * sink_total


def test_pattern_compressed_indices_are_bit_identical(product, checker, monkeypatch):
    """The sweep reads the column indices through 16-bit link patterns (2 B/node instead of 40 B).
    The map is verified on the device at finalize; forcing the explicit index array must give
    bit-identical potentials (integer work: exact)."""
    cat = Catchment(33, 27, 5)
    runs = []
    for explicit in (False, True):
        if explicit:
            monkeypatch.setenv("SF3D_EXPLICIT_INDEX", "1")
        else:
            monkeypatch.delenv("SF3D_EXPLICIT_INDEX", raising=False)
        setup(product, cat)
        dts = run_hours(product, cat, [30.0], max_steps=25)
        runs.append((dts, product.get_field(Field.TOTAL_POTENTIAL, 0, cat.n_nodes), product.counters()["sweeps"]))
    assert runs[0][0] == runs[1][0] and runs[0][2] == runs[1][2]
    assert np.array_equal(runs[0][1], runs[1][1])
