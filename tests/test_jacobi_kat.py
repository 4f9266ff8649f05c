"""Known-answer test of the hot kernel: one Jacobi sweep on linear systems captured from INSIDE the
reference (tests/golden/jacobi_kat.npz: the compact-row matrix, b, x that Water::JacobiWaterCPU saw, and
the x it produced).  The row arithmetic has no libm call, so x must be bit-identical everywhere
(the product is compiled with -fmad=false); the norm differs only by summation order."""
from pathlib import Path

import numpy as np
import pytest

from criteria3d_b200 import SoilFluxes3D
from oracle import ORACLE_LIB, REFERENCE_LIB

KAT = Path(__file__).parent / "golden" / "jacobi_kat.npz"


def _cases():
    with np.load(KAT) as z:
        d = {k: z[k] for k in z.files}
    for k in (0, 1):
        yield (int(d[f"ns{k}"]), d[f"ncols{k}"], d[f"col{k}"], d[f"val{k}"], d[f"b{k}"], d[f"z{k}"], d[f"x_in{k}"],
               d[f"x_out{k}"], float(d[f"norm{k}"]))


def _check(sf, norm_rel):
    for ns, ncols, col, val, b, z, x_in, x_ref, norm_ref in _cases():
        x, norm = sf.jacobi_sweep(ns, ncols, col, val, b, z, x_in)
        assert np.array_equal(x, x_ref), f"max diff {np.max(np.abs(x - x_ref))}"
        assert abs(norm - norm_ref) <= norm_rel * abs(norm_ref)
        assert (ncols < 11).any() and (ncols == 11).any()        # ragged rows and full rows are both present


def test_restatement_sweep_is_bit_exact():
    if not ORACLE_LIB.exists():
        pytest.skip("oracle library not built")
    _check(SoilFluxes3D(ORACLE_LIB), 0.0)


def test_reference_reproduces_its_own_capture():
    if not REFERENCE_LIB.exists():
        pytest.skip("oracle/_ref not built here")
    from criteria3d_b200.synth import Catchment, setup
    ref = SoilFluxes3D(REFERENCE_LIB)
    setup(ref, Catchment(4, 4, 2), threads=1)                   # the real sweep needs an initialised solver object
    _check(ref, 0.0)


@pytest.mark.gpu
def test_product_sweep_is_bit_exact_on_x(product):
    _check(product, 1e-12)
