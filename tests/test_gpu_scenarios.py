"""GPU: every scenario of tests/scenarios.py through the product's C ABI against the oracle
(the reference itself when oracle/_ref travelled with the snapshot), fp64 tolerances of
tests/test_gpu_parity.py."""
import pytest

from scenarios import HEAT_SCENARIOS, SCENARIOS, compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_scenario(product, checker, name):
    a = SCENARIOS[name](product)
    b = SCENARIOS[name](checker)
    compare(a, b, exact=False)


@pytest.mark.parametrize("name", sorted(HEAT_SCENARIOS))
def test_heat_scenario(product, checker, name):
    """coupled heat (Jacobi on the GPU vs the reference's Gauss-Seidel: same fixed point)"""
    a = HEAT_SCENARIOS[name](product)
    b = HEAT_SCENARIOS[name](checker)
    compare(a, b, exact=False)
