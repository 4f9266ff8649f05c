"""GPU: every scenario of tests/scenarios.py through the product's C ABI against the oracle
(the reference itself when oracle/_ref travelled with the snapshot), fp64 tolerances of
tests/test_gpu_parity.py."""
import pytest

from scenarios import SCENARIOS, compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_scenario(product, checker, name):
    a = SCENARIOS[name](product)
    b = SCENARIOS[name](checker)
    compare(a, b, exact=False)
