"""GPU: every scenario of tests/scenarios.py through the product's C ABI against the oracle
(the reference itself when oracle/_ref travelled with the snapshot), fp64 tolerances of
tests/test_gpu_parity.py."""
import pytest

from scenarios import HEAT_SCENARIOS, SCENARIOS, TOLERANCES, compare

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
    compare(a, b, exact=False, **TOLERANCES.get(name, {}))


def test_heat_coupled_medium(product, checker):
    """32 x 32 x 11 coupled run (heat parity past toy sizes); also reports how far the Jacobi heat
    solve stays from its 4x sweep cap compared with the reference's Gauss-Seidel (Q6)"""
    from scenarios import heat_coupled_medium
    a = heat_coupled_medium(product)
    b = heat_coupled_medium(checker)
    compare(a, b, exact=False)
    print(f"[heat_coupled_medium] heat sub-steps {int(a['heat_counters'][0])}, heat sweeps product (Jacobi) "
          f"{int(a['heat_counters'][1])} vs reference (Gauss-Seidel) {int(b['heat_counters'][1])}, cap hits "
          f"{int(a['heat_cap_hits'])} / {int(b['heat_cap_hits'])}")


def test_culvert_against_restatement(product):
    """culvert boundary: product vs the C restatement (the reference cannot run it)"""
    import numpy as np
    from criteria3d_b200 import SoilFluxes3D
    from oracle import ORACLE_LIB
    from scenarios import culvert_outlet
    if not ORACLE_LIB.exists():
        pytest.skip("oracle library not built")
    a = culvert_outlet(product)
    b = culvert_outlet(SoilFluxes3D(ORACLE_LIB))
    assert b["culvert_total"] < 0.0                      # water does leave through the culvert
    assert abs(a["culvert_total"] - b["culvert_total"]) <= 1e-6 * abs(b["culvert_total"])
    a.pop("culvert_total"); b.pop("culvert_total")
    compare(a, b, exact=False)
