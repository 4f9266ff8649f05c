"""Golden vectors produced by the unmodified reference (tests/golden/make_golden.py).
CPU: the oracle restatement reproduces them bit for bit (this pins the oracle on machines where
/root/reference is absent).  GPU: the product matches them within the stated fp64 tolerances."""
from pathlib import Path

import numpy as np
import pytest

from criteria3d_b200 import SoilFluxes3D
from oracle import ORACLE_LIB
from scenarios import HEAT_SCENARIOS, SCENARIOS, TOLERANCES, compare

GOLDEN = Path(__file__).parent / "golden"


def _load(name):
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", sorted({**SCENARIOS, **HEAT_SCENARIOS}))
def test_oracle_matches_reference_golden(name):
    if not ORACLE_LIB.exists():
        pytest.skip("oracle/libsf3d_oracle.so not built")
    port = SoilFluxes3D(ORACLE_LIB)
    compare({**SCENARIOS, **HEAT_SCENARIOS}[name](port), _load(name), exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted({**SCENARIOS, **HEAT_SCENARIOS}))
def test_product_matches_reference_golden(product, name):
    fn = {**SCENARIOS, **HEAT_SCENARIOS}[name]
    compare(fn(product), _load(name), exact=False, **TOLERANCES.get(name, {}))
