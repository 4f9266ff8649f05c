import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def checker():
    """The oracle used as checker: the unmodified reference behind the C ABI when its prebuilt
    library is present (oracle/_ref/, built here from /root/reference and shipped with the
    snapshot), else the CPU restatement (oracle/libsf3d_oracle.so)."""
    from criteria3d_b200 import SoilFluxes3D
    from oracle import ORACLE_LIB, REFERENCE_LIB
    if REFERENCE_LIB.exists():
        return SoilFluxes3D(REFERENCE_LIB)
    if ORACLE_LIB.exists():
        return SoilFluxes3D(ORACLE_LIB)
    pytest.skip("no oracle library built")


@pytest.fixture(scope="session")
def product():
    """The CUDA product through its C ABI.  No fallback: missing library or device = failure."""
    from criteria3d_b200 import load_product
    return load_product()
