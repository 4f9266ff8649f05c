"""TEST INFRASTRUCTURE: where the two CPU checkers live.  Only tests/, __graft_entry__.smoke() and bench.py's
CPU legs (cpu_baseline, --impl reference) import this; nothing under criteria3d_b200/ does
(tests/test_abi.py::test_product_does_not_link_or_reference_the_oracle)."""
from pathlib import Path

_HERE = Path(__file__).resolve().parent
ORACLE_LIB = _HERE / "libsf3d_oracle.so"                 # C restatement (oracle/sf3d_oracle.c), `make -C oracle port`
REFERENCE_LIB = _HERE / "_ref" / "libsf3d_ref.so"        # the unmodified reference behind the ABI, `make -C oracle ref`
GIS_REFERENCE_LIB = _HERE / "_ref" / "libgis_ref.so"


def checker_path() -> Path:
    """the reference itself when its prebuilt library travelled with the snapshot, else the restatement"""
    return REFERENCE_LIB if REFERENCE_LIB.exists() else ORACLE_LIB
