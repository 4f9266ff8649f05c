"""GPU: raster preparation on the device (include/sf3d_gis.h) against the reference's own gis code: the golden
maps of the bundled STH DEM (generated from agrolib/gis by tests/golden/make_golden.py), the reference library itself
where it travelled (oracle/_ref/libgis_ref.so), and the host restatement (criteria3d_b200/raster.py, pinned bit-identical
to the reference) on a 2048 x 2048 DEM with holes.  Bar: identical boundary mask, identical float maps."""
import sys
from pathlib import Path

import numpy as np
import pytest

from criteria3d_b200.raster import boundary_runoff, boundary_slope_tan, prepare_on_device, slope_aspect

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
GIS_REF = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "libgis_ref.so"
NODATA = np.float32(-9999)


def _same(dev, ref, dem, what):
    valid = dem != NODATA
    for a, b, name in zip(dev[:3], ref[:3], ("slope", "aspect", "boundary")):
        assert np.array_equal(a, b), f"{what}: {name}: {int((a != b).sum())} cells differ"
    assert np.array_equal(dev[3][valid], ref[3][valid]), f"{what}: tan(slope)"


def test_bundled_dem_matches_the_reference_golden():
    with np.load(GOLDEN / "config1_sth_inputs.npz") as z:
        dem, cell = z["dem"], float(z["cell"])
    with np.load(GOLDEN / "gis_sth.npz") as g:
        _same(prepare_on_device(dem, cell), (g["slope"], g["aspect"], g["boundary"], g["tan"]), dem, "STH")


def _cases():
    rng = np.random.default_rng(11)
    ragged = (200 + rng.random((61, 47)) * 30).astype(np.float32)
    ragged[rng.random(ragged.shape) < 0.15] = -9999
    bowl = (np.hypot(*np.mgrid[-8:9, -10:11]) * 0.7 + 50).astype(np.float32)
    bowl[0:3, 0:4] = -9999
    return {"ragged": (ragged, 10.0), "flat": (np.full((9, 9), 100, np.float32), 5.0), "bowl": (bowl, 2.0),
            "one row": ((100 + np.arange(12, dtype=np.float32))[None, :], 4.0), "one cell": (np.full((1, 1), 7, np.float32), 1.0),
            "all nodata": (np.full((4, 5), -9999, np.float32), 3.0)}


@pytest.mark.parametrize("name", sorted(_cases()))
def test_edge_rasters(name):
    dem, cell = _cases()[name]
    if GIS_REF.exists():
        sys.path.insert(0, str(GOLDEN))
        from make_golden import gis_reference
        ref = gis_reference(dem, cell)
    else:
        slope, aspect = slope_aspect(dem, cell)
        ref = (slope, aspect, boundary_runoff(dem, aspect), boundary_slope_tan(slope))
    _same(prepare_on_device(dem, cell), ref, dem, name)


def test_large_dem_with_holes():
    rng = np.random.default_rng(5)
    n = 2048
    y, x = np.mgrid[0:n, 0:n].astype(np.float32)
    dem = (300 + 0.05 * x + 0.03 * y + 8 * np.sin(x / 37) * np.cos(y / 53) + rng.random((n, n), dtype=np.float32)).astype(np.float32)
    dem[(x - 700) ** 2 + (y - 900) ** 2 < 150 ** 2] = -9999          # a lake
    dem[:, :5] = -9999
    slope, aspect = slope_aspect(dem, 10.0)
    ref = (slope, aspect, boundary_runoff(dem, aspect), boundary_slope_tan(slope))
    dev = prepare_on_device(dem, 10.0)
    _same(dev, ref, dem, "2048^2")
    assert int(dev[2].sum()) > 0


def test_bad_arguments():
    with pytest.raises(RuntimeError):
        prepare_on_device(np.zeros((0, 5), np.float32), 1.0)
