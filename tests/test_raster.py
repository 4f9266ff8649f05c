"""ESRI float grid I/O of the harness (the format of the reference's DEMs, soil maps and saved states)."""
import numpy as np

from criteria3d_b200.raster import EsriGrid, layer_to_grid, read_flt, write_flt
from criteria3d_b200.synth import Catchment


def test_flt_roundtrip(tmp_path):
    vals = np.arange(12, dtype=np.float32).reshape(3, 4)
    vals[1, 2] = -9999
    g = EsriGrid(vals, 641947.15, 5724524.79, 2.0, -9999.0)
    write_flt(tmp_path / "dem.flt", g)
    r = read_flt(tmp_path / "dem.flt")
    assert np.array_equal(r.values, vals) and r.cell == 2.0 and r.nodata == -9999.0
    assert r.valid.sum() == 11 and abs(r.xll - g.xll) < 1e-9


def test_layer_to_grid_places_values_by_cell_rank():
    valid = np.array([[True, False, True], [True, True, False]])
    cat = Catchment(2, 3, 2, valid=valid)
    layer = np.array([10.0, 11.0, 12.0, 13.0])
    like = EsriGrid(np.zeros((2, 3), np.float32), 0, 0, 10.0, -9999.0)
    g = layer_to_grid(layer, cat.cell_rank, like)
    assert np.array_equal(g.values, np.array([[10, -9999, 11], [12, 13, -9999]], np.float32))
    assert cat.n_surface == 4 and cat.rain_sink_source(3.6).shape == (4,)
