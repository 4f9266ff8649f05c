"""GPU: sf3d_ext_get_layer_rasters_async / sf3d_ext_wait_rasters (output maps copied on a second stream while the next
step runs) return exactly the maps of the synchronous call for the state at the time of the call."""
import numpy as np
import pytest
import torch

from criteria3d_b200 import Field
from criteria3d_b200.synth import Catchment, setup

pytestmark = pytest.mark.gpu


def test_async_maps_equal_sync_maps_while_the_next_steps_run(product):
    cat = Catchment(96, 80, 6)
    setup(product, cat)
    assert product.set_forcing_rasters(precipitation=cat.rain_raster(30.0)) == 0
    shape = (cat.rows, cat.cols)
    bufs = [torch.empty((cat.layers, *shape), dtype=torch.float32).pin_memory().numpy() for _ in range(3)]
    want = []
    for k in range(3):
        product.computeStep(3600.0)
        want.append(product.get_layer_rasters(Field.MATRIC_POTENTIAL, 0, cat.layers, shape).copy())
        product.get_layer_rasters_async(Field.MATRIC_POTENTIAL, 0, cat.layers, shape, bufs[k])     # three in flight: the third waits for the first
    product.computeStep(3600.0)                                 # runs while the copies land
    product.wait_rasters()
    for k in range(3):
        assert np.array_equal(bufs[k], want[k]), f"maps of step {k}"
    assert not np.array_equal(want[0], want[2])                 # the state did change between the calls
    # pageable destination: still correct
    out = np.empty((cat.layers, *shape), np.float32)
    product.get_layer_rasters_async(Field.WATER_CONTENT, 0, cat.layers, shape, out)
    product.wait_rasters()
    assert np.array_equal(out, product.get_layer_rasters(Field.WATER_CONTENT, 0, cat.layers, shape))

