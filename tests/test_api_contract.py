"""Return codes and sentinels of the plugin API (soilFluxes3D.cpp): the same call sequence on the
implementation under test and on the oracle must give identical codes.  CPU: restatement vs
reference; GPU: product vs oracle."""
import numpy as np
import pytest

from criteria3d_b200 import BoundaryType, Field, SoilFluxes3D
from oracle import ORACLE_LIB, REFERENCE_LIB
from criteria3d_b200.synth import Catchment, setup


def contract_codes(sf):
    out = []
    sf.cleanSF3D()
    # not initialised: MemoryError code / sentinel
    out += [sf.setNode(0, 0, 0, 0, 1, True, 0, 0, 0), sf.setNodeWaterSinkSource(0, 1.0), sf.getNodeWaterContent(0),
            sf.getNodeTemperature(0), sf.setNodeLink(0, 1, 1, 1.0)]
    cat = Catchment(6, 5, 3, heat=True)
    setup(sf, cat, threads=1)
    ns, n = cat.n_surface, cat.n_nodes
    hs, soil2, bottom = ns + 2, 2 * ns + 2, n - 1                  # HeatSurface node, plain soil node, FreeDrainage node
    out += [sf.setNodeBoundaryWindSpeed(hs, -1.0), sf.setNodeBoundaryWindSpeed(hs, 2000.0), sf.setNodeBoundaryWindSpeed(hs, 3.0),
            sf.setNodeBoundaryRoughness(hs, -0.1), sf.setNodeBoundaryRoughness(hs, 0.02),
            sf.setNodeBoundaryHeightWind(soil2, 2.0), sf.setNodeBoundaryTemperature(soil2, 290.0),       # NoBoundary -> BoundaryError
            sf.setNodeBoundaryNetIrradiance(n + 4, 1.0),                                                    # IndexError
            sf.setNodeBoundaryFixedTemperature(hs, 280.0, 0.3),                                             # wrong boundary type
            sf.setNodeBoundaryFixedTemperature(bottom, 280.0, 0.3),
            sf.setNodeTemperature(n, 280.0), sf.setNodeHeatSinkSource(soil2, 5.0),
            sf.setNodePrescribedTotalPotential(soil2, 1.0)]
    out += [sf.getNodeTemperature(0),                      # surface -> TopographyError sentinel
            sf.getNodeTemperature(n + 1), sf.getNodeHeatConductivity(1),
            sf.getNodeBoundarySensibleFlux(soil2), sf.getNodeBoundaryLatentFlux(bottom),      # not HeatSurface -> BoundaryError
            sf.getNodeBoundaryRadiativeFlux(hs), sf.getNodeBoundaryAdvectiveFlux(hs),
            sf.getNodeHeatMaxFlux(0, 1, 0), sf.getNodeHeatMaxFlux(soil2, 1, 3),              # Total mode: other types NODATA
            sf.getNodeHeatMaxFlux(soil2, 0, 0), sf.getNodeHeatMaxFlux(soil2, 3, 0),
            sf.getNodeBoundaryWaterFlow(soil2), sf.getNodeMaximumWaterContent(0), sf.getNodeMinimumWaterContent(soil2),
            sf.getNodePond(soil2), sf.getNodePond(1), sf.getNodeWaterDeficit(1, 3.0), sf.getHeatMBR(), sf.getWaterMBR()]
    # raster-facing extensions: shape mismatch, too many sink layers, layer out of range, unknown field
    import ctypes as C
    from criteria3d_b200.capi import ForcingDesc
    bad = np.zeros((cat.rows + 1, cat.cols), np.float32)
    out += [sf.set_forcing_rasters(precipitation=bad), sf.set_forcing_rasters(layer_sink=np.zeros((cat.layers + 1, cat.rows, cat.cols), np.float32)),
            sf.set_forcing_rasters(precipitation=np.zeros((cat.rows, cat.cols), np.float32)), sf.lib.sf3d_ext_set_forcing_rasters(None)]
    buf = np.zeros((cat.rows, cat.cols), np.float32)
    P = buf.ctypes.data_as(C.POINTER(C.c_float))
    out += [sf.lib.sf3d_ext_get_layer_rasters(int(Field.WATER_CONTENT), cat.layers, 1, -9999.0, P),
            sf.lib.sf3d_ext_get_layer_rasters(int(Field.WATER_CONTENT), cat.layers - 1, 2, -9999.0, P),
            sf.lib.sf3d_ext_get_layer_rasters(int(Field.WATER_CONTENT), 0, 0, -9999.0, P),
            sf.lib.sf3d_ext_get_layer_rasters(int(Field.WATER_SINK_SOURCE), 0, 1, -9999.0, P),
            sf.lib.sf3d_ext_get_layer_rasters(int(Field.WATER_CONTENT), 0, 1, -9999.0, None),
            sf.lib.sf3d_ext_get_layer_rasters(int(Field.WATER_CONTENT), 0, 1, -9999.0, P)]
    return np.array(out, dtype=np.float64)


@pytest.mark.skipif(not (ORACLE_LIB.exists() and REFERENCE_LIB.exists()), reason="needs both CPU libraries")
def test_restatement_matches_reference_codes():
    a = contract_codes(SoilFluxes3D(ORACLE_LIB))
    b = contract_codes(SoilFluxes3D(REFERENCE_LIB))
    assert np.array_equal(a, b, equal_nan=True), np.where(a != b)


@pytest.mark.gpu
def test_product_matches_oracle_codes(product, checker):
    a = contract_codes(product)
    b = contract_codes(checker)
    assert np.array_equal(a, b, equal_nan=True), (np.where(a != b), a[a != b], b[a != b])


@pytest.mark.parametrize("lib", [ORACLE_LIB, REFERENCE_LIB], ids=["oracle", "reference"])
def test_async_layer_rasters_on_the_cpu_libraries_are_the_synchronous_call(lib):
    """sf3d_ext_get_layer_rasters_async / sf3d_ext_wait_rasters exist in every implementation of the ABI; the CPU
    libraries have nothing to overlap and return the same maps at once"""
    if not lib.exists():
        pytest.skip(f"{lib.name} not built")
    sf = SoilFluxes3D(lib)
    cat = Catchment(7, 6, 3)
    setup(sf, cat, threads=1)
    shape = (cat.rows, cat.cols)
    want = sf.get_layer_rasters(Field.MATRIC_POTENTIAL, 0, cat.layers, shape)
    got = np.empty_like(want)
    sf.get_layer_rasters_async(Field.MATRIC_POTENTIAL, 0, cat.layers, shape, got)
    sf.wait_rasters()
    assert np.array_equal(got, want)
    with pytest.raises(RuntimeError):
        sf.get_layer_rasters_async(Field.MATRIC_POTENTIAL, cat.layers, 1, shape, got[:1])      # layer out of range
