"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsf3d_ref.so, built from
/root/reference by oracle/Makefile).  Run in the build container, where the reference exists:

    python tests/golden/make_golden.py

Each file holds the outputs of one scenario of tests/scenarios.py (1 OpenMP thread): potentials,
water contents, conductivities, flow sums, accepted time steps, counters, balance scalars.  The
inputs are the seeded generator itself (criteria3d_b200/synth.py), so the fixtures stay small."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from criteria3d_b200 import SoilFluxes3D  # noqa: E402
from oracle import REFERENCE_LIB  # noqa: E402
from scenarios import HEAT_SCENARIOS, SCENARIOS  # noqa: E402


def config1_inputs():
    """rasters of the bundled STH sample project -> tests/golden/config1_sth_inputs.npz"""
    from criteria3d_b200.raster import read_flt
    base = Path("/root/reference/DATA/PROJECT/STH")
    dem = read_flt(base / "MAPS" / "DEM_STH.flt")
    soil = read_flt(base / "SOIL" / "soilMap_STH.flt")
    np.savez_compressed(Path(__file__).parent / "config1_sth_inputs.npz", dem=dem.values, soil=soil.values,
                        cell=np.float64(dem.cell), xll=np.float64(dem.xll), yll=np.float64(dem.yll))


def jacobi_kat(ref):
    """linear systems captured from inside the reference (the matrix JacobiWaterCPU actually sees, compact
    rows with dropped zero conductances) -> tests/golden/jacobi_kat.npz"""
    import ctypes as C
    from criteria3d_b200.synth import Catchment, run_hours, setup
    lib = ref.lib
    lib.sf3d_ref_captured_rows.restype = C.c_uint32
    lib.sf3d_ref_captured_norm.restype = C.c_double
    cat = Catchment(14, 11, 4)
    setup(ref, cat, threads=1)
    lib.sf3d_ref_capture_jacobi(1)
    out = {}
    for k, mm in enumerate((25.0, 40.0)):
        run_hours(ref, cat, [mm], max_steps=6 + 5 * k)
        n = lib.sf3d_ref_captured_rows()
        ncols = np.empty(n, np.uint8); col = np.empty(n * 11, np.uint32); val = np.empty(n * 11, np.float64)
        b = np.empty(n); x_in = np.empty(n); x_out = np.empty(n)
        P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        lib.sf3d_ref_captured_copy(P(ncols, C.c_uint8), P(col, C.c_uint32), P(val, C.c_double), P(b, C.c_double),
                                   P(x_in, C.c_double), P(x_out, C.c_double))
        z = ref.get_field(4, 0, n) - ref.get_field(3, 0, n)          # H - psi
        out.update({f"ncols{k}": ncols, f"col{k}": col, f"val{k}": val, f"b{k}": b, f"z{k}": z, f"x_in{k}": x_in,
                    f"x_out{k}": x_out, f"norm{k}": np.float64(lib.sf3d_ref_captured_norm()), f"ns{k}": np.int64(cat.n_surface)})
    lib.sf3d_ref_capture_jacobi(0)
    np.savez_compressed(Path(__file__).parent / "jacobi_kat.npz", **out)
    print("jacobi_kat: 2 systems of", n, "rows")


def gis_reference(dem, cell, flag=-9999.0):
    """slope [deg], aspect [deg], runoff boundary mask and tan(slope) from the reference's own gis code
    (oracle/_ref/libgis_ref.so = agrolib/gis compiled where it lies + oracle/gis_ref_capi.cpp)"""
    import ctypes as C
    lib = C.CDLL(str(ROOT / "oracle" / "_ref" / "libgis_ref.so"))
    d = np.ascontiguousarray(dem, np.float32)
    rows, cols = d.shape
    slope, aspect, tan = np.empty_like(d), np.empty_like(d), np.empty_like(d)
    boundary = np.empty(d.shape, np.uint8)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib.gisref_slope_aspect_boundary(rows, cols, C.c_double(cell), C.c_float(flag), P(d, C.c_float), P(slope, C.c_float),
                                          P(aspect, C.c_float), P(boundary, C.c_uint8), P(tan, C.c_float))
    assert rc == 0
    return slope, aspect, boundary, tan


def gis_golden():
    """the reference's slope / aspect / runoff boundary of the bundled STH DEM -> tests/golden/gis_sth.npz"""
    with np.load(Path(__file__).parent / "config1_sth_inputs.npz") as z:
        dem, cell = z["dem"], float(z["cell"])
    slope, aspect, boundary, tan = gis_reference(dem, cell)
    np.savez_compressed(Path(__file__).parent / "gis_sth.npz", slope=slope, aspect=aspect, boundary=boundary, tan=tan)
    print("gis_sth:", int(boundary.sum()), "runoff boundary cells")


def main():
    config1_inputs()
    gis_golden()
    jacobi_kat(SoilFluxes3D(REFERENCE_LIB))
    ref = SoilFluxes3D(REFERENCE_LIB)
    assert ref.backend == "reference"
    for name, fn in sorted({**SCENARIOS, **HEAT_SCENARIOS}.items()):
        out = fn(ref)
        path = Path(__file__).parent / f"{name}.npz"
        np.savez_compressed(path, **out)
        print(f"{name}: {path.stat().st_size} bytes, {len(out['dts'])} steps")


if __name__ == "__main__":
    main()
