"""ESRI float grid (.flt + .hdr) reader/writer: the raster format CRITERIA-3D uses for DEMs, soil maps
and saved states (agrolib/gis/gisIO.cpp:1587 readEsriGridFlt; criteria3DProject.cpp:2275-2301 writes
WP_<depth>.flt).  Harness-side I/O for the callers on either side of the time step."""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np


@dataclass
class EsriGrid:
    values: np.ndarray          # rows x cols float32, row 0 = northernmost
    xll: float
    yll: float
    cell: float
    nodata: float

    @property
    def valid(self) -> np.ndarray:
        return self.values != np.float32(self.nodata)


def read_flt(path: str | Path) -> EsriGrid:
    path = Path(path)
    hdr = {}
    for line in path.with_suffix(".hdr").read_text().splitlines():
        parts = line.split()
        if len(parts) >= 2:
            hdr[parts[0].lower()] = parts[1]
    rows, cols = int(hdr["nrows"]), int(hdr["ncols"])
    order = "<" if hdr.get("byteorder", "LSBFIRST").upper().startswith("LSB") else ">"
    data = np.fromfile(path.with_suffix(".flt"), dtype=order + "f4")
    if data.size != rows * cols:
        raise ValueError(f"{path}: {data.size} values, header says {rows}x{cols}")
    return EsriGrid(values=np.ascontiguousarray(data.reshape(rows, cols).astype(np.float32)),
                    xll=float(hdr["xllcorner"]), yll=float(hdr["yllcorner"]), cell=float(hdr["cellsize"]),
                    nodata=float(hdr.get("nodata_value", -9999)))


def write_flt(path: str | Path, grid: EsriGrid) -> None:
    path = Path(path)
    rows, cols = grid.values.shape
    path.with_suffix(".hdr").write_text(
        f"ncols         {cols}\nnrows         {rows}\nxllcorner     {grid.xll}\nyllcorner     {grid.yll}\n"
        f"cellsize      {grid.cell}\nNODATA_value  {grid.nodata}\nbyteorder     LSBFIRST\n")
    grid.values.astype("<f4").tofile(path.with_suffix(".flt"))


def layer_to_grid(values: np.ndarray, cell_rank: np.ndarray, like: EsriGrid) -> EsriGrid:
    """per-valid-cell values of one layer (bulk getter output) -> raster, as computeCriteria3DMap does"""
    out = np.full(cell_rank.shape, np.float32(like.nodata), np.float32)
    m = cell_rank >= 0
    out[m] = values[cell_rank[m]].astype(np.float32)
    return EsriGrid(out, like.xll, like.yll, like.cell, like.nodata)


# ---- slope / aspect / runoff boundary of a DEM, as the reference's caller prepares them ----------
# Restatement of agrolib/gis/gis.cpp (Qt-free) in numpy, pinned cell by cell against the reference's own code
# (oracle/_ref/libgis_ref.so, tests/test_raster.py).  Input preparation on the host: O(cells), once per project.
EPSILON = 0.00001               # mathFunctions/commonConstants.h:252
RAD_TO_DEG = 57.295779513       # :256
DEG_TO_RAD = 0.01745329252      # :255


def _is_flag(a: np.ndarray, flag: float) -> np.ndarray:
    """isEqual(value, flag), basicMath.h:25-26: fabs(double(a) - double(b)) < EPSILON"""
    return np.abs(a.astype(np.float64) - np.float64(np.float32(flag))) < EPSILON


def _neighbourhood(z: np.ndarray, flag: float):
    """nb(di, dj) = value at (row + di, col + dj), flag outside the grid (Crit3DRasterGrid::getValueFromRowCol, gis.cpp:521-530)"""
    R, C = z.shape
    pad = np.full((R + 2, C + 2), np.float32(flag), np.float32)
    pad[1:-1, 1:-1] = z
    return lambda di, dj: pad[1 + di: 1 + di + R, 1 + dj: 1 + dj + C]


def is_boundary(dem: np.ndarray, nodata: float = -9999.0) -> np.ndarray:
    """gis::isBoundary (gis.cpp:1494-1510): valid cell with at least one NODATA / out-of-grid neighbour"""
    z = np.ascontiguousarray(dem, np.float32)
    nb = _neighbourhood(z, nodata)
    any_missing = np.zeros(z.shape, bool)
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            if di or dj:
                any_missing |= _is_flag(nb(di, dj), nodata)
    return ~_is_flag(z, nodata) & any_missing


def slope_aspect(dem: np.ndarray, cell: float, nodata: float = -9999.0) -> tuple[np.ndarray, np.ndarray]:
    """gis::computeSlopeAspectMaps (gis.cpp:1190-1268): Horn's 3x3 derivatives in the interior,
    computeSlopeAspectBoundary (:1114-1186) on cells with a missing neighbour.  Returns (slope [deg], aspect [deg,
    0 = north, clockwise]) as float32 maps, NODATA where the DEM is."""
    z = np.ascontiguousarray(dem, np.float32)
    nb = _neighbourhood(z, nodata)
    valid = ~_is_flag(z, nodata)
    rim = is_boundary(z, nodata)
    d = lambda di, dj: nb(di, dj).astype(np.float64)

    # interior: Horn (:1219-1254)
    z1, z2, z3, z4, z6, z7, z8, z9 = d(-1, -1), d(-1, 0), d(-1, 1), d(0, -1), d(0, 1), d(1, -1), d(1, 0), d(1, 1)
    dzdx = ((z3 + 2 * z6 + z9) - (z1 + 2 * z4 + z7)) / (8.0 * cell)
    dzdy = ((z7 + 2 * z8 + z9) - (z1 + 2 * z2 + z3)) / (8.0 * cell)
    flat = (np.abs(dzdx) < EPSILON) & (np.abs(dzdy) < EPSILON)
    slope_in = (np.arctan(np.sqrt(dzdx * dzdx + dzdy * dzdy)) * RAD_TO_DEG).astype(np.float32)
    aspect = 90.0 - np.arctan2(dzdy, -dzdx) * RAD_TO_DEG
    aspect_in = np.where(aspect < 0, aspect + 360.0, aspect).astype(np.float32)
    slope_in[flat] = 0.0
    aspect_in[flat] = 0.0

    # rim: one-sided sums over the valid neighbours (:1126-1166); (z - z1) and its product with i are float operations
    def one_sided(outer_axis_rows: bool):
        dz = np.zeros(z.shape, np.float64)
        dl = np.zeros(z.shape, np.float64)
        for a in (-1, 1):
            for b in (-1, 0, 1):
                di, dj = (a, b) if outer_axis_rows else (b, a)
                z1f = nb(di, dj)
                ok = ~_is_flag(z1f, nodata)
                term = (np.float32(a) * (z - z1f)).astype(np.float64)
                dz = dz + np.where(ok, term, 0.0)
                dl = dl + np.where(ok, cell, 0.0)
        return dz / np.maximum(dl, EPSILON)
    with np.errstate(invalid="ignore", over="ignore"):
        dz_dy = one_sided(True)
        dz_dx = one_sided(False)
        slope_rim = (np.arctan(np.sqrt(dz_dx * dz_dx + dz_dy * dz_dy)) * RAD_TO_DEG).astype(np.float32)
        aspect = 90.0 - np.arctan2(-dz_dy, dz_dx) * RAD_TO_DEG
        aspect_rim = np.where(aspect < 0, aspect + 360, aspect).astype(np.float32)

    slope = np.where(rim, slope_rim, slope_in).astype(np.float32)
    asp = np.where(rim, aspect_rim, aspect_in).astype(np.float32)
    slope[~valid] = np.float32(nodata)
    asp[~valid] = np.float32(nodata)
    return slope, asp


def boundary_runoff(dem: np.ndarray, aspect: np.ndarray, nodata: float = -9999.0) -> np.ndarray:
    """gis::isBoundaryRunoff (gis.cpp:1452-1488) for every cell, with the surface index map of
    Project3D::setIndexMaps (every valid DEM cell is a node): rim cells that are strict minima, or whose
    aspect points at a cell outside the catchment.  This is Project3D::setLateralBoundary (project3D.cpp:851-873)."""
    z = np.ascontiguousarray(dem, np.float32)
    R, C = z.shape
    nb = _neighbourhood(z, nodata)
    valid = ~_is_flag(z, nodata)
    rim = is_boundary(z, nodata)
    strict_min = np.ones(z.shape, bool)                       # isMinimum(dtm, true, ...) :1395-1427
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            if di or dj:
                adj = nb(di, dj)
                strict_min &= _is_flag(adj, nodata) | ~(z >= adj)
    a = np.ascontiguousarray(aspect, np.float32)
    has_aspect = ~_is_flag(a, nodata)
    r = np.where((a >= 135) & (a <= 225), 1, np.where((a <= 45) | (a >= 315), -1, 0))
    c = np.where((a >= 45) & (a <= 135), 1, np.where((a >= 225) & (a <= 315), -1, 0))
    vpad = np.zeros((R + 2, C + 2), bool)
    vpad[1:-1, 1:-1] = valid
    rr, cc = np.mgrid[0:R, 0:C]
    target_missing = ~vpad[rr + r + 1, cc + c + 1]
    return (rim & valid & (strict_min | (has_aspect & target_missing))).astype(np.uint8)


def boundary_slope_tan(slope_deg: np.ndarray) -> np.ndarray:
    """float boundarySlope = tan(slopeDegree * DEG_TO_RAD), Project3D::setCrit3DTopography (project3D.cpp:964-965)"""
    return np.tan(np.ascontiguousarray(slope_deg, np.float32).astype(np.float64) * DEG_TO_RAD).astype(np.float32)


# ---- the same preparation on the GPU (include/sf3d_gis.h, criteria3d_b200/csrc/sf3d_gis.cu) ------------------------
def prepare_on_device(dem: np.ndarray, cell: float, nodata: float = -9999.0):
    """slope [deg], aspect [deg], runoff-boundary mask and tan(slope) of a DEM in one call of the product's
    sf3d_gis_slope_aspect_boundary (one CUDA thread per cell).  Fails loudly when the CUDA library is missing or no
    device is present: there is no CPU fallback (slope_aspect / boundary_runoff above are the host restatement the
    tests pin against the reference's own gis code)."""
    import ctypes
    from .capi import PRODUCT_LIB
    if not PRODUCT_LIB.exists():
        raise RuntimeError(f"{PRODUCT_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first")
    lib = ctypes.CDLL(str(PRODUCT_LIB))
    fn = lib.sf3d_gis_slope_aspect_boundary
    fp, up = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8)
    fn.restype = ctypes.c_uint8
    fn.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_double, ctypes.c_float, fp, fp, fp, up, fp]
    z = np.ascontiguousarray(dem, np.float32)
    rows, cols = z.shape
    slope = np.empty_like(z); aspect = np.empty_like(z); tan = np.empty_like(z)
    mask = np.empty(z.shape, np.uint8)
    rc = fn(rows, cols, float(cell), np.float32(nodata), z.ctypes.data_as(fp), slope.ctypes.data_as(fp), aspect.ctypes.data_as(fp),
            mask.ctypes.data_as(up), tan.ctypes.data_as(fp))
    if rc != 0:
        raise RuntimeError(f"sf3d_gis_slope_aspect_boundary returned {rc} (2 = no CUDA device / allocation failed, 6 = bad raster)")
    return slope, aspect, mask, tan
