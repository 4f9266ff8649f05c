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
