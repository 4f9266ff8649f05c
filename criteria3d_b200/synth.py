"""Deterministic synthetic DEM catchments and the driver that feeds
them to any implementation of the sf3d C ABI in the order Project3D::initialize3DModel uses
(src/project3D/project3D.cpp:456-616).

Everything here is host-side input preparation (numpy); no computation of the time step.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from .capi import (BoundaryType, Field, GridDesc, HeatFluxSaveMode, MeanType, SoilFluxes3D,
                   WRCModel)

SEED = 20240601

# (alpha [m-1], n, he [m], theta_r, theta_s, Ksat [m s-1], L, organic matter, clay) per horizon.
# Four soils x three horizons; values are typical van Genuchten sets (sand, loam, silt loam,
# clay), Ksat decreasing with depth.  The generator is the specification: the same table is
# fed to the reference, to the CPU restatement and to the product.
SOIL_TABLE = [
    # sand
    [(14.5, 2.68, 0.010, 0.045, 0.43, 8.25e-5, 0.5, 0.010, 0.03),
     (14.5, 2.68, 0.010, 0.045, 0.41, 6.00e-5, 0.5, 0.005, 0.03),
     (12.4, 2.28, 0.012, 0.057, 0.41, 4.05e-5, 0.5, 0.003, 0.05)],
    # loam
    [(3.6, 1.56, 0.020, 0.078, 0.43, 2.89e-6, 0.5, 0.020, 0.20),
     (3.6, 1.56, 0.020, 0.078, 0.41, 2.00e-6, 0.5, 0.010, 0.22),
     (2.7, 1.45, 0.025, 0.080, 0.40, 1.20e-6, 0.5, 0.005, 0.25)],
    # silt loam
    [(2.0, 1.41, 0.030, 0.067, 0.45, 1.25e-6, 0.5, 0.025, 0.15),
     (2.0, 1.41, 0.030, 0.067, 0.43, 9.00e-7, 0.5, 0.012, 0.17),
     (1.6, 1.37, 0.035, 0.070, 0.42, 6.00e-7, 0.5, 0.006, 0.20)],
    # clay
    [(0.8, 1.09, 0.050, 0.068, 0.38, 5.56e-7, 0.5, 0.030, 0.50),
     (0.8, 1.09, 0.050, 0.068, 0.37, 3.50e-7, 0.5, 0.015, 0.52),
     (0.6, 1.08, 0.060, 0.070, 0.36, 2.00e-7, 0.5, 0.008, 0.55)],
]
SURFACE_TABLE = [(0.05, 0.002), (0.24, 0.01)]      # (Manning roughness [s m-1/3], pond [m])
STORM_MM_H = [5.0, 20.0, 40.0, 25.0, 10.0, 2.0]     # C2 6 h hyetograph


def _hash01(seed: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Counter-based hash of (seed, a, b) -> U[0,1) (portable; no libc rand)."""
    x = (a.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
         + b.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
         + np.uint64((seed * 0x165667B19E3779F9) & 0xFFFFFFFFFFFFFFFF))
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _raster_slope_and_outlet(dem: np.ndarray, valid: np.ndarray, cell: float):
    """Slope and runoff boundary of a real raster exactly as the reference's caller prepares them:
    gis::computeSlopeAspectMaps, Project3D::setLateralBoundary (gis::isBoundaryRunoff) and the
    tan(slope) of Project3D::setCrit3DTopography, restated in criteria3d_b200/raster.py and pinned bit for bit
    against the reference's own gis code (tests/test_raster.py)."""
    from .raster import boundary_runoff, boundary_slope_tan, prepare_on_device, slope_aspect
    nodata = -9999.0
    z = np.where(valid, dem, np.float32(nodata)).astype(np.float32)
    import torch
    if torch.cuda.is_available():
        # on a GPU box the maps come from the product's device kernel (include/sf3d_gis.h), bit-identical to the host
        # restatement on every test raster (tests/test_gpu_raster_prep.py)
        _, _, outlet, tan_all = prepare_on_device(z, cell, nodata)
        tan = np.where(valid, tan_all, np.float32(0.0)).astype(np.float32)
        return np.ascontiguousarray(tan), np.ascontiguousarray(outlet)
    slope_deg, aspect = slope_aspect(z, cell, nodata)
    tan = np.where(valid, boundary_slope_tan(slope_deg), np.float32(0.0)).astype(np.float32)
    return np.ascontiguousarray(tan), np.ascontiguousarray(boundary_runoff(z, aspect, nodata))


def soil_layers(n_soil_layers: int, min_t=0.02, max_t=0.10, max_t_depth=0.40):
    """Layer thickness/centre-depth progression of Project3D::setSoilLayers/setLayersDepth
    (project3D.cpp:1568-1661), truncated/extended to exactly `n_soil_layers` soil layers.
    Returns (depth[L+1], thickness[L+1]) with layer 0 = surface (0, 0)."""
    if min_t == max_t:
        best = 1.0
    else:
        factor, best, best_err = 1.01, 1.01, 99.0
        while factor <= 2.0:
            upper, t = 0.0, min_t
            depth = upper + t * 0.5
            while t < max_t:
                upper += t
                t = min(t * factor, max_t)
                depth = upper + t * 0.5
            err = abs(depth - max_t_depth)
            if err < best_err:
                best_err, best = err, factor
            factor += 0.01
    thick = [0.0, min_t]
    depth = [0.0, min_t * 0.5]
    cur = min_t
    for i in range(2, n_soil_layers + 1):
        t = min(max_t, thick[i - 1] * best)
        thick.append(t)
        depth.append(cur + t * 0.5)
        cur += t
    return np.array(depth, np.float64), np.array(thick, np.float64)


@dataclass
class Catchment:
    rows: int
    cols: int
    n_soil_layers: int
    cell: float = 10.0
    heat: bool = False
    saturated_bottom: bool = False       # C4: lower third of the layers start at psi = +0.1 m
    seed: int = SEED
    initial_psi: float = -2.0
    row0: int = 0                        # first global DEM row of this (slab of the) raster
    global_rows: int | None = None       # rows of the whole catchment (None: this raster is the whole)
    valid: np.ndarray | None = field(default=None, repr=False)      # bool rows x cols, False = NODATA cell
    dem_override: np.ndarray | None = field(default=None, repr=False)   # use this DEM instead of the generator
    soil_override: np.ndarray | None = field(default=None, repr=False)  # soil id per cell
    # filled by __post_init__
    dem: np.ndarray = field(init=False, repr=False)
    slope_tan: np.ndarray = field(init=False, repr=False)
    cell_rank: np.ndarray = field(init=False, repr=False)
    outlet: np.ndarray = field(init=False, repr=False)
    soil_id: np.ndarray = field(init=False, repr=False)
    surface_id: np.ndarray = field(init=False, repr=False)
    pond: np.ndarray = field(init=False, repr=False)
    layer_depth: np.ndarray = field(init=False, repr=False)
    layer_thickness: np.ndarray = field(init=False, repr=False)
    layer_horizon: np.ndarray = field(init=False, repr=False)

    def __post_init__(self):
        R, Cc, cell = self.rows, self.cols, self.cell
        RG = self.global_rows if self.global_rows is not None else R      # generator works in global rows
        r = (self.row0 + np.arange(R, dtype=np.float64))[:, None]
        c = np.arange(Cc, dtype=np.float64)[None, :]
        ri = (self.row0 + np.arange(R, dtype=np.int64))[:, None] + np.zeros((1, Cc), np.int64)
        ci = np.arange(Cc, dtype=np.int64)[None, :] + np.zeros((R, 1), np.int64)
        u = _hash01(self.seed, ri, ci)
        z = (200.0 + cell * (0.03 * (RG - 1 - r) + 0.01 * c)
             + 5.0 * np.sin(2 * np.pi * r / 257.0) * np.cos(2 * np.pi * c / 193.0) + 0.25 * u)
        self.dem = np.ascontiguousarray(z, dtype=np.float32)
        if self.dem_override is not None:
            self.dem = np.ascontiguousarray(self.dem_override, dtype=np.float32)
            assert self.dem.shape == (R, Cc)
        # analytic slope of the smooth part of the DEM (independent of how the raster is cut into slabs)
        gy = -0.03 + 5.0 * (2 * np.pi / 257.0 / cell) * np.cos(2 * np.pi * r / 257.0) * np.cos(2 * np.pi * c / 193.0)
        gx = 0.01 - 5.0 * (2 * np.pi / 193.0 / cell) * np.sin(2 * np.pi * r / 257.0) * np.sin(2 * np.pi * c / 193.0)
        self.slope_tan = np.ascontiguousarray(np.sqrt(gx * gx + gy * gy), dtype=np.float32)
        self.outlet = np.zeros((R, Cc), np.uint8)
        if self.valid is None:
            self.cell_rank = np.arange(R * Cc, dtype=np.int32).reshape(R, Cc)
            if self.row0 + R == RG:
                self.outlet[R - 1, :] = 1               # outlet edge = last row of the whole catchment
        else:
            valid = np.ascontiguousarray(self.valid, dtype=bool)
            assert valid.shape == (R, Cc)
            rank = np.full((R, Cc), -1, np.int32)
            rank[valid] = np.arange(int(valid.sum()), dtype=np.int32)
            self.cell_rank = np.ascontiguousarray(rank)
            if self.dem_override is not None:
                self.slope_tan, self.outlet = _raster_slope_and_outlet(self.dem, valid, cell)
            elif self.row0 + R == RG:
                self.outlet[R - 1, :] = valid[R - 1, :]
        self.soil_id = (np.floor(_hash01(self.seed + 1, ri // 64, ci // 64) * 4).astype(np.uint16) % 4)
        if self.soil_override is not None:
            self.soil_id = np.ascontiguousarray(self.soil_override, dtype=np.uint16)
        rough = _hash01(self.seed + 2, ri, ci) < 0.10
        self.surface_id = np.ascontiguousarray(rough.astype(np.uint16))
        self.pond = np.where(rough, SURFACE_TABLE[1][1], SURFACE_TABLE[0][1]).astype(np.float64)
        self.layer_depth, self.layer_thickness = soil_layers(self.n_soil_layers)
        self.layer_horizon = np.where(self.layer_depth < 0.30, 0,
                                      np.where(self.layer_depth < 0.70, 1, 2)).astype(np.uint16)

    # ---- sizes -----------------------------------------------------------------------
    @property
    def layers(self) -> int:
        return self.n_soil_layers + 1

    @property
    def n_surface(self) -> int:
        return self.rows * self.cols if self.valid is None else int(np.count_nonzero(self.valid))

    @property
    def n_nodes(self) -> int:
        return self.layers * self.n_surface

    def min_delta_t(self) -> float:
        # Project3D::setAccuracy at accuracy 3 (project3D.cpp:619-635): vMax = 20 m/s
        return min(6.0, self.cell / 20.0)

    def grid_desc(self) -> GridDesc:
        d = GridDesc()
        d.rows, d.cols, d.layers, d.n_valid = self.rows, self.cols, self.layers, self.n_surface
        RG = self.global_rows if self.global_rows is not None else self.rows
        d.cell, d.x_ll, d.y_ll = self.cell, 0.0, self.cell * (RG - self.row0 - self.rows)
        d.dem = self.dem.ctypes.data_as(C.POINTER(C.c_float))
        d.slope_tan = self.slope_tan.ctypes.data_as(C.POINTER(C.c_float))
        d.cell_rank = self.cell_rank.ctypes.data_as(C.POINTER(C.c_int32))
        d.outlet = self.outlet.ctypes.data_as(C.POINTER(C.c_uint8))
        d.soil_id = self.soil_id.ctypes.data_as(C.POINTER(C.c_uint16))
        d.surface_id = self.surface_id.ctypes.data_as(C.POINTER(C.c_uint16))
        d.pond = self.pond.ctypes.data_as(C.POINTER(C.c_double))
        d.layer_depth = self.layer_depth.ctypes.data_as(C.POINTER(C.c_double))
        d.layer_thickness = self.layer_thickness.ctypes.data_as(C.POINTER(C.c_double))
        d.layer_horizon = self.layer_horizon.ctypes.data_as(C.POINTER(C.c_uint16))
        d.boundary_l1 = None
        d.free_catchment_runoff = d.free_lateral_drainage = d.free_bottom_drainage = 1
        d.heat_surface_layer1 = 1 if self.heat else 0
        return d

    # ---- forcing ---------------------------------------------------------------------
    def rain_sink_source(self, mm_per_hour: float) -> np.ndarray:
        """Surface sink/source [m3 s-1] as assignPrecipitation does
        (criteria3DProject.cpp:914-968): area * mm / 1000 / 3600, scaled in space by
        1 + 0.3 sin(2 pi c / C)."""
        c = np.arange(self.cols, dtype=np.float64)[None, :]
        scale = 1.0 + 0.3 * np.sin(2 * np.pi * c / self.cols) + np.zeros((self.rows, 1))
        area = self.cell * self.cell
        q = area * mm_per_hour * scale / 1000.0 / 3600.0
        return np.ascontiguousarray(q.reshape(-1) if self.valid is None else q[np.asarray(self.valid, bool)])

    def rain_raster(self, mm_per_hour: float, nodata: float = -9999.0) -> np.ndarray:
        """The hourly precipitation map [mm h-1] (float32, as the reference's meteo maps) whose
        assignPrecipitation result is rain_sink_source up to the float rounding of the map."""
        c = np.arange(self.cols, dtype=np.float64)[None, :]
        mm = (mm_per_hour * (1.0 + 0.3 * np.sin(2 * np.pi * c / self.cols)) + np.zeros((self.rows, 1))).astype(np.float32)
        if self.valid is not None:
            mm[~np.asarray(self.valid, bool)] = nodata
        return np.ascontiguousarray(mm)

    def initial_matric_potential(self) -> np.ndarray:
        psi = np.full(self.n_nodes, self.initial_psi, np.float64)
        psi[: self.n_surface] = 0.0
        if self.saturated_bottom:
            first_sat = self.layers - max(1, self.n_soil_layers // 3)
            psi[first_sat * self.n_surface:] = 0.1
        return psi


def setup(sf: SoilFluxes3D, cat: Catchment, threads: int = 0,
          numerics: tuple | None = None, heat_flux_mode: int | None = None, balance: bool = True) -> None:
    """initialize3DModel's call sequence (project3D.cpp:456-616) on implementation `sf`."""
    hf = HeatFluxSaveMode.Total if cat.heat else HeatFluxSaveMode.None_
    if heat_flux_mode is not None:
        hf = HeatFluxSaveMode(heat_flux_mode)
    sf.reset_solver()       # same starting deltaTcurr as a fresh process (see sf3d_ext_reset_solver)
    _ok(sf.initializeSF3D(cat.n_nodes, cat.n_surface, 8, True, cat.heat, False, int(hf)), "initializeSF3D")
    for i, (rough, _pond) in enumerate(SURFACE_TABLE):
        _ok(sf.setSurfaceProperties(i, rough), "setSurfaceProperties")
    for s, horizons in enumerate(SOIL_TABLE):
        for h, (a, n, he, tr, ts, ks, l, om, clay) in enumerate(horizons):
            _ok(sf.setSoilProperties(s, h, a, n, 1.0 - 1.0 / n, he, tr, ts, ks, l, om, clay), "setSoilProperties")
    desc = cat.grid_desc()
    _ok(sf.build_grid(desc), "sf3d_ext_build_grid")
    _ok(sf.setHydraulicProperties(int(WRCModel.ModifiedVanGenuchten), int(MeanType.Logarithmic), 10.0),
        "setHydraulicProperties")
    if numerics is None:
        numerics = (cat.min_delta_t(), 3600.0, 150, 10, 10, 3)
    _ok(sf.setNumericalParameters(*numerics), "setNumericalParameters")
    sf.setThreadsNumber(threads)
    _ok(sf.set_field(Field.MATRIC_POTENTIAL, 0, cat.initial_matric_potential()), "initial matric potential")
    if cat.heat:
        setup_heat(sf, cat, mode=int(hf))
    if balance:         # row slabs initialise the balance after the halo is wired (reductions over ranks)
        _ok(sf.initializeBalance(), "initializeBalance")


def setup_heat(sf: SoilFluxes3D, cat: Catchment, hour: int = 0, advection: bool = False, latent: bool = True,
               mode: int = int(HeatFluxSaveMode.Total)) -> None:
    """C3 heat configuration: T0 = 288.15 K; HeatSurface boundary on the first soil layer
    (set by the grid builder) with 2 m measurement heights and 0.01 m roughness; fixed 285.15 K at
    0.3 m below the free-drainage bottom nodes; total heat flux saved; latent heat on.
    Advection defaults to OFF: with it on, the reference's temperatures run away (SURVEY Appendix B Q1:
    its advective term is not conservative) and reach NaN within one 600 s water step on these catchments;
    the advective term is pinned on short bounded steps instead (tests/scenarios.py heat_advective_*)."""
    ns, n = cat.n_surface, cat.n_nodes
    _ok(sf.initializeHeatFlag(int(mode), advection, latent), "initializeHeatFlag")
    _ok(sf.set_field(Field.TEMPERATURE, ns, np.full(n - ns, 288.15)), "setNodeTemperature")
    for f, val in ((Field.BOUNDARY_HEIGHT_WIND, 2.0), (Field.BOUNDARY_HEIGHT_TEMPERATURE, 2.0), (Field.BOUNDARY_ROUGHNESS, 0.01)):
        _ok(sf.set_field(f, ns, np.full(ns, val)), f.name)
    _ok(sf.set_fixed_temperature(n - ns, np.full(ns, 285.15), 0.3), "setNodeBoundaryFixedTemperature")
    set_heat_forcing(sf, cat, hour)


def set_heat_forcing(sf: SoilFluxes3D, cat: Catchment, hour: int) -> None:
    """hourly atmospheric forcing on the HeatSurface nodes: air T 293.15 + 5 sin(2 pi h / 24) K, RH 60 %,
    wind 2 m/s, net irradiance 400 sin(pi h / 12) W/m2 by day (0 at night)"""
    ns = cat.n_surface
    air_t = 293.15 + 5.0 * math.sin(2 * math.pi * hour / 24.0)
    irr = max(0.0, 400.0 * math.sin(math.pi * (hour % 24) / 12.0))
    for f, val in ((Field.BOUNDARY_TEMPERATURE, air_t), (Field.BOUNDARY_RELATIVE_HUMIDITY, 60.0),
                   (Field.BOUNDARY_WIND_SPEED, 2.0), (Field.BOUNDARY_NET_IRRADIANCE, irr)):
        _ok(sf.set_field(f, ns, np.full(ns, val)), f.name)


def run_hours(sf: SoilFluxes3D, cat: Catchment, hours_mm: list[float], max_steps: int | None = None):
    """Per model hour: setSinkSource then the computeStep loop of runWaterFluxes3DModel
    (project3D.cpp:1307-1386).  Returns the accepted time-step sequence."""
    dts: list[float] = []
    sink = np.zeros(cat.n_nodes, np.float64)
    for mm in hours_mm:
        sink[: cat.n_surface] = cat.rain_sink_source(mm)
        _ok(sf.set_field(Field.WATER_SINK_SOURCE, 0, sink), "setNodeWaterSinkSource")
        t = 0.0
        while t < 3600.0:
            dt = sf.computeStep(3600.0 - t)
            dts.append(dt)
            t += dt
            if max_steps is not None and len(dts) >= max_steps:
                return dts
    return dts


def _ok(rc: int, what: str) -> None:
    if rc:
        raise RuntimeError(f"{what} failed with SF3Derror {rc}")
