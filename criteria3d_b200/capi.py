"""ctypes binding of include/sf3d.h.

`SoilFluxes3D(path)` loads ONE implementation of the C ABI and exposes it under the
reference's own function names (agrolib/soilFluxes3D/soilFluxes3D.h:9-104), so harness and
test code reads like a caller of the reference plugin API:

    sf = SoilFluxes3D(PRODUCT_LIB)
    sf.initializeSF3D(nrNodes, nrSurface, 8, True, False, False)
    sf.setNode(i, x, y, z, area, True, BoundaryType.Runoff, slope, width)
    dt = sf.computeStep(3600.0)

Three libraries implement the ABI (see include/sf3d.h): the CUDA product, the CPU
restatement and the unmodified reference; the paths of the two checkers live in oracle/__init__.py
(test infrastructure), not here.  This module is plumbing only; it never chooses an
implementation on behalf of the caller and has no fallback.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
PRODUCT_LIB = ROOT / "criteria3d_b200" / "libsf3d_b200.so"


class SF3Derror(enum.IntEnum):  # types.h:39-40
    SF3Dok = 0
    IndexError = 1
    MemoryError = 2
    TopographyError = 3
    BoundaryError = 4
    MissingDataError = 5
    ParameterError = 6
    SolverError = 7
    FileError = 8


class BoundaryType(enum.IntEnum):  # types.h:98-99
    NoBoundary = 0
    Runoff = 1
    FreeDrainage = 2
    FreeLateralDrainage = 3
    PrescribedTotalWaterPotential = 4
    Urban = 5
    Road = 6
    Culvert = 7
    HeatSurface = 8
    SoluteFlux = 9


class LinkType(enum.IntEnum):  # types.h:101
    NoLink = 0
    Up = 1
    Down = 2
    Lateral = 3


class WRCModel(enum.IntEnum):  # types.h:135
    VanGenuchten = 0
    ModifiedVanGenuchten = 1
    Campbell = 2


class MeanType(enum.IntEnum):  # types.h:36
    Arithmetic = 0
    Geometric = 1
    Logarithmic = 2


class HeatFluxSaveMode(enum.IntEnum):  # types.h:186
    None_ = 0
    Total = 1
    All = 2


class FluxType(enum.IntEnum):  # types.h:199
    HeatTotal = 0
    HeatDiffusive = 1
    HeatLatentIsothermal = 2
    HeatLatentThermal = 3
    HeatAdvective = 4
    WaterLiquidIsothermal = 5
    WaterLiquidThermal = 6
    WaterVaporIsothermal = 7
    WaterVaporThermal = 8


class Field(enum.IntEnum):  # include/sf3d.h enum sf3d_field
    WATER_CONTENT = 0
    DEGREE_OF_SATURATION = 1
    WATER_CONDUCTIVITY = 2
    MATRIC_POTENTIAL = 3
    TOTAL_POTENTIAL = 4
    POND = 5
    WATER_SINK_SOURCE = 6
    BOUNDARY_WATER_FLOW = 7
    PRESCRIBED_POTENTIAL = 8
    TEMPERATURE = 9
    HEAT_SINK_SOURCE = 10
    SUM_LATERAL_FLOW = 11
    MAX_FLOW_UP = 12
    MAX_FLOW_DOWN = 13
    MAX_FLOW_LATERAL = 14
    HEAT_CONDUCTIVITY = 15
    BOUNDARY_NET_IRRADIANCE = 16
    BOUNDARY_TEMPERATURE = 17
    BOUNDARY_RELATIVE_HUMIDITY = 18
    BOUNDARY_WIND_SPEED = 19
    BOUNDARY_HEIGHT_WIND = 20
    BOUNDARY_HEIGHT_TEMPERATURE = 21
    BOUNDARY_ROUGHNESS = 22


# sentinel doubles of getDoubleErrorValue (types.h:42-64)
INDEX_ERROR = -1111.0
MEMORY_ERROR = -2222.0
TOPOGRAPHY_ERROR = -3333.0
BOUNDARY_ERROR = -4444.0
MISSING_DATA_ERROR = -9999.0
PARAMETER_ERROR = -7777.0
NODATA = -9999.0


class GridDesc(C.Structure):
    _fields_ = [
        ("rows", C.c_uint32), ("cols", C.c_uint32), ("layers", C.c_uint32), ("n_valid", C.c_uint32),
        ("cell", C.c_double), ("x_ll", C.c_double), ("y_ll", C.c_double),
        ("dem", C.POINTER(C.c_float)), ("slope_tan", C.POINTER(C.c_float)),
        ("cell_rank", C.POINTER(C.c_int32)), ("outlet", C.POINTER(C.c_uint8)),
        ("soil_id", C.POINTER(C.c_uint16)), ("surface_id", C.POINTER(C.c_uint16)),
        ("pond", C.POINTER(C.c_double)),
        ("layer_depth", C.POINTER(C.c_double)), ("layer_thickness", C.POINTER(C.c_double)),
        ("layer_horizon", C.POINTER(C.c_uint16)), ("boundary_l1", C.POINTER(C.c_uint8)),
        ("free_catchment_runoff", C.c_int), ("free_lateral_drainage", C.c_int),
        ("free_bottom_drainage", C.c_int), ("heat_surface_layer1", C.c_int),
    ]


class ForcingDesc(C.Structure):
    _fields_ = [
        ("rows", C.c_uint32), ("cols", C.c_uint32),
        ("precipitation", C.POINTER(C.c_float)), ("precipitation_nodata", C.c_float),
        ("n_sink_layers", C.c_uint32), ("layer_sink", C.POINTER(C.c_float)), ("sink_nodata", C.c_float),
        ("accumulate", C.c_int),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("steps", C.c_uint64), ("tries", C.c_uint64), ("approximations", C.c_uint64),
        ("sweeps", C.c_uint64), ("heat_steps", C.c_uint64), ("heat_sweeps", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("delta_t_curr", C.c_double), ("last_courant", C.c_double),
        ("last_mbr", C.c_double), ("last_mbe", C.c_double), ("links", C.c_uint64),
        ("heat_cap_hits", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


KERNEL_NAMES = ["begin_try", "node_phase", "assemble", "jacobi", "post", "accept", "other", "heat_coeffs",
                "heat_flux_snapshot", "heat_boundary", "heat_assemble", "heat_jacobi", "heat_post", "heat_accept", "comm"]


class KernelTimes(C.Structure):
    _fields_ = [("ms", C.c_double * 15), ("launches", C.c_uint64 * 15)]


u8, u16, u32, dbl, cint = C.c_uint8, C.c_uint16, C.c_uint32, C.c_double, C.c_int

# (C symbol, reference name, restype, argtypes) -- one row per soilFluxes3D.h declaration
_API = [
    ("sf3d_initialize", "initializeSF3D", u8, [u32, u32, u8, cint, cint, cint, u8]),
    ("sf3d_initialize_balance", "initializeBalance", u8, []),
    ("sf3d_clean", "cleanSF3D", u8, []),
    ("sf3d_initialize_heat_flag", "initializeHeatFlag", u8, [u8, cint, cint]),
    ("sf3d_set_threads_number", "setThreadsNumber", u32, [u32]),
    ("sf3d_set_use_lineal", "setUseLineal", None, [cint]),
    ("sf3d_set_lineal_method", "setLinealMethod", None, [cint]),
    ("sf3d_set_soil_properties", "setSoilProperties", u8, [u16, u8] + [dbl] * 10),
    ("sf3d_set_surface_properties", "setSurfaceProperties", u8, [u16, dbl]),
    ("sf3d_set_numerical_parameters", "setNumericalParameters", u8, [dbl, dbl, u16, u16, u8, u8]),
    ("sf3d_set_hydraulic_properties", "setHydraulicProperties", u8, [u8, u8, C.c_float]),
    ("sf3d_set_culvert", "setCulvert", u8, [u32, dbl, dbl, dbl, dbl]),
    ("sf3d_set_node", "setNode", u8, [u32, dbl, dbl, dbl, dbl, cint, u8, dbl, dbl]),
    ("sf3d_set_node_link", "setNodeLink", u8, [u32, u32, u8, dbl]),
    ("sf3d_set_node_boundary", "setNodeBoundary", u8, [u32, u8, dbl, dbl]),
    ("sf3d_set_node_soil", "setNodeSoil", u8, [u32, u16, u16]),
    ("sf3d_set_node_surface", "setNodeSurface", u8, [u32, u16]),
    ("sf3d_set_node_pond", "setNodePond", u8, [u32, dbl]),
    ("sf3d_set_node_water_content", "setNodeWaterContent", u8, [u32, dbl]),
    ("sf3d_set_node_degree_of_saturation", "setNodeDegreeOfSaturation", u8, [u32, dbl]),
    ("sf3d_set_node_matric_potential", "setNodeMatricPotential", u8, [u32, dbl]),
    ("sf3d_set_node_total_potential", "setNodeTotalPotential", u8, [u32, dbl]),
    ("sf3d_set_node_water_sink_source", "setNodeWaterSinkSource", u8, [u32, dbl]),
    ("sf3d_set_node_prescribed_total_potential", "setNodePrescribedTotalPotential", u8, [u32, dbl]),
    ("sf3d_get_node_water_content", "getNodeWaterContent", dbl, [u32]),
    ("sf3d_get_node_maximum_water_content", "getNodeMaximumWaterContent", dbl, [u32]),
    ("sf3d_get_node_minimum_water_content", "getNodeMinimumWaterContent", dbl, [u32]),
    ("sf3d_get_node_available_water_content", "getNodeAvailableWaterContent", dbl, [u32]),
    ("sf3d_get_node_water_deficit", "getNodeWaterDeficit", dbl, [u32, dbl]),
    ("sf3d_get_node_degree_of_saturation", "getNodeDegreeOfSaturation", dbl, [u32]),
    ("sf3d_get_node_water_conductivity", "getNodeWaterConductivity", dbl, [u32]),
    ("sf3d_get_node_matric_potential", "getNodeMatricPotential", dbl, [u32]),
    ("sf3d_get_node_total_potential", "getNodeTotalPotential", dbl, [u32]),
    ("sf3d_get_node_pond", "getNodePond", dbl, [u32]),
    ("sf3d_get_node_max_water_flow", "getNodeMaxWaterFlow", dbl, [u32, u8]),
    ("sf3d_get_node_sum_lateral_water_flow", "getNodeSumLateralWaterFlow", dbl, [u32]),
    ("sf3d_get_node_sum_lateral_water_flow_in", "getNodeSumLateralWaterFlowIn", dbl, [u32]),
    ("sf3d_get_node_sum_lateral_water_flow_out", "getNodeSumLateralWaterFlowOut", dbl, [u32]),
    ("sf3d_get_node_boundary_water_flow", "getNodeBoundaryWaterFlow", dbl, [u32]),
    ("sf3d_get_total_boundary_water_flow", "getTotalBoundaryWaterFlow", dbl, [u8]),
    ("sf3d_get_total_water_content", "getTotalWaterContent", dbl, []),
    ("sf3d_get_water_storage", "getWaterStorage", dbl, []),
    ("sf3d_get_water_mbr", "getWaterMBR", dbl, []),
    ("sf3d_set_node_heat_sink_source", "setNodeHeatSinkSource", u8, [u32, dbl]),
    ("sf3d_set_node_temperature", "setNodeTemperature", u8, [u32, dbl]),
    ("sf3d_set_node_boundary_fixed_temperature", "setNodeBoundaryFixedTemperature", u8, [u32, dbl, dbl]),
    ("sf3d_set_node_boundary_height_wind", "setNodeBoundaryHeightWind", u8, [u32, dbl]),
    ("sf3d_set_node_boundary_height_temperature", "setNodeBoundaryHeightTemperature", u8, [u32, dbl]),
    ("sf3d_set_node_boundary_net_irradiance", "setNodeBoundaryNetIrradiance", u8, [u32, dbl]),
    ("sf3d_set_node_boundary_temperature", "setNodeBoundaryTemperature", u8, [u32, dbl]),
    ("sf3d_set_node_boundary_relative_humidity", "setNodeBoundaryRelativeHumidity", u8, [u32, dbl]),
    ("sf3d_set_node_boundary_roughness", "setNodeBoundaryRoughness", u8, [u32, dbl]),
    ("sf3d_set_node_boundary_wind_speed", "setNodeBoundaryWindSpeed", u8, [u32, dbl]),
    ("sf3d_get_node_temperature", "getNodeTemperature", dbl, [u32]),
    ("sf3d_get_node_heat_conductivity", "getNodeHeatConductivity", dbl, [u32]),
    ("sf3d_get_node_vapor", "getNodeVapor", dbl, [u32]),
    ("sf3d_get_node_heat_storage", "getNodeHeatStorage", dbl, [u32, dbl]),
    ("sf3d_get_node_heat_max_flux", "getNodeHeatMaxFlux", dbl, [u32, u8, u8]),
    ("sf3d_get_node_boundary_advective_flux", "getNodeBoundaryAdvectiveFlux", dbl, [u32]),
    ("sf3d_get_node_boundary_latent_flux", "getNodeBoundaryLatentFlux", dbl, [u32]),
    ("sf3d_get_node_boundary_radiative_flux", "getNodeBoundaryRadiativeFlux", dbl, [u32]),
    ("sf3d_get_node_boundary_sensible_flux", "getNodeBoundarySensibleFlux", dbl, [u32]),
    ("sf3d_get_node_boundary_aerodynamic_conductance", "getNodeBoundaryAerodynamicConductance", dbl, [u32]),
    ("sf3d_get_node_boundary_soil_conductance", "getNodeBoundarySoilConductance", dbl, [u32]),
    ("sf3d_get_heat_mbr", "getHeatMBR", dbl, []),
    ("sf3d_get_heat_mbe", "getHeatMBE", dbl, []),
    ("sf3d_compute_period", "computePeriod", None, [dbl]),
    ("sf3d_compute_step", "computeStep", dbl, [dbl]),
]

_EXT = [
    ("sf3d_ext_get_field", u8, [cint, u32, u32, C.POINTER(dbl)]),
    ("sf3d_ext_set_field", u8, [cint, u32, u32, C.POINTER(dbl)]),
    ("sf3d_ext_get_link_table", u8, [u8, u32, u32, C.POINTER(u8), C.POINTER(u32), C.POINTER(dbl)]),
    ("sf3d_ext_get_node_meta", u8, [u32, u32, C.POINTER(u8), C.POINTER(u8), C.POINTER(u8)]),
    ("sf3d_ext_build_grid", u8, [C.POINTER(GridDesc)]),
    ("sf3d_ext_set_forcing_rasters", u8, [C.POINTER(ForcingDesc)]),
    ("sf3d_ext_get_layer_rasters", u8, [C.c_int, u32, u32, C.c_float, C.POINTER(C.c_float)]),
    ("sf3d_ext_get_layer_rasters_async", u8, [C.c_int, u32, u32, C.c_float, C.POINTER(C.c_float)]),
    ("sf3d_ext_wait_rasters", u8, []),
    ("sf3d_ext_set_fixed_temperature", u8, [u32, u32, C.POINTER(dbl), dbl]),
    ("sf3d_ext_get_counters", u8, [C.POINTER(Counters)]),
    ("sf3d_ext_reset_counters", u8, []),
    ("sf3d_ext_last_error", u8, []),
    ("sf3d_ext_backend", C.c_char_p, []),
    ("sf3d_ext_set_device", u8, [cint]),
    ("sf3d_ext_reset_solver", u8, []),
    ("sf3d_ext_set_time_step", u8, [dbl]),
    ("sf3d_ext_jacobi_sweep", u8, [u32, u32, C.POINTER(u8), C.POINTER(u32), C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl),
                                   C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl)]),
    ("sf3d_ext_comm_unique_id", u8, [C.POINTER(u8)]),
    ("sf3d_ext_comm_init", u8, [cint, cint, C.POINTER(u8)]),
    ("sf3d_ext_comm_finalize", u8, []),
    ("sf3d_ext_ipc_export", u8, [C.POINTER(u8)]),
    ("sf3d_ext_ipc_import", u8, [cint, C.POINTER(u8), u32, C.POINTER(u32)]),
    ("sf3d_ext_mailbox_export", u8, [C.POINTER(u8)]),
    ("sf3d_ext_mailbox_import", u8, [cint, C.POINTER(u8)]),
    ("sf3d_ext_set_halo", u8, [u32, C.POINTER(C.c_int32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32),
                               C.POINTER(u32), C.c_uint64]),
    ("sf3d_ext_stream", C.c_void_p, []),
    ("sf3d_ext_profile", u8, [cint]),
    ("sf3d_ext_get_kernel_times", u8, [C.POINTER(KernelTimes)]),
]

ALL_SYMBOLS = [s for s, *_ in _API] + [s for s, *_ in _EXT]


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class SoilFluxes3D:
    """One loaded implementation of the sf3d C ABI, under the reference's function names."""

    def __init__(self, path: os.PathLike | str):
        path = Path(path)
        if not path.exists():
            raise FileNotFoundError(
                f"{path} not found: build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self.lib = C.CDLL(str(path), mode=C.RTLD_LOCAL)
        for sym, name, res, args in _API:
            fn = getattr(self.lib, sym)
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)
        for sym, res, args in _EXT:
            fn = getattr(self.lib, sym)
            fn.restype, fn.argtypes = res, args
        self.backend = self.lib.sf3d_ext_backend().decode()

    # ---- bulk extensions ---------------------------------------------------------
    def get_field(self, field: Field, first: int, count: int, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(count, dtype=np.float64)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size >= count
        rc = self.lib.sf3d_ext_get_field(int(field), first, count, _ptr(out, dbl))
        if rc:
            raise RuntimeError(f"sf3d_ext_get_field({field!r}) -> {SF3Derror(rc).name}")
        return out

    def set_field(self, field: Field, first: int, values: np.ndarray) -> int:
        v = np.ascontiguousarray(values, dtype=np.float64)
        return self.lib.sf3d_ext_set_field(int(field), first, v.size, _ptr(v, dbl))

    def link_table(self, slot: int, first: int, count: int):
        lt = np.empty(count, np.uint8)
        li = np.empty(count, np.uint32)
        ar = np.empty(count, np.float64)
        rc = self.lib.sf3d_ext_get_link_table(slot, first, count, _ptr(lt, u8), _ptr(li, u32), _ptr(ar, dbl))
        if rc:
            raise RuntimeError(f"sf3d_ext_get_link_table -> {SF3Derror(rc).name}")
        return lt, li, ar

    def node_meta(self, first: int, count: int):
        sfl = np.empty(count, np.uint8)
        bt = np.empty(count, np.uint8)
        nl = np.empty(count, np.uint8)
        rc = self.lib.sf3d_ext_get_node_meta(first, count, _ptr(sfl, u8), _ptr(bt, u8), _ptr(nl, u8))
        if rc:
            raise RuntimeError(f"sf3d_ext_get_node_meta -> {SF3Derror(rc).name}")
        return sfl, bt, nl

    def build_grid(self, desc: GridDesc) -> int:
        return self.lib.sf3d_ext_build_grid(C.byref(desc))

    def set_forcing_rasters(self, precipitation=None, layer_sink=None, *, nodata: float = -9999.0,
                            accumulate: bool = False) -> int:
        """Hourly forcing from float rasters [mm h-1] (assignPrecipitation / assignETreal / setSinkSource):
        precipitation [rows, cols]; layer_sink [n_layers, rows, cols] water removed per layer (0 = surface)."""
        d = ForcingDesc()
        keep = []
        if precipitation is not None:
            p = np.ascontiguousarray(precipitation, dtype=np.float32); keep.append(p)
            d.rows, d.cols = p.shape
            d.precipitation = _ptr(p, C.c_float)
        if layer_sink is not None:
            q = np.ascontiguousarray(layer_sink, dtype=np.float32); keep.append(q)
            d.n_sink_layers, d.rows, d.cols = q.shape
            d.layer_sink = _ptr(q, C.c_float)
        d.precipitation_nodata = nodata; d.sink_nodata = nodata
        d.accumulate = int(accumulate)
        return self.lib.sf3d_ext_set_forcing_rasters(C.byref(d))

    def get_layer_raster(self, field: int, layer: int, shape, nodata: float = -9999.0) -> np.ndarray:
        """Output map of one layer (computeCriteria3DMap): float32 [rows, cols]."""
        return self.get_layer_rasters(field, layer, 1, shape, nodata)[0]

    def get_layer_rasters(self, field: int, first_layer: int, n_layers: int, shape, nodata: float = -9999.0,
                          out: np.ndarray | None = None) -> np.ndarray:
        """Output maps of n_layers consecutive layers: float32 [n_layers, rows, cols]."""
        if out is None:
            out = np.empty((n_layers, *shape), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == n_layers * shape[0] * shape[1]
        rc = self.lib.sf3d_ext_get_layer_rasters(int(field), first_layer, n_layers, nodata, _ptr(out, C.c_float))
        if rc:
            raise RuntimeError(f"sf3d_ext_get_layer_rasters -> {SF3Derror(rc).name}")
        return out

    def get_layer_rasters_async(self, field: int, first_layer: int, n_layers: int, shape, out: np.ndarray, nodata: float = -9999.0) -> None:
        """The same maps, copied to `out` (page-locked float32 [n_layers, rows, cols]) while later calls run; read `out`
        only after wait_rasters()."""
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == n_layers * shape[0] * shape[1]
        rc = self.lib.sf3d_ext_get_layer_rasters_async(int(field), first_layer, n_layers, nodata, _ptr(out, C.c_float))
        if rc:
            raise RuntimeError(f"sf3d_ext_get_layer_rasters_async -> {SF3Derror(rc).name}")

    def wait_rasters(self) -> None:
        rc = self.lib.sf3d_ext_wait_rasters()
        if rc:
            raise RuntimeError(f"sf3d_ext_wait_rasters -> {SF3Derror(rc).name}")

    def set_fixed_temperature(self, first: int, temperature: np.ndarray, depth: float) -> int:
        t = np.ascontiguousarray(temperature, dtype=np.float64)
        return self.lib.sf3d_ext_set_fixed_temperature(first, t.size, _ptr(t, dbl), depth)

    def counters(self) -> dict:
        c = Counters()
        rc = self.lib.sf3d_ext_get_counters(C.byref(c))
        if rc:
            raise RuntimeError(f"sf3d_ext_get_counters -> {SF3Derror(rc).name}")
        return c.as_dict()

    def reset_counters(self) -> None:
        self.lib.sf3d_ext_reset_counters()

    def set_device(self, device: int) -> int:
        return self.lib.sf3d_ext_set_device(device)

    # ---- multi-GPU slabs (product only) -----------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = (u8 * 128)()
        rc = self.lib.sf3d_ext_comm_unique_id(buf)
        if rc:
            raise RuntimeError(f"sf3d_ext_comm_unique_id -> {SF3Derror(rc).name}")
        return bytes(buf)

    def comm_init(self, rank: int, world: int, uid: bytes | None) -> int:
        """uid None: peer memory only, no NCCL communicator"""
        buf = (u8 * 128).from_buffer_copy(uid) if uid is not None else None
        return self.lib.sf3d_ext_comm_init(rank, world, buf)

    def ipc_export(self) -> bytes:
        buf = (u8 * 128)()
        rc = self.lib.sf3d_ext_ipc_export(buf)
        if rc:
            raise RuntimeError(f"sf3d_ext_ipc_export -> {SF3Derror(rc).name}")
        return bytes(buf)

    def ipc_import(self, peer: int, handles: bytes, remote_idx) -> int:
        buf = (u8 * 128).from_buffer_copy(handles)
        r = np.ascontiguousarray(remote_idx, dtype=np.uint32)
        return self.lib.sf3d_ext_ipc_import(peer, buf, r.size, _ptr(r, u32))

    def mailbox_export(self) -> bytes:
        buf = (u8 * 64)()
        rc = self.lib.sf3d_ext_mailbox_export(buf)
        if rc:
            raise RuntimeError(f"sf3d_ext_mailbox_export -> {SF3Derror(rc).name}")
        return bytes(buf)

    def mailbox_import(self, peer: int, handle: bytes) -> int:
        return self.lib.sf3d_ext_mailbox_import(peer, (u8 * 64).from_buffer_copy(handle))

    def comm_finalize(self) -> int:
        return self.lib.sf3d_ext_comm_finalize()

    def set_halo(self, peers, send_lists, recv_lists, n_global_nodes: int) -> int:
        peers_a = np.ascontiguousarray(peers, dtype=np.int32)
        sc = np.ascontiguousarray([len(x) for x in send_lists], dtype=np.uint32)
        rc_ = np.ascontiguousarray([len(x) for x in recv_lists], dtype=np.uint32)
        si = np.ascontiguousarray(np.concatenate(send_lists) if len(send_lists) else np.zeros(0), dtype=np.uint32)
        ri = np.ascontiguousarray(np.concatenate(recv_lists) if len(recv_lists) else np.zeros(0), dtype=np.uint32)
        return self.lib.sf3d_ext_set_halo(len(peers_a), _ptr(peers_a, C.c_int32), _ptr(sc, u32), _ptr(si, u32),
                                          _ptr(rc_, u32), _ptr(ri, u32), int(n_global_nodes))

    def jacobi_sweep(self, n_surface, ncols, col, val, b, z, x_in):
        """one sweep on a system in the reference's compact row layout; returns (x_out, norm)"""
        n = len(b)
        ncols = np.ascontiguousarray(ncols, np.uint8); col = np.ascontiguousarray(col, np.uint32)
        val = np.ascontiguousarray(val, np.float64); b = np.ascontiguousarray(b, np.float64)
        z = np.ascontiguousarray(z, np.float64); x_in = np.ascontiguousarray(x_in, np.float64)
        x_out = np.empty(n, np.float64)
        norm = dbl(0.0)
        rc = self.lib.sf3d_ext_jacobi_sweep(n, n_surface, _ptr(ncols, u8), _ptr(col, u32), _ptr(val, dbl), _ptr(b, dbl),
                                            _ptr(z, dbl), _ptr(x_in, dbl), _ptr(x_out, dbl), C.byref(norm))
        if rc:
            raise RuntimeError(f"sf3d_ext_jacobi_sweep -> {SF3Derror(rc).name}")
        return x_out, norm.value

    def last_error(self) -> int:
        """error of the most recent computeStep / computePeriod (then cleared)"""
        return self.lib.sf3d_ext_last_error()

    def reset_solver(self) -> int:
        return self.lib.sf3d_ext_reset_solver()

    def set_time_step(self, delta_t: float) -> int:
        return self.lib.sf3d_ext_set_time_step(delta_t)

    def stream(self) -> int | None:
        return self.lib.sf3d_ext_stream()

    def profile(self, enable: bool) -> int:
        return self.lib.sf3d_ext_profile(1 if enable else 0)

    def kernel_times(self) -> dict:
        kt = KernelTimes()
        rc = self.lib.sf3d_ext_get_kernel_times(C.byref(kt))
        if rc:
            raise RuntimeError(f"sf3d_ext_get_kernel_times -> {SF3Derror(rc).name}")
        return {n: {"ms": kt.ms[i], "launches": int(kt.launches[i])} for i, n in enumerate(KERNEL_NAMES)}


def load_product() -> SoilFluxes3D:
    """The CUDA product.  Fails loudly when the library is missing: there is no fallback."""
    if not PRODUCT_LIB.exists():
        raise RuntimeError(
            f"{PRODUCT_LIB} is missing. The B200 product is CUDA only (no CPU fallback): "
            "run `python -c 'import __graft_entry__ as g; g.build()'` first.")
    return SoilFluxes3D(PRODUCT_LIB)
