/*
 * ref_capi.cpp -- TEST INFRASTRUCTURE.  The UNMODIFIED reference library
 * (/root/reference/agrolib/soilFluxes3D/*.cpp, compiled where it lies by oracle/Makefile)
 * placed behind the C ABI of include/sf3d.h, so that the same harness drives the product,
 * the CPU restatement and the reference itself.  Nothing here computes: every function
 * forwards to the soilFluxes3D::v2 call of the same name.  Output: oracle/_ref/libsf3d_ref.so
 * (git-ignored).  Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline load it.
 *
 * Counters: the reference keeps no sweep/approximation counters, so the link step uses
 * GNU ld --wrap on two functions that cpusolver.o calls across translation units
 * (Water::JacobiWaterCPU, water.cpp:565; Water::computeCapacity, water.cpp:279;
 * Heat::GaussSeidelHeatCPU, heat.cpp:664).  The wrappers below count and call the real
 * function; no reference source is edited.
 */
#include <cmath>
#include <cstring>
#include <vector>
#include "soilFluxes3D.h"      // the reference's own header (-I /root/reference/...)
#include "types_cpu.h"
#include "solver.h"
#include "sf3d.h"

namespace sf = soilFluxes3D::v2;

#include "cpusolver.h"
namespace soilFluxes3D::v2 {
    extern CPUSolver CPUSolverObject;
    extern nodesData_t nodeGrid;
    extern Solver* solver;
    extern balanceData_t balanceDataCurrentTimeStep;
}

static sf3d_counters g_cnt;
/* heat_cap_hits bookkeeping: tolerance and sweep cap of the heat solve (defaults of SolverParameters, types.h:291-315) */
static double g_tolerance = 1e-10;
static uint32_t g_heatCap = 200, g_heatSolveSweeps = 0;

/* optional capture of the linear system seen by the last Jacobi call (kernel-level KATs) */
static int g_capture = 0;
struct Capture {
    uint32_t n = 0;
    std::vector<uint8_t> ncols;
    std::vector<uint32_t> col;     // n * 11
    std::vector<double> val;       // n * 11
    std::vector<double> b, x_in, x_out;
    double norm = 0;
} g_cap;

extern "C" {
double __real__ZN12soilFluxes3D2v25Water14JacobiWaterCPUERNS0_9VectorCPUES3_RKNS0_9MatrixCPUERKS2_(
    sf::VectorCPU&, sf::VectorCPU&, const sf::MatrixCPU&, const sf::VectorCPU&);
double __wrap__ZN12soilFluxes3D2v25Water14JacobiWaterCPUERNS0_9VectorCPUES3_RKNS0_9MatrixCPUERKS2_(
    sf::VectorCPU& x, sf::VectorCPU& xn, const sf::MatrixCPU& A, const sf::VectorCPU& b)
{
    ++g_cnt.sweeps;
    if (g_capture)
    {
        const uint32_t n = A.numRows;
        g_cap.n = n;
        g_cap.ncols.assign(A.numColsInRow, A.numColsInRow + n);
        g_cap.col.assign((size_t)n * 11, 0u);
        g_cap.val.assign((size_t)n * 11, 0.0);
        for (uint32_t r = 0; r < n; ++r)
            for (uint8_t c = 0; c < A.numColsInRow[r]; ++c)
            {
                g_cap.col[(size_t)r * 11 + c] = A.columnIndeces[r][c];
                g_cap.val[(size_t)r * 11 + c] = A.values[r][c];
            }
        g_cap.b.assign(b.values, b.values + n);
        g_cap.x_in.assign(x.values, x.values + n);
    }
    double norm = __real__ZN12soilFluxes3D2v25Water14JacobiWaterCPUERNS0_9VectorCPUES3_RKNS0_9MatrixCPUERKS2_(x, xn, A, b);
    if (g_capture)
    {
        g_cap.x_out.assign(x.values, x.values + A.numRows);
        g_cap.norm = norm;
    }
    return norm;
}

void __real__ZN12soilFluxes3D2v25Water15computeCapacityERNS0_9VectorCPUE(sf::VectorCPU&);
void __wrap__ZN12soilFluxes3D2v25Water15computeCapacityERNS0_9VectorCPUE(sf::VectorCPU& c)
{
    ++g_cnt.approximations;
    __real__ZN12soilFluxes3D2v25Water15computeCapacityERNS0_9VectorCPUE(c);
}

double __real__ZN12soilFluxes3D2v24Heat18GaussSeidelHeatCPUERNS0_9VectorCPUERKNS0_9MatrixCPUERKS2_(
    sf::VectorCPU&, const sf::MatrixCPU&, const sf::VectorCPU&);
double __wrap__ZN12soilFluxes3D2v24Heat18GaussSeidelHeatCPUERNS0_9VectorCPUERKNS0_9MatrixCPUERKS2_(
    sf::VectorCPU& x, const sf::MatrixCPU& A, const sf::VectorCPU& b)
{
    ++g_cnt.heat_sweeps;
    const double norm = __real__ZN12soilFluxes3D2v24Heat18GaussSeidelHeatCPUERNS0_9VectorCPUERKNS0_9MatrixCPUERKS2_(x, A, b);
    /* CPUSolver::solveLinearSystem (cpusolver.cpp:676-700) leaves its loop when the norm is below the
       tolerance or after calcCurrentMaxIterationNumber(maxApprox - 1) sweeps: count the second exit */
    ++g_heatSolveSweeps;
    if (norm < g_tolerance) g_heatSolveSweeps = 0;
    else if (g_heatSolveSweeps >= g_heatCap) { ++g_cnt.heat_cap_hits; g_heatSolveSweeps = 0; }
    return norm;
}
/* accepted heat sub-steps: Heat::updateHeatBalanceData (heat.cpp:393) is called once per accepted
   heatLoop (cpusolver.cpp:593) */
void __real__ZN12soilFluxes3D2v24Heat21updateHeatBalanceDataEv(void);
void __wrap__ZN12soilFluxes3D2v24Heat21updateHeatBalanceDataEv(void)
{
    ++g_cnt.heat_steps;
    __real__ZN12soilFluxes3D2v24Heat21updateHeatBalanceDataEv();
}
} // extern "C"

#define E8(call) static_cast<uint8_t>(call)
#define BT(v) static_cast<sf::boundaryType_t>(v)
#define LT(v) static_cast<sf::linkType_t>(v)

extern "C" {

uint8_t sf3d_initialize(uint32_t n, uint32_t ns, uint8_t nl, int w, int h, int s, uint8_t hf)
{
    std::memset(&g_cnt, 0, sizeof g_cnt);
    return E8(sf::initializeSF3D(n, ns, nl, w != 0, h != 0, s != 0, static_cast<sf::heatFluxSaveMode_t>(hf)));
}
uint8_t  sf3d_initialize_balance(void) { return E8(sf::initializeBalance()); }
uint8_t  sf3d_clean(void) { return E8(sf::cleanSF3D()); }
uint8_t  sf3d_initialize_heat_flag(uint8_t m, int adv, int lat)
{ return E8(sf::initializeHeatFlag(static_cast<sf::heatFluxSaveMode_t>(m), adv != 0, lat != 0)); }
uint32_t sf3d_set_threads_number(uint32_t n) { return sf::setThreadsNumber(n); }
void     sf3d_set_use_lineal(int v) { sf::setUseLineal(v != 0); }
void     sf3d_set_lineal_method(int v) { sf::setLinealMethod(v); }

uint8_t sf3d_set_soil_properties(uint16_t a, uint8_t b, double c, double d, double e, double f, double g,
                                 double h, double i, double j, double k, double l)
{ return E8(sf::setSoilProperties(a, b, c, d, e, f, g, h, i, j, k, l)); }
uint8_t sf3d_set_surface_properties(uint16_t i, double r) { return E8(sf::setSurfaceProperties(i, r)); }
uint8_t sf3d_set_numerical_parameters(double a, double b, uint16_t c, uint16_t d, uint8_t e, uint8_t f)
{
    /* mirror of the clamps (soilFluxes3D.cpp:491-504) and of Solver::calcCurrentMaxIterationNumber (solver.h:55-59),
       only for the heat_cap_hits counter */
    const uint16_t it = c < 20 ? 20 : (c > 1000 ? 1000 : c), ap = d < 1 ? 1 : (d > 50 ? 50 : d);
    const uint8_t ex = e < 5 ? 5 : (e > 12 ? 12 : e);
    g_tolerance = std::pow(10.0, -ex);
    const uint32_t cap = static_cast<uint32_t>(ap * (static_cast<float>(it) / static_cast<float>(ap)));
    g_heatCap = cap < 25u ? 25u : cap;
    return E8(sf::setNumericalParameters(a, b, c, d, e, f));
}
uint8_t sf3d_set_hydraulic_properties(uint8_t w, uint8_t m, float r)
{ return E8(sf::setHydraulicProperties(static_cast<sf::WRCModel>(w), static_cast<sf::meanType_t>(m), r)); }

uint8_t sf3d_set_culvert(uint32_t i, double r, double s, double w, double h)
{
    /* soilFluxes3D.cpp:586 writes through nodeGrid.culvertPtr, which the reference never
       allocates (:146 commented out): calling it would crash.  See SURVEY Appendix B Q5. */
    (void)i; (void)r; (void)s; (void)w; (void)h;
    return SF3D_MEMORY_ERROR;
}
uint8_t sf3d_set_node(uint32_t i, double x, double y, double z, double v, int surf, uint8_t bt, double sl, double ba)
{ return E8(sf::setNode(i, x, y, z, v, surf != 0, BT(bt), sl, ba)); }
uint8_t sf3d_set_node_link(uint32_t i, uint32_t j, uint8_t d, double a) { return E8(sf::setNodeLink(i, j, LT(d), a)); }
uint8_t sf3d_set_node_boundary(uint32_t i, uint8_t bt, double s, double a)
{
    /* the reference does no checks here (soilFluxes3D.cpp:689-725); guard the raw index */
    if (!sf::nodeGrid.isInitialized) return SF3D_MEMORY_ERROR;
    if (i >= sf::nodeGrid.nrNodes) return SF3D_INDEX_ERROR;
    return E8(sf::setNodeBoundary(i, BT(bt), s, a));
}
uint8_t sf3d_set_node_soil(uint32_t i, uint16_t s, uint16_t h) { return E8(sf::setNodeSoil(i, s, h)); }
uint8_t sf3d_set_node_surface(uint32_t i, uint16_t s) { return E8(sf::setNodeSurface(i, s)); }

uint8_t sf3d_set_node_pond(uint32_t i, double v) { return E8(sf::setNodePond(i, v)); }
uint8_t sf3d_set_node_water_content(uint32_t i, double v) { return E8(sf::setNodeWaterContent(i, v)); }
uint8_t sf3d_set_node_degree_of_saturation(uint32_t i, double v) { return E8(sf::setNodeDegreeOfSaturation(i, v)); }
uint8_t sf3d_set_node_matric_potential(uint32_t i, double v) { return E8(sf::setNodeMatricPotential(i, v)); }
uint8_t sf3d_set_node_total_potential(uint32_t i, double v) { return E8(sf::setNodeTotalPotential(i, v)); }
uint8_t sf3d_set_node_water_sink_source(uint32_t i, double v) { return E8(sf::setNodeWaterSinkSource(i, v)); }
uint8_t sf3d_set_node_prescribed_total_potential(uint32_t i, double v) { return E8(sf::setNodePrescribedTotalPotential(i, v)); }

double sf3d_get_node_water_content(uint32_t i) { return sf::getNodeWaterContent(i); }
double sf3d_get_node_maximum_water_content(uint32_t i) { return sf::getNodeMaximumWaterContent(i); }
double sf3d_get_node_minimum_water_content(uint32_t i) { return sf::getNodeMinimumWaterContent(i); }
double sf3d_get_node_available_water_content(uint32_t i) { return sf::getNodeAvailableWaterContent(i); }
double sf3d_get_node_water_deficit(uint32_t i, double fc) { return sf::getNodeWaterDeficit(i, fc); }
double sf3d_get_node_degree_of_saturation(uint32_t i) { return sf::getNodeDegreeOfSaturation(i); }
double sf3d_get_node_water_conductivity(uint32_t i) { return sf::getNodeWaterConductivity(i); }
double sf3d_get_node_matric_potential(uint32_t i) { return sf::getNodeMatricPotential(i); }
double sf3d_get_node_total_potential(uint32_t i) { return sf::getNodeTotalPotential(i); }
double sf3d_get_node_pond(uint32_t i) { return sf::getNodePond(i); }
double sf3d_get_node_max_water_flow(uint32_t i, uint8_t d) { return sf::getNodeMaxWaterFlow(i, LT(d)); }
double sf3d_get_node_sum_lateral_water_flow(uint32_t i) { return sf::getNodeSumLateralWaterFlow(i); }
double sf3d_get_node_sum_lateral_water_flow_in(uint32_t i) { return sf::getNodeSumLateralWaterFlowIn(i); }
double sf3d_get_node_sum_lateral_water_flow_out(uint32_t i) { return sf::getNodeSumLateralWaterFlowOut(i); }
double sf3d_get_node_boundary_water_flow(uint32_t i) { return sf::getNodeBoundaryWaterFlow(i); }
double sf3d_get_total_boundary_water_flow(uint8_t bt) { return sf::getTotalBoundaryWaterFlow(BT(bt)); }
double sf3d_get_total_water_content(void) { return sf::getTotalWaterContent(); }
double sf3d_get_water_storage(void) { return sf::getWaterStorage(); }
double sf3d_get_water_mbr(void) { return sf::getWaterMBR(); }

uint8_t sf3d_set_node_heat_sink_source(uint32_t i, double v) { return E8(sf::setNodeHeatSinkSource(i, v)); }
uint8_t sf3d_set_node_temperature(uint32_t i, double v) { return E8(sf::setNodeTemperature(i, v)); }
uint8_t sf3d_set_node_boundary_fixed_temperature(uint32_t i, double t, double d) { return E8(sf::setNodeBoundaryFixedTemperature(i, t, d)); }
uint8_t sf3d_set_node_boundary_height_wind(uint32_t i, double v) { return E8(sf::setNodeBoundaryHeightWind(i, v)); }
uint8_t sf3d_set_node_boundary_height_temperature(uint32_t i, double v) { return E8(sf::setNodeBoundaryHeightTemperature(i, v)); }
uint8_t sf3d_set_node_boundary_net_irradiance(uint32_t i, double v) { return E8(sf::setNodeBoundaryNetIrradiance(i, v)); }
uint8_t sf3d_set_node_boundary_temperature(uint32_t i, double v) { return E8(sf::setNodeBoundaryTemperature(i, v)); }
uint8_t sf3d_set_node_boundary_relative_humidity(uint32_t i, double v) { return E8(sf::setNodeBoundaryRelativeHumidity(i, v)); }
uint8_t sf3d_set_node_boundary_roughness(uint32_t i, double v) { return E8(sf::setNodeBoundaryRoughness(i, v)); }
uint8_t sf3d_set_node_boundary_wind_speed(uint32_t i, double v) { return E8(sf::setNodeBoundaryWindSpeed(i, v)); }

double sf3d_get_node_temperature(uint32_t i) { return sf::getNodeTemperature(i); }
double sf3d_get_node_heat_conductivity(uint32_t i) { return sf::getNodeHeatConductivity(i); }
double sf3d_get_node_vapor(uint32_t i) { return sf::getNodeVapor(i); }
double sf3d_get_node_heat_storage(uint32_t i, double h) { return sf::getNodeHeatStorage(i, h); }
double sf3d_get_node_heat_max_flux(uint32_t i, uint8_t d, uint8_t f) { return sf::getNodeHeatMaxFlux(i, LT(d), static_cast<sf::fluxTypes_t>(f)); }
double sf3d_get_node_boundary_advective_flux(uint32_t i) { return sf::getNodeBoundaryAdvectiveFlux(i); }
double sf3d_get_node_boundary_latent_flux(uint32_t i) { return sf::getNodeBoundaryLatentFlux(i); }
double sf3d_get_node_boundary_radiative_flux(uint32_t i) { return sf::getNodeBoundaryRadiativeFlux(i); }
double sf3d_get_node_boundary_sensible_flux(uint32_t i) { return sf::getNodeBoundarySensibleFlux(i); }
double sf3d_get_node_boundary_aerodynamic_conductance(uint32_t i) { return sf::getNodeBoundaryAerodynamicConductance(i); }
double sf3d_get_node_boundary_soil_conductance(uint32_t i) { return sf::getNodeBoundarySoilConductance(i); }
double sf3d_get_heat_mbr(void) { return sf::getHeatMBR(); }
double sf3d_get_heat_mbe(void) { return sf::getHeatMBE(); }

void sf3d_compute_period(double t)
{
    /* computePeriod (soilFluxes3D.cpp:1760-1777) loops on computeStep internally */
    sf::computePeriod(t);
}
double sf3d_compute_step(double maxDt)
{
    double dt = sf::computeStep(maxDt);
    ++g_cnt.steps;
    return dt;
}

/* ---------------- extensions: plain loops over the scalar API ---------------- */
#include "field_loops.inc"
#include "raster_loops.inc"
#include "grid_builder_scalar.inc"

uint8_t sf3d_ext_get_link_table(uint8_t slot, uint32_t first, uint32_t count,
                                uint8_t *lt, uint32_t *li, double *area)
{
    if (!sf::nodeGrid.isInitialized) return SF3D_MEMORY_ERROR;
    if (slot >= SF3D_MAX_TOTAL_LINK || (uint64_t)first + count > sf::nodeGrid.nrNodes) return SF3D_INDEX_ERROR;
    const sf::linkData_t &ld = sf::nodeGrid.linkData[slot];
    for (uint32_t k = 0; k < count; ++k)
    {
        if (lt)   lt[k]   = static_cast<uint8_t>(ld.linkType[first + k]);
        if (li)   li[k]   = ld.linkIndex[first + k];
        if (area) area[k] = ld.interfaceArea[first + k];
    }
    return SF3D_OK;
}

uint8_t sf3d_ext_get_node_meta(uint32_t first, uint32_t count, uint8_t *sfl, uint8_t *bt, uint8_t *nl)
{
    if (!sf::nodeGrid.isInitialized) return SF3D_MEMORY_ERROR;
    if ((uint64_t)first + count > sf::nodeGrid.nrNodes) return SF3D_INDEX_ERROR;
    for (uint32_t k = 0; k < count; ++k)
    {
        if (sfl) sfl[k] = sf::nodeGrid.surfaceFlag[first + k] ? 1 : 0;
        if (bt)  bt[k]  = static_cast<uint8_t>(sf::nodeGrid.boundaryData.boundaryType[first + k]);
        if (nl)  nl[k]  = sf::nodeGrid.numLateralLink[first + k];
    }
    return SF3D_OK;
}

uint8_t sf3d_ext_get_counters(sf3d_counters *out)
{
    if (!out) return SF3D_PARAMETER_ERROR;
    g_cnt.tries = 0;            /* not observable without editing the reference */
    g_cnt.kernel_launches = 0;
    g_cnt.delta_t_curr = sf::solver ? sf::solver->getTimeStep() : -9999.;
    g_cnt.last_courant = sf::nodeGrid.CourantWater;
    g_cnt.last_mbr = sf::balanceDataCurrentTimeStep.waterMBR;
    g_cnt.last_mbe = sf::balanceDataCurrentTimeStep.waterMBE;
    g_cnt.links = 0;
    if (sf::nodeGrid.isInitialized)
        for (int s = 0; s < SF3D_MAX_TOTAL_LINK; ++s)
            for (uint32_t i = 0; i < sf::nodeGrid.nrNodes; ++i)
                g_cnt.links += (sf::nodeGrid.linkData[s].linkType[i] != sf::linkType_t::NoLink);
    *out = g_cnt;
    return SF3D_OK;
}
uint8_t sf3d_ext_reset_counters(void) { std::memset(&g_cnt, 0, sizeof g_cnt); return SF3D_OK; }
uint8_t sf3d_ext_last_error(void) { return SF3D_OK; }
const char *sf3d_ext_backend(void) { return "reference"; }
uint8_t sf3d_ext_set_device(int) { return SF3D_PARAMETER_ERROR; }
uint8_t sf3d_ext_comm_unique_id(uint8_t id[128]) { (void)id; return SF3D_PARAMETER_ERROR; }
uint8_t sf3d_ext_comm_init(int rank, int world, const uint8_t id[128]) { (void)rank; (void)world; (void)id; return SF3D_PARAMETER_ERROR; }
uint8_t sf3d_ext_comm_finalize(void) { return SF3D_PARAMETER_ERROR; }
uint8_t sf3d_ext_ipc_export(uint8_t h[128]) { (void)h; return SF3D_PARAMETER_ERROR; }
uint8_t sf3d_ext_mailbox_export(uint8_t h[64]) { (void)h; return SF3D_PARAMETER_ERROR; }
uint8_t sf3d_ext_mailbox_import(int p, const uint8_t h[64]) { (void)p; (void)h; return SF3D_PARAMETER_ERROR; }
uint8_t sf3d_ext_ipc_import(int p, const uint8_t h[128], uint32_t n, const uint32_t *r) { (void)p; (void)h; (void)n; (void)r; return SF3D_PARAMETER_ERROR; }
uint8_t sf3d_ext_set_halo(uint32_t n, const int32_t *p, const uint32_t *sc, const uint32_t *si, const uint32_t *rc, const uint32_t *ri, uint64_t ng)
{ (void)n; (void)p; (void)sc; (void)si; (void)rc; (void)ri; (void)ng; return SF3D_PARAMETER_ERROR; }
void *sf3d_ext_stream(void) { return nullptr; }
uint8_t sf3d_ext_profile(int) { return SF3D_PARAMETER_ERROR; }
uint8_t sf3d_ext_get_kernel_times(sf3d_kernel_times *) { return SF3D_PARAMETER_ERROR; }
uint8_t sf3d_ext_jacobi_sweep(uint32_t n, uint32_t ns, const uint8_t *ncols, const uint32_t *col, const double *val,
                              const double *b, const double *z, const double *x_in, double *x_out, double *norm)
{
    /* the REAL Water::JacobiWaterCPU on temporary MatrixCPU / VectorCPU objects; it reads nodeGrid.z and
       nodeGrid.nrSurfaceNodes, which are pointed at the caller's data for the duration of the call */
    if (!ncols || !col || !val || !b || !z || !x_in || !x_out || !norm || n == 0) return SF3D_PARAMETER_ERROR;
    if (!sf::solver) return SF3D_MEMORY_ERROR;          /* __ompStatus dereferences the solver */
    std::vector<uint32_t *> colp(n); std::vector<double *> valp(n);
    std::vector<uint32_t> colc(col, col + (size_t)n * 11); std::vector<double> valc(val, val + (size_t)n * 11);
    for (uint32_t r = 0; r < n; ++r) { colp[r] = &colc[(size_t)r * 11]; valp[r] = &valc[(size_t)r * 11]; }
    std::vector<uint8_t> nc(ncols, ncols + n);
    std::vector<double> x(x_in, x_in + n), xn(n, 0.0), bb(b, b + n);
    sf::MatrixCPU A; A.numRows = n; A.numColsInRow = nc.data(); A.columnIndeces = colp.data(); A.values = valp.data();
    sf::VectorCPU vx{n, x.data()}, vxn{n, xn.data()}, vb{n, bb.data()};
    double *saveZ = sf::nodeGrid.z; uint32_t saveNs = sf::nodeGrid.nrSurfaceNodes;
    sf::nodeGrid.z = const_cast<double *>(z); sf::nodeGrid.nrSurfaceNodes = ns;
    const uint64_t sweeps = g_cnt.sweeps;
    *norm = __wrap__ZN12soilFluxes3D2v25Water14JacobiWaterCPUERNS0_9VectorCPUES3_RKNS0_9MatrixCPUERKS2_(vx, vxn, A, vb);
    g_cnt.sweeps = sweeps;
    sf::nodeGrid.z = saveZ; sf::nodeGrid.nrSurfaceNodes = saveNs;
    std::memcpy(x_out, vx.values, (size_t)n * sizeof(double));      /* the function swaps the two vectors */
    return SF3D_OK;
}
uint8_t sf3d_ext_set_time_step(double delta_t)
{
    sf::SolverParametersPartial p;
    p.deltaTcurr = delta_t;
    sf::CPUSolverObject.updateParameters(p);
    return SF3D_OK;
}
uint8_t sf3d_ext_reset_solver(void)
{
    /* every field of SolverParameters back to its default (types.h:291-315); deltaTcurr = NODATA
       makes CPUSolver::initialize pick deltaTmax again (cpusolver.cpp:30-31) */
    const sf::SolverParameters d;
    sf::SolverParametersPartial p;
    p.MBRThreshold = d.MBRThreshold; p.residualTolerance = d.residualTolerance;
    p.deltaTmin = d.deltaTmin; p.deltaTmax = d.deltaTmax; p.deltaTcurr = d.deltaTcurr;
    p.maxApproximationsNumber = d.maxApproximationsNumber; p.maxIterationsNumber = d.maxIterationsNumber;
    p.waterRetentionCurveModel = d.waterRetentionCurveModel; p.meanType = d.meanType;
    p.lateralVerticalRatio = static_cast<float>(d.lateralVerticalRatio);
    sf::CPUSolverObject.updateParameters(p);
    return SF3D_OK;
}

/* reference-only helpers for kernel-level known-answer tests */
void sf3d_ref_capture_jacobi(int enable) { g_capture = enable; }
uint32_t sf3d_ref_captured_rows(void) { return g_cap.n; }
double sf3d_ref_captured_norm(void) { return g_cap.norm; }
void sf3d_ref_captured_copy(uint8_t *ncols, uint32_t *col, double *val, double *b, double *x_in, double *x_out)
{
    const size_t n = g_cap.n;
    if (ncols) std::memcpy(ncols, g_cap.ncols.data(), n);
    if (col)   std::memcpy(col, g_cap.col.data(), n * 11 * sizeof(uint32_t));
    if (val)   std::memcpy(val, g_cap.val.data(), n * 11 * sizeof(double));
    if (b)     std::memcpy(b, g_cap.b.data(), n * sizeof(double));
    if (x_in)  std::memcpy(x_in, g_cap.x_in.data(), n * sizeof(double));
    if (x_out) std::memcpy(x_out, g_cap.x_out.data(), n * sizeof(double));
}

} // extern "C"
