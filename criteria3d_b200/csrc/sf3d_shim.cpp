// sf3d_shim.cpp -- namespace soilFluxes3D::v2 (the reference's C++ plugin API,
// agrolib/soilFluxes3D/soilFluxes3D.h:9-104) forwarding 1:1 to the C ABI of include/sf3d.h.
// Same function names, argument order, default arguments and return conventions, so that
// project3D.cpp / criteria3DProject.cpp / CRITERIA-1D link against libsf3d_b200.so unchanged.
#include "soilFluxes3D.h"
#include "sf3d.h"

namespace soilFluxes3D { inline namespace v2 {

#define E(x) static_cast<SF3Derror_t>(x)
#define U(x) static_cast<std::uint8_t>(x)

// MATLAB .mat logging is out of scope (MCR_ENABLED off by default, parallel.pri:15)
SF3Derror_t initializeLog(const std::string &, const std::string &) { return SF3Derror_t::SF3Dok; }
SF3Derror_t closeLog() { return SF3Derror_t::SF3Dok; }

SF3Derror_t initializeSF3D(SF3Duint_t nrNodes, SF3Duint_t nrSurfaceNodes, u8_t nrLateralLinks, bool isComputeWater, bool isComputeHeat, bool isComputeSolutes, heatFluxSaveMode_t HFsm) { return E(sf3d_initialize(nrNodes, nrSurfaceNodes, nrLateralLinks, isComputeWater ? 1 : 0, isComputeHeat ? 1 : 0, isComputeSolutes ? 1 : 0, U(HFsm))); }
SF3Derror_t initializeBalance() { return E(sf3d_initialize_balance()); }
SF3Derror_t cleanSF3D() { return E(sf3d_clean()); }
SF3Derror_t initializeHeatFlag(heatFluxSaveMode_t saveModeHeat, bool isComputeAdvectiveFlux, bool isComputeLatentHeat) { return E(sf3d_initialize_heat_flag(U(saveModeHeat), isComputeAdvectiveFlux ? 1 : 0, isComputeLatentHeat ? 1 : 0)); }
u32_t setThreadsNumber(u32_t nrThreads) { return sf3d_set_threads_number(nrThreads); }
void setUseLineal(bool value) { sf3d_set_use_lineal(value ? 1 : 0); }
void setLinealMethod(int value) { sf3d_set_lineal_method(value); }
SF3Derror_t setSoilProperties(u16_t nrSoil, u8_t nrHorizon, double VG_alpha, double VG_n, double VG_m, double VG_he, double thetaR, double thetaS, double kSat, double MualemL, double organicMatter, double clay) { return E(sf3d_set_soil_properties(nrSoil, nrHorizon, VG_alpha, VG_n, VG_m, VG_he, thetaR, thetaS, kSat, MualemL, organicMatter, clay)); }
SF3Derror_t setSurfaceProperties(u16_t surfaceIndex, double roughness) { return E(sf3d_set_surface_properties(surfaceIndex, roughness)); }
SF3Derror_t setNumericalParameters(double minDeltaT, double maxDeltaT, u16_t maxIterationNumber, u16_t maxApproximationsNumber, u8_t ResidualToleranceExponent, u8_t MBRThresholdExponent) { return E(sf3d_set_numerical_parameters(minDeltaT, maxDeltaT, maxIterationNumber, maxApproximationsNumber, ResidualToleranceExponent, MBRThresholdExponent)); }
SF3Derror_t setHydraulicProperties(WRCModel waterRetentionCurve, meanType_t conductivityMeanType, float conductivityHorizVertRatio) { return E(sf3d_set_hydraulic_properties(U(waterRetentionCurve), U(conductivityMeanType), conductivityHorizVertRatio)); }
SF3Derror_t setCulvert(SF3Duint_t nodeIndex, double roughness, double slope, double width, double height) { return E(sf3d_set_culvert(nodeIndex, roughness, slope, width, height)); }
SF3Derror_t setNode(SF3Duint_t index, double x, double y, double z, double volume_or_area, bool isSurface, boundaryType_t boundaryType, double slope, double boundaryArea) { return E(sf3d_set_node(index, x, y, z, volume_or_area, isSurface ? 1 : 0, U(boundaryType), slope, boundaryArea)); }
SF3Derror_t setNodeLink(SF3Duint_t nodeIndex, SF3Duint_t linkIndex, linkType_t direction, double interfaceArea) { return E(sf3d_set_node_link(nodeIndex, linkIndex, U(direction), interfaceArea)); }
SF3Derror_t setNodeBoundary(SF3Duint_t nodeIndex, boundaryType_t boundaryType, double slope, double boundaryArea) { return E(sf3d_set_node_boundary(nodeIndex, U(boundaryType), slope, boundaryArea)); }
SF3Derror_t setNodeSoil(SF3Duint_t nodeIndex, u16_t soilIndex, u16_t horizonIndex) { return E(sf3d_set_node_soil(nodeIndex, soilIndex, horizonIndex)); }
SF3Derror_t setNodeSurface(SF3Duint_t nodeIndex, u16_t surfaceIndex) { return E(sf3d_set_node_surface(nodeIndex, surfaceIndex)); }
SF3Derror_t setNodePond(SF3Duint_t nodeIndex, double pond) { return E(sf3d_set_node_pond(nodeIndex, pond)); }
SF3Derror_t setNodeWaterContent(SF3Duint_t nodeIndex, double waterContent) { return E(sf3d_set_node_water_content(nodeIndex, waterContent)); }
SF3Derror_t setNodeDegreeOfSaturation(SF3Duint_t nodeIndex, double degreeOfSaturation) { return E(sf3d_set_node_degree_of_saturation(nodeIndex, degreeOfSaturation)); }
SF3Derror_t setNodeMatricPotential(SF3Duint_t nodeIndex, double matricPotential) { return E(sf3d_set_node_matric_potential(nodeIndex, matricPotential)); }
SF3Derror_t setNodeTotalPotential(SF3Duint_t nodeIndex, double totalPotential) { return E(sf3d_set_node_total_potential(nodeIndex, totalPotential)); }
SF3Derror_t setNodeWaterSinkSource(SF3Duint_t nodeIndex, double waterSinkSource) { return E(sf3d_set_node_water_sink_source(nodeIndex, waterSinkSource)); }
SF3Derror_t setNodePrescribedTotalPotential(SF3Duint_t nodeIndex, double prescribedTotalPotential) { return E(sf3d_set_node_prescribed_total_potential(nodeIndex, prescribedTotalPotential)); }
double getNodeWaterContent(SF3Duint_t nodeIndex) { return sf3d_get_node_water_content(nodeIndex); }
double getNodeMaximumWaterContent(SF3Duint_t nodeIndex) { return sf3d_get_node_maximum_water_content(nodeIndex); }
double getNodeMinimumWaterContent(SF3Duint_t nodeIndex) { return sf3d_get_node_minimum_water_content(nodeIndex); }
double getNodeAvailableWaterContent(SF3Duint_t nodeIndex) { return sf3d_get_node_available_water_content(nodeIndex); }
double getNodeWaterDeficit(SF3Duint_t nodeIndex, double fieldCapacity) { return sf3d_get_node_water_deficit(nodeIndex, fieldCapacity); }
double getNodeDegreeOfSaturation(SF3Duint_t nodeIndex) { return sf3d_get_node_degree_of_saturation(nodeIndex); }
double getNodeWaterConductivity(SF3Duint_t nodeIndex) { return sf3d_get_node_water_conductivity(nodeIndex); }
double getNodeMatricPotential(SF3Duint_t nodeIndex) { return sf3d_get_node_matric_potential(nodeIndex); }
double getNodeTotalPotential(SF3Duint_t nodeIndex) { return sf3d_get_node_total_potential(nodeIndex); }
double getNodePond(SF3Duint_t nodeIndex) { return sf3d_get_node_pond(nodeIndex); }
double getNodeMaxWaterFlow(SF3Duint_t nodeIndex, linkType_t linkDirection) { return sf3d_get_node_max_water_flow(nodeIndex, U(linkDirection)); }
double getNodeSumLateralWaterFlow(SF3Duint_t nodeIndex) { return sf3d_get_node_sum_lateral_water_flow(nodeIndex); }
double getNodeSumLateralWaterFlowIn(SF3Duint_t nodeIndex) { return sf3d_get_node_sum_lateral_water_flow_in(nodeIndex); }
double getNodeSumLateralWaterFlowOut(SF3Duint_t nodeIndex) { return sf3d_get_node_sum_lateral_water_flow_out(nodeIndex); }
double getNodeBoundaryWaterFlow(SF3Duint_t nodeIndex) { return sf3d_get_node_boundary_water_flow(nodeIndex); }
double getTotalBoundaryWaterFlow(boundaryType_t boundaryType) { return sf3d_get_total_boundary_water_flow(U(boundaryType)); }
double getTotalWaterContent() { return sf3d_get_total_water_content(); }
double getWaterStorage() { return sf3d_get_water_storage(); }
double getWaterMBR() { return sf3d_get_water_mbr(); }
SF3Derror_t setNodeHeatSinkSource(SF3Duint_t nodeIndex, double heatSinkSource) { return E(sf3d_set_node_heat_sink_source(nodeIndex, heatSinkSource)); }
SF3Derror_t setNodeTemperature(SF3Duint_t nodeIndex, double temperature) { return E(sf3d_set_node_temperature(nodeIndex, temperature)); }
SF3Derror_t setNodeBoundaryFixedTemperature(SF3Duint_t nodeIndex, double fixedTemperature, double depth) { return E(sf3d_set_node_boundary_fixed_temperature(nodeIndex, fixedTemperature, depth)); }
SF3Derror_t setNodeBoundaryHeightWind(SF3Duint_t nodeIndex, double heightWind) { return E(sf3d_set_node_boundary_height_wind(nodeIndex, heightWind)); }
SF3Derror_t setNodeBoundaryHeightTemperature(SF3Duint_t nodeIndex, double heightTemperature) { return E(sf3d_set_node_boundary_height_temperature(nodeIndex, heightTemperature)); }
SF3Derror_t setNodeBoundaryNetIrradiance(SF3Duint_t nodeIndex, double netIrradiance) { return E(sf3d_set_node_boundary_net_irradiance(nodeIndex, netIrradiance)); }
SF3Derror_t setNodeBoundaryTemperature(SF3Duint_t nodeIndex, double temperature) { return E(sf3d_set_node_boundary_temperature(nodeIndex, temperature)); }
SF3Derror_t setNodeBoundaryRelativeHumidity(SF3Duint_t nodeIndex, double relativeHumidity) { return E(sf3d_set_node_boundary_relative_humidity(nodeIndex, relativeHumidity)); }
SF3Derror_t setNodeBoundaryRoughness(SF3Duint_t nodeIndex, double roughness) { return E(sf3d_set_node_boundary_roughness(nodeIndex, roughness)); }
SF3Derror_t setNodeBoundaryWindSpeed(SF3Duint_t nodeIndex, double windSpeed) { return E(sf3d_set_node_boundary_wind_speed(nodeIndex, windSpeed)); }
double getNodeTemperature(SF3Duint_t nodeIndex) { return sf3d_get_node_temperature(nodeIndex); }
double getNodeHeatConductivity(SF3Duint_t nodeIndex) { return sf3d_get_node_heat_conductivity(nodeIndex); }
double getNodeVapor(SF3Duint_t nodeIndex) { return sf3d_get_node_vapor(nodeIndex); }
double getNodeHeatStorage(SF3Duint_t nodeIndex, double h) { return sf3d_get_node_heat_storage(nodeIndex, h); }
double getNodeHeatMaxFlux(SF3Duint_t nodeIndex, linkType_t linkDirection, fluxTypes_t fluxType) { return sf3d_get_node_heat_max_flux(nodeIndex, U(linkDirection), U(fluxType)); }
double getNodeBoundaryAdvectiveFlux(SF3Duint_t nodeIndex) { return sf3d_get_node_boundary_advective_flux(nodeIndex); }
double getNodeBoundaryLatentFlux(SF3Duint_t nodeIndex) { return sf3d_get_node_boundary_latent_flux(nodeIndex); }
double getNodeBoundaryRadiativeFlux(SF3Duint_t nodeIndex) { return sf3d_get_node_boundary_radiative_flux(nodeIndex); }
double getNodeBoundarySensibleFlux(SF3Duint_t nodeIndex) { return sf3d_get_node_boundary_sensible_flux(nodeIndex); }
double getNodeBoundaryAerodynamicConductance(SF3Duint_t nodeIndex) { return sf3d_get_node_boundary_aerodynamic_conductance(nodeIndex); }
double getNodeBoundarySoilConductance(SF3Duint_t nodeIndex) { return sf3d_get_node_boundary_soil_conductance(nodeIndex); }
double getHeatMBR() { return sf3d_get_heat_mbr(); }
double getHeatMBE() { return sf3d_get_heat_mbe(); }
void computePeriod(double timePeriod) { sf3d_compute_period(timePeriod); }
double computeStep(double maxTimeStep) { return sf3d_compute_step(maxTimeStep); }

}}  // namespace soilFluxes3D::v2
