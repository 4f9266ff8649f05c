/*
 * sf3d_gis.h -- C ABI of the raster preparation that precedes the soilFluxes3D set-up: what
 * CRITERIA-3D's Project3D does with the DEM before it creates the nodes.
 *
 *   slope / aspect maps   gis::computeSlopeAspectMaps          /root/reference/agrolib/gis/gis.cpp:1190-1268
 *                         (cells with a missing neighbour: computeSlopeAspectBoundary, gis.cpp:1114-1186)
 *   runoff boundary       gis::isBoundaryRunoff over the surface index map (gis.cpp:1452-1488, isMinimum :1395-1427,
 *                         isBoundary :1494-1510), i.e. Project3D::setLateralBoundary (src/project3D/project3D.cpp:851-873)
 *   boundary slope        float boundarySlope = tan(slopeDegree * DEG_TO_RAD), Project3D::setCrit3DTopography
 *                         (project3D.cpp:964-965)
 *
 * One call for the whole raster; plain pointers to HOST buffers, no torch types.  The product computes the maps on the
 * GPU (criteria3d_b200/csrc/sf3d_gis.cu); the argument list is the one of the reference-side checker
 * oracle/gis_ref_capi.cpp (gisref_slope_aspect_boundary), which forwards to the reference's own functions.
 */
#ifndef SF3D_GIS_H
#define SF3D_GIS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* dem: rows*cols floats, row 0 = northernmost, `flag` = NODATA value of the header.  Outputs (each may be NULL):
 *   slope_deg, aspect_deg   rows*cols floats, `flag` where the DEM is NODATA (aspect: 0 = north, clockwise)
 *   boundary_runoff         rows*cols bytes, 1 where the cell gets a Runoff boundary
 *   boundary_slope_tan      rows*cols floats, tan of the slope (the reference evaluates it for every cell)
 * Returns SF3D_OK (0), SF3D_PARAMETER_ERROR (6) for an empty raster or a NULL dem, SF3D_MEMORY_ERROR (2) when no CUDA
 * device is available or an allocation fails (there is no CPU fallback). */
uint8_t sf3d_gis_slope_aspect_boundary(uint32_t rows, uint32_t cols, double cell_size, float flag, const float *dem,
                                       float *slope_deg, float *aspect_deg, uint8_t *boundary_runoff,
                                       float *boundary_slope_tan);

#ifdef __cplusplus
}
#endif
#endif
