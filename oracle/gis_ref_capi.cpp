/*
 * gis_ref_capi.cpp -- TEST INFRASTRUCTURE (oracle side only): the reference's own raster helpers behind a C call.
 *
 * Compiled by oracle/Makefile into oracle/_ref/libgis_ref.so together with the UNMODIFIED reference sources
 * agrolib/gis/{gis,gisIO,color}.cpp, agrolib/mathFunctions/{basicMath,statistics,furtherMathFunctions,physics}.cpp and
 * agrolib/crit3dDate/{crit3dDate,crit3dTime}.cpp (Qt-free), where they lie under /root/reference.  This file
 * only builds the grids and forwards to
 *   gis::computeSlopeAspectMaps      (gis.cpp:1190-1268; boundary cells: computeSlopeAspectBoundary :1114-1186)
 *   gis::isBoundaryRunoff            (gis.cpp:1452-1488) over the surface index map, as Project3D::setLateralBoundary
 *                                    (src/project3D/project3D.cpp:851-873) does
 * and evaluates  boundarySlope = tan(slopeDegree * DEG_TO_RAD)  as Project3D::setCrit3DTopography (:964-965).
 * It is the checker of criteria3d_b200/raster.py; nothing of the product links it.
 */
#include <cmath>
#include <cstdint>

#include "commonConstants.h"
#include "gis.h"

extern "C" int gisref_slope_aspect_boundary(int rows, int cols, double cell, float flag, const float *dem,
                                            float *slopeDeg, float *aspectDeg, uint8_t *boundaryRunoff, float *boundarySlopeTan)
{
    gis::Crit3DRasterHeader header;
    header.nrRows = rows; header.nrCols = cols; header.cellSize = cell; header.flag = flag;
    header.llCorner.x = 0.; header.llCorner.y = 0.;

    gis::Crit3DRasterGrid DEM;
    if (!DEM.initializeGrid(header)) return 1;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) DEM.value[r][c] = dem[(size_t)r * cols + c];
    DEM.isLoaded = true;

    gis::Crit3DRasterGrid slopeMap, aspectMap;
    if (!gis::computeSlopeAspectMaps(DEM, &slopeMap, &aspectMap)) return 2;

    /* surface index map of Project3D::setIndexMaps (project3D.cpp:758-818) with every DEM cell in a land unit */
    gis::Crit3DIndexGrid indexMap;
    indexMap.initializeGrid(*(DEM.header));
    const long noIndex = static_cast<long>(indexMap.header->flag);
    long current = 0;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c)
            indexMap.value[r][c] = (DEM.value[r][c] == DEM.header->flag) ? noIndex : current++;

    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c)
        {
            const size_t k = (size_t)r * cols + c;
            slopeDeg[k] = slopeMap.value[r][c];
            aspectDeg[k] = aspectMap.value[r][c];
            boundaryRunoff[k] = gis::isBoundaryRunoff(indexMap, DEM, aspectMap, r, c) ? 1 : 0;
            const float slopeDegree = slopeMap.value[r][c];
            const float boundarySlope = tan(slopeDegree * DEG_TO_RAD);
            boundarySlopeTan[k] = boundarySlope;
        }
    return 0;
}

/* ESRI float grid I/O of the reference (agrolib/gis/gisIO.cpp: readEsriGridFlt :1587, writeEsriGrid :1575),
 * fileName WITHOUT extension as the reference passes it.  Checker of criteria3d_b200/raster.py read_flt / write_flt. */
extern "C" int gisref_read_flt(const char *fileNameNoExt, int *rows, int *cols, double *cell, double *xll, double *yll,
                               float *flag, float *values, long capacity)
{
    gis::Crit3DRasterGrid grid;
    std::string error;
    if (!gis::readEsriGridFlt(fileNameNoExt, &grid, error)) return 1;
    *rows = grid.header->nrRows; *cols = grid.header->nrCols; *cell = grid.header->cellSize;
    *xll = grid.header->llCorner.x; *yll = grid.header->llCorner.y; *flag = grid.header->flag;
    if ((long)grid.header->nrRows * grid.header->nrCols > capacity) return 2;
    for (int r = 0; r < grid.header->nrRows; ++r)
        for (int c = 0; c < grid.header->nrCols; ++c) values[(size_t)r * grid.header->nrCols + c] = grid.value[r][c];
    return 0;
}

extern "C" int gisref_write_flt(const char *fileNameNoExt, int rows, int cols, double cell, double xll, double yll,
                                float flag, const float *values)
{
    gis::Crit3DRasterHeader header;
    header.nrRows = rows; header.nrCols = cols; header.cellSize = cell; header.flag = flag;
    header.llCorner.x = xll; header.llCorner.y = yll;
    gis::Crit3DRasterGrid grid;
    if (!grid.initializeGrid(header)) return 1;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) grid.value[r][c] = values[(size_t)r * cols + c];
    grid.isLoaded = true;
    std::string error;
    return gis::writeEsriGrid(fileNameNoExt, &grid, error) ? 0 : 3;
}
