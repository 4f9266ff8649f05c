// sf3d_gis.cu -- raster preparation on the device (C ABI in include/sf3d_gis.h): slope / aspect maps of a DEM,
// the runoff-boundary mask and tan(slope), one thread per cell.  Arithmetic follows agrolib/gis/gis.cpp expression by
// expression (float where the reference computes in float, double where it computes in double; the file is compiled with
// -fmad=false like the rest of the product), so the float maps agree with the reference's bit for bit wherever the
// device's atan / atan2 / tan round to the same float as glibc's (every cell of the test rasters).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include "sf3d_backend.h"
#include "sf3d_gis.h"

namespace {

constexpr double kEpsilon = 0.00001;            // mathFunctions/commonConstants.h:252
constexpr double kRadToDeg = 57.295779513;      // :256
constexpr double kDegToRad = 0.01745329252;     // :255

struct GisGrid { int rows, cols; double cell; float flag; const float *dem; };

// isEqual(float, float), basicMath.h:25-26
__device__ __forceinline__ bool is_flag(float a, float flag) { return fabs((double)a - (double)flag) < kEpsilon; }
// Crit3DRasterGrid::getValueFromRowCol (gis.cpp:521-530): the flag outside the grid
__device__ __forceinline__ float cell_value(const GisGrid &g, int r, int c)
{ return (r < 0 || r >= g.rows || c < 0 || c >= g.cols) ? g.flag : g.dem[(size_t)r * g.cols + c]; }

__global__ void __launch_bounds__(256) kern_gis_prepare(GisGrid g, float *__restrict__ slopeDeg, float *__restrict__ aspectDeg,
                                                        uint8_t *__restrict__ boundaryRunoff, float *__restrict__ boundarySlopeTan)
{
    const size_t cells = (size_t)g.rows * g.cols;
    for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < cells; k += (size_t)gridDim.x * 256)
    {
        const int row = (int)(k / g.cols), col = (int)(k - (size_t)row * g.cols);
        float nb[3][3];
        bool missing[3][3];
        bool anyMissing = false;
        #pragma unroll
        for (int i = -1; i <= 1; ++i)
            #pragma unroll
            for (int j = -1; j <= 1; ++j)
            {
                const float zz = cell_value(g, row + i, col + j);
                nb[i + 1][j + 1] = zz;
                missing[i + 1][j + 1] = is_flag(zz, g.flag);
                if (i != 0 || j != 0) anyMissing = anyMissing || missing[i + 1][j + 1];
            }
        const float z = nb[1][1];
        const bool valid = !missing[1][1];
        const bool rim = valid && anyMissing;                       // gis::isBoundary, gis.cpp:1494-1510
        float slope = g.flag, aspect = g.flag;
        if (valid && rim)
        {
            // computeSlopeAspectBoundary, gis.cpp:1126-1180: one-sided sums over the valid neighbours; i * (z - z1) is a float product
            double dz = 0., dl = 0.;
            #pragma unroll
            for (int i = -1; i <= 1; i += 2)
                #pragma unroll
                for (int j = -1; j <= 1; ++j)
                    if (!missing[i + 1][j + 1]) { dz += (double)((float)i * (z - nb[i + 1][j + 1])); dl += g.cell; }
            const double dz_dy = dz / fmax(dl, kEpsilon);
            dz = 0.; dl = 0.;
            #pragma unroll
            for (int j = -1; j <= 1; j += 2)
                #pragma unroll
                for (int i = -1; i <= 1; ++i)
                    if (!missing[i + 1][j + 1]) { dz += (double)((float)j * (z - nb[i + 1][j + 1])); dl += g.cell; }
            const double dz_dx = dz / fmax(dl, kEpsilon);
            slope = (float)(atan(sqrt(dz_dx * dz_dx + dz_dy * dz_dy)) * kRadToDeg);
            double a = atan2(-dz_dy, dz_dx);
            a = 90.0 - a * kRadToDeg;
            if (a < 0) a += 360;
            aspect = (float)a;
        }
        else if (valid)
        {
            // Horn's 3x3 derivatives, gis.cpp:1219-1254
            const double z1 = nb[0][0], z2 = nb[0][1], z3 = nb[0][2], z4 = nb[1][0], z6 = nb[1][2], z7 = nb[2][0], z8 = nb[2][1], z9 = nb[2][2];
            const double dzdx = ((z3 + 2 * z6 + z9) - (z1 + 2 * z4 + z7)) / (8.0 * g.cell);
            const double dzdy = ((z7 + 2 * z8 + z9) - (z1 + 2 * z2 + z3)) / (8.0 * g.cell);
            if (fabs(dzdx) < kEpsilon && fabs(dzdy) < kEpsilon) { slope = 0.f; aspect = 0.f; }
            else
            {
                const double slopeRad = atan(sqrt(dzdx * dzdx + dzdy * dzdy));
                slope = (float)(slopeRad * kRadToDeg);
                double a = atan2(dzdy, -dzdx);
                a = 90.0 - a * kRadToDeg;
                if (a < 0) a += 360.0;
                aspect = (float)a;
            }
        }
        if (slopeDeg) slopeDeg[k] = slope;
        if (aspectDeg) aspectDeg[k] = aspect;
        if (boundarySlopeTan) boundarySlopeTan[k] = (float)tan((double)slope * kDegToRad);      // project3D.cpp:964-965
        if (boundaryRunoff)
        {
            // gis::isBoundaryRunoff, gis.cpp:1452-1488 (index map: every valid DEM cell is a node)
            bool out = false;
            if (rim)
            {
                bool strictMin = true;                               // isMinimum(dtm, true, ...), gis.cpp:1395-1427
                #pragma unroll
                for (int i = 0; i < 3; ++i)
                    #pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if ((i != 1 || j != 1) && !missing[i][j] && z >= nb[i][j]) strictMin = false;
                if (strictMin) out = true;
                else if (!is_flag(aspect, g.flag))
                {
                    int r = 0, c = 0;
                    if (aspect >= 135 && aspect <= 225) r = 1;
                    else if ((aspect <= 45) || (aspect >= 315)) r = -1;
                    if (aspect >= 45 && aspect <= 135) c = 1;
                    else if (aspect >= 225 && aspect <= 315) c = -1;
                    out = missing[1 + r][1 + c];
                }
            }
            boundaryRunoff[k] = out ? 1 : 0;
        }
    }
}

}  // namespace

extern "C" uint8_t sf3d_gis_slope_aspect_boundary(uint32_t rows, uint32_t cols, double cell_size, float flag, const float *dem,
                                                  float *slope_deg, float *aspect_deg, uint8_t *boundary_runoff,
                                                  float *boundary_slope_tan)
{
    using namespace sf3d;
    if (!dem || rows == 0 || cols == 0 || rows > 0x7FFFFFFFu || cols > 0x7FFFFFFFu) return SF3D_PARAMETER_ERROR;
    const size_t cells = (size_t)rows * cols;
    float *dDem = nullptr, *dSlope = nullptr, *dAspect = nullptr, *dTan = nullptr;
    uint8_t *dMask = nullptr;
    uint8_t rc = SF3D_OK;
    try
    {
        dDem = (float *)dev_alloc(cells * sizeof(float));
        h2d(dDem, dem, cells * sizeof(float));
        if (slope_deg) dSlope = (float *)dev_alloc(cells * sizeof(float));
        if (aspect_deg) dAspect = (float *)dev_alloc(cells * sizeof(float));
        if (boundary_slope_tan) dTan = (float *)dev_alloc(cells * sizeof(float));
        if (boundary_runoff) dMask = (uint8_t *)dev_alloc(cells);
        GisGrid g{(int)rows, (int)cols, cell_size, flag, dDem};
        const size_t want = (cells + 255) / 256;
        const int grid = (int)(want < (size_t)148 * 8 ? want : (size_t)148 * 8);
        kern_gis_prepare<<<grid, 256, 0, (cudaStream_t)dev_stream()>>>(g, dSlope, dAspect, dMask, dTan);
        if (cudaGetLastError() != cudaSuccess) throw DeviceError{-1, "launch failed", "kern_gis_prepare"};
        if (slope_deg) d2h(slope_deg, dSlope, cells * sizeof(float));
        if (aspect_deg) d2h(aspect_deg, dAspect, cells * sizeof(float));
        if (boundary_slope_tan) d2h(boundary_slope_tan, dTan, cells * sizeof(float));
        if (boundary_runoff) d2h(boundary_runoff, dMask, cells);
    }
    catch (const DeviceError &e)
    {
        fprintf(stderr, "[sf3d_b200] %s: %s\n", e.where ? e.where : "sf3d_gis", e.what ? e.what : "device error");
        rc = SF3D_MEMORY_ERROR;
    }
    dev_free(dDem); dev_free(dSlope); dev_free(dAspect); dev_free(dTan); dev_free(dMask);
    return rc;
}
