// sf3d_fields.h -- per-node values of the bulk getters and the raster-facing forcing / output rows
// (SURVEY 8 f3, f4), as __host__ __device__ functions shared by the kernels.
#pragma once
#include "sf3d.h"
#include "sf3d_backend.h"
#include "sf3d_rows.h"

// same value as the scalar getter of the node (soilFluxes3D.cpp:951-1234)
SF3D_HD double sf3d_field_value(const SF3DView &v, int field, uint32_t i)
{
    const size_t N = v.N;
    const uint32_t m = v.meta[i];
    const bool surface = META_SURFACE(m);
    switch (field)
    {
        case SF3D_F_WATER_CONTENT:
            return surface ? (v.H[i] - v.z[i]) : sf3d_theta_from_se(v.soil[v.tab[i]], v.Se[i]);
        case SF3D_F_DEGREE_OF_SATURATION:
        {
            if (!surface) return v.Se[i];
            const double curPot = v.H[i] - v.z[i], maxPot = 0.001;
            return curPot <= 0 ? 0 : (curPot > maxPot ? 1. : curPot / maxPot);
        }
        case SF3D_F_WATER_CONDUCTIVITY: return v.K[i];
        case SF3D_F_MATRIC_POTENTIAL:   return v.H[i] - v.z[i];
        case SF3D_F_TOTAL_POTENTIAL:    return v.H[i];
        case SF3D_F_POND:               return surface ? v.pond[i] : -1111.;
        case SF3D_F_BOUNDARY_WATER_FLOW: return (META_BT(m) == BT_NONE) ? -4444. : v.bSum[i];
        case SF3D_F_SUM_LATERAL_FLOW:
        {
            double s = 0.;
            for (uint32_t l = 0; l < META_NLAT(m); ++l) s += v.lflow[(size_t)(2 + l) * N + i];
            return s;
        }
        case SF3D_F_MAX_FLOW_UP:   return v.lflow[i];
        case SF3D_F_MAX_FLOW_DOWN: return v.lflow[N + i];
        case SF3D_F_MAX_FLOW_LATERAL:
        {
            double mx = 0.;
            for (uint32_t l = 0; l < META_NLAT(m); ++l) mx = sf3d_max(mx, v.lflow[(size_t)(2 + l) * N + i]);
            return mx;
        }
        case SF3D_F_TEMPERATURE: return (v.computeHeat && !surface) ? v.T[i] : -3333.;
        default: return -1111.;
    }
}

// hourly forcing of one raster cell (criteria3DProject.cpp:2121-2160; see include/sf3d.h)
SF3D_HD void sf3d_cell_forcing(const SF3DView &v, const sf3d::RasterDev &g, const sf3d::ForcingDev &f, uint64_t cell)
{
    const int32_t rank = g.rank[cell];
    if (rank < 0) return;
    const uint64_t cells = (uint64_t)g.rows * g.cols;
    const double area = g.cell * g.cell;
    for (uint32_t layer = 0; layer < g.layers; ++layer)
    {
        const uint32_t i = layer * g.nValid + (uint32_t)rank;
        double q = f.accumulate ? v.sink[i] : 0.;
        if (f.layerSink && layer < f.nSinkLayers)
        {
            const float s = f.layerSink[(uint64_t)layer * cells + cell];
            if (s != f.sinkNodata && s > 0.f) q -= area * (s / 1000.) / 3600.;
        }
        if (layer == 0 && f.precipitation)
        {
            const float p = f.precipitation[cell];
            if (p != f.precipitationNodata && p > 0.f)
            {
                const double flow = area * (p / 1000.);               // [m3 h-1]
                if ((flow / 3600.) > 0.) q += flow / 3600.;          // [m3 s-1]
            }
        }
        v.sink[i] = q;
    }
}

// one cell of an output map (Project3D::computeCriteria3DMap, project3D.cpp:1917-1944)
SF3D_HD float sf3d_cell_output(const SF3DView &v, const sf3d::RasterDev &g, int field, uint32_t layer, float nodata, uint64_t cell)
{
    const int32_t rank = g.rank[cell];
    if (rank < 0) return nodata;
    double value = sf3d_field_value(v, field, layer * g.nValid + (uint32_t)rank);
    if (value == SF3D_NODATA) return nodata;
    if (field == SF3D_F_WATER_CONTENT && layer == 0) value *= 1000;          // [m] -> [mm]
    return (float)value;
}
