// dropin_caller.cpp -- a caller of the reference's C++ plugin API, written against include/soilFluxes3D.h
// only, in the call order Project3D uses (src/project3D/project3D.cpp: initializeSF3D :558, soil and surface
// tables :1164-1238, setNode / setNodeLink :941-1103, setHydraulicProperties / setNumericalParameters
// :573-590, initial state :1106-1160, initializeBalance, then per hour setNodeWaterSinkSource + the
// computeStep loop :1307-1386 and the getters).  The SAME object file links against the product
// (libsf3d_b200.so) and against the unmodified reference built behind oracle/_ref (both export the
// soilFluxes3D::v2::* symbols): tests/test_cpp_dropin.py swaps the library and compares the output.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "soilFluxes3D.h"

using namespace soilFluxes3D;

static void check(SF3Derror_t rc, const char *what)
{
    std::string name;
    if (getSF3DerrorName(rc, name)) { std::fprintf(stderr, "%s: %s\n", what, name.c_str()); std::exit(2); }
}

int main()
{
    // a 5 x 4 raster, one surface layer + 3 soil layers, 10 m cells on a tilted plane, outlet on the last row
    const int R = 5, C = 4, L = 4;
    const double cell = 10., area = cell * cell;
    const double depth[L] = {0., 0.05, 0.15, 0.30}, thick[L] = {0., 0.10, 0.10, 0.20};
    const SF3Duint_t nCells = R * C, n = nCells * L;
    auto id = [&](int l, int r, int c) { return SF3Duint_t(l * nCells + r * C + c); };

    check(initializeSF3D(n, nCells, 8, true, false, false), "initializeSF3D");
    setThreadsNumber(1);
    check(setSoilProperties(0, 0, 3.6, 1.56, 1. - 1. / 1.56, 0.02, 0.078, 0.43, 2.9e-6, 0.5, 0.02, 0.2), "setSoilProperties");
    check(setSoilProperties(0, 1, 1.9, 1.31, 1. - 1. / 1.31, 0.03, 0.095, 0.41, 7.2e-7, 0.5, 0.01, 0.3), "setSoilProperties");
    check(setSurfaceProperties(0, 0.24), "setSurfaceProperties");

    for (int l = 0; l < L; ++l)
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C; ++c)
            {
                const double z = 100. + 0.4 * (R - 1 - r) + 0.1 * c - depth[l];
                const bool surface = (l == 0), outlet = (r == R - 1);
                boundaryType_t bt = boundaryType_t::NoBoundary;
                double slope = 0., bArea = 0.;
                if (surface && outlet) { bt = boundaryType_t::Runoff; slope = 0.04; bArea = cell; }
                else if (l == L - 1) { bt = boundaryType_t::FreeDrainage; bArea = area; }
                else if (!surface && outlet) { bt = boundaryType_t::FreeLateralDrainage; slope = 0.04; bArea = cell * thick[l]; }
                check(setNode(id(l, r, c), cell * (c + 0.5), cell * (R - r - 0.5), z, surface ? area : area * thick[l], surface, bt, slope, bArea), "setNode");
            }
    for (int l = 0; l < L; ++l)
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C; ++c)
            {
                const SF3Duint_t i = id(l, r, c);
                if (l > 0) check(setNodeLink(i, id(l - 1, r, c), linkType_t::Up, area), "setNodeLink up");
                if (l < L - 1) check(setNodeLink(i, id(l + 1, r, c), linkType_t::Down, area), "setNodeLink down");
                for (int dr = -1; dr <= 1; ++dr)
                    for (int dc = -1; dc <= 1; ++dc)
                    {
                        if ((dr == 0 && dc == 0) || r + dr < 0 || r + dr >= R || c + dc < 0 || c + dc >= C) continue;
                        const double lateral = (l == 0) ? cell : cell * thick[l];
                        check(setNodeLink(i, id(l, r + dr, c + dc), linkType_t::Lateral, lateral * 0.5), "setNodeLink lateral");
                    }
                if (l == 0) { check(setNodeSurface(i, 0), "setNodeSurface"); check(setNodePond(i, 0.002), "setNodePond"); }
                else check(setNodeSoil(i, 0, l < 2 ? 0 : 1), "setNodeSoil");
            }

    check(setHydraulicProperties(WRCModel::ModifiedVanGenuchten, meanType_t::Logarithmic, 10.f), "setHydraulicProperties");
    check(setNumericalParameters(1., 600., 200, 10, 12, 3), "setNumericalParameters");
    for (SF3Duint_t i = 0; i < n; ++i) check(setNodeMatricPotential(i, i < nCells ? 0. : -2.0), "setNodeMatricPotential");
    check(initializeBalance(), "initializeBalance");

    const double rainMmH[2] = {30., 8.};
    int steps = 0;
    for (double mm : rainMmH)
    {
        for (SF3Duint_t i = 0; i < n; ++i)
            check(setNodeWaterSinkSource(i, i < nCells ? area * mm / 1000. / 3600. : 0.), "setNodeWaterSinkSource");
        for (double t = 0.; t < 1800.;) { t += computeStep(1800. - t); ++steps; }
    }
    std::printf("steps %d\n", steps);
    std::printf("runoff %.12e drainage %.12e lateral %.12e\n", getTotalBoundaryWaterFlow(boundaryType_t::Runoff),
                getTotalBoundaryWaterFlow(boundaryType_t::FreeDrainage), getTotalBoundaryWaterFlow(boundaryType_t::FreeLateralDrainage));
    for (SF3Duint_t i = 0; i < n; i += 7)
        std::printf("%u H %.12e theta %.12e Se %.12e K %.12e\n", i, getNodeTotalPotential(i), getNodeWaterContent(i),
                    getNodeDegreeOfSaturation(i), getNodeWaterConductivity(i));
    std::printf("index_error %.1f\n", getNodeWaterContent(n + 3));
    check(cleanSF3D(), "cleanSF3D");
    return 0;
}
