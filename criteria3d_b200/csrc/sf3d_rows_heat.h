// sf3d_rows_heat.h -- heat-coupling hooks of the water rows and the rows of the heat update.
// (Filled in by the heat milestone; until then the hooks are neutral and the C ABI refuses
//  isComputeHeat = true with SF3D_PARAMETER_ERROR, so nothing silently runs without them.)
#pragma once
#include "sf3d_rows.h"

SF3D_HD double sf3d_heat_vapor_K(const SF3DView &, uint32_t) { return 0.; }
SF3D_HD double sf3d_heat_dthetav_dh(const SF3DView &, uint32_t, double) { return 0.; }
SF3D_HD double sf3d_heat_surface_boundary(const SF3DView &, uint32_t, double, double *) { return 0.; }
SF3D_HD double sf3d_heat_surface_pull(const SF3DView &, uint32_t, double, int *) { return 0.; }
SF3D_HD double sf3d_heat_thermal_invariant(const SF3DView &, uint32_t, int, uint32_t) { return 0.; }
