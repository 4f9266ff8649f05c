"""criteria3d_b200 -- B200-native soilFluxes3D time step (CUDA, sm_100a) behind the
reference's plugin API.  See DESIGN.md.  Python here is harness plumbing only."""
from .capi import (BoundaryType, Field, LinkType, MeanType, SF3Derror, SoilFluxes3D, WRCModel,
                   PRODUCT_LIB, load_product)
