"""reyna_b200 -- B200-native DGFEM assemble-and-solve path behind reyna's Python API.

Drop-in surface (mattevs24/reyna): ``DGFEMGeometry(poly_mesh)``, ``DGFEM(geometry, polynomial_degree)``,
``DGFEM.add_data(...)``, ``DGFEM.dgfem(solve=...)``, ``DGFEM.errors(...)``.  The compute path is
``libreyna_b200.so`` (hand-written sm_100a CUDA behind the C-ABI of ``include/reyna_b200.h``); there is no CPU fallback.
"""
from .mesh import PolyMesh, RectangleDomain, FlatRegions, voronoi_mesh, tiled_voronoi_mesh, morton_order
from .geometry import DGFEMGeometry, flat_geometry, save_geometry_cache, load_geometry_cache
from .dgfem import DGFEM
from .boundary_information import BoundaryInformation
from .tools import evaluate, evaluate_batch, plot_data, plot_DG
from ._lib import ReynaB200Error, build, lib, LIB_PATH

__version__ = "0.1.0"
__all__ = ["PolyMesh", "RectangleDomain", "FlatRegions", "voronoi_mesh", "tiled_voronoi_mesh", "morton_order",
           "DGFEMGeometry", "flat_geometry", "save_geometry_cache", "load_geometry_cache", "DGFEM", "BoundaryInformation", "evaluate", "evaluate_batch", "plot_data", "plot_DG", "ReynaB200Error", "build", "lib", "LIB_PATH"]
