"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference (mattevs24/reyna v1.3.0 at
/root/reference) in the build container.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

The reference ships no golden vectors (its tests draw a fresh random mesh each run, tests/test_DGFEM.py:12), so these
files ARE the pin: each holds one reference mesh + DGFEMGeometry (flat arrays) and, per (problem, p) case, the
reference's B (CSR triplet), L, spsolve solution and errors().  The coefficient callables are referenced by name
(reyna_b200/problems.py); nothing from /root/reference is copied, only its outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from reyna_b200.problems import PROBLEMS  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

MESHES = {
    # name: (domain factory name, domain args, poly_mesher kwargs)
    "square64": ("RectangleDomain", dict(bounding_box=np.array([[0.0, 1.0], [0.0, 1.0]])),
                 dict(max_iterations=10, n_points=64, cleaned=True, seed=1337)),
    "circle48": ("CircleDomain", dict(),
                 dict(max_iterations=6, n_points=48, cleaned=True, seed=7)),
    "rect24": ("RectangleDomain", dict(bounding_box=np.array([[0.0, 0.25], [0.0, 0.25]])),
               dict(max_iterations=3, n_points=24, cleaned=False, seed=11)),
}

CASES = {
    "square64": [("adr", 1), ("adr", 2), ("adr", 3), ("diffusion", 1), ("diffusion", 2), ("hyperbolic", 1),
                 ("hyperbolic", 2), ("adr_tensor", 3), ("variable", 2), ("variable_hyperbolic", 2)],
    "circle48": [("adr", 2), ("variable", 1), ("variable", 3), ("variable_hyperbolic", 1), ("hyperbolic", 3)],
    "rect24": [("adr", 2), ("variable", 2)],
}


def geometry_arrays(geo) -> dict:
    regions = [np.asarray(r, dtype=np.int64) for r in geo.mesh.filtered_regions]
    off = np.zeros(len(regions) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in regions], out=off[1:])
    return dict(
        vertices=np.asarray(geo.mesh.vertices, dtype=np.float64),
        region_offsets=off, region_indices=np.concatenate(regions),
        filtered_points=np.asarray(geo.mesh.filtered_points, dtype=np.float64),
        nodes=np.asarray(geo.nodes, dtype=np.float64),
        subtriangulation=np.asarray(geo.subtriangulation, dtype=np.int64),
        triangle_to_polygon=np.asarray(geo.triangle_to_polygon, dtype=np.int64),
        elem_bounding_boxes=np.asarray(geo.elem_bounding_boxes, dtype=np.float64),
        interior_edges=np.asarray(geo.interior_edges, dtype=np.int64),
        interior_edges_to_element=np.asarray(geo.interior_edges_to_element, dtype=np.int64),
        interior_normals=np.asarray(geo.interior_normals, dtype=np.float64),
        boundary_edges=np.asarray(geo.boundary_edges, dtype=np.int64),
        boundary_edges_to_element=np.asarray(geo.boundary_edges_to_element, dtype=np.int64),
        boundary_normals=np.asarray(geo.boundary_normals, dtype=np.float64),
        areas=np.asarray(geo.areas, dtype=np.float64), h_s=np.asarray(geo.h_s, dtype=np.float64),
        h=np.float64(geo.h), n_elements=np.int64(geo.n_elements), n_triangles=np.int64(geo.n_triangles),
    )


def main() -> None:
    ref = ref_harness.load()
    os.makedirs(GOLDEN, exist_ok=True)
    for mname, (dname, dargs, mkw) in MESHES.items():
        np.random.seed(1142)                     # tests/test_DGFEM.py:12
        dom = getattr(ref.domains, dname)(**dargs)
        mesh = ref.poly_mesher(dom, **mkw)
        geo = ref.DGFEMGeometry(mesh)
        out = {"g_" + k: v for k, v in geometry_arrays(geo).items()}
        names = []
        for pname, p in CASES[mname]:
            pr = PROBLEMS[pname]
            dg = ref.DGFEM(geo, polynomial_degree=p)
            dg.add_data(**pr["data"])
            dg.dgfem(solve=True)
            key = f"{pname}_p{p}"
            names.append(key)
            out[key + "_indptr"] = dg.B.indptr.copy()
            out[key + "_indices"] = dg.B.indices.copy()
            out[key + "_data"] = dg.B.data.copy()
            out[key + "_L"] = np.asarray(dg.L.todense()).reshape(-1)
            out[key + "_solution"] = np.asarray(dg.solution, dtype=np.float64)
            bi = dg.boundary_information
            out[key + "_elliptic"] = (np.asarray(bi.elliptical_indecies, dtype=np.int64)
                                      if bi.elliptical_indecies is not None else np.zeros(0, np.int64) - 1)
            out[key + "_inflow"] = (np.asarray(bi.inflow_indecies, dtype=np.int64)
                                    if bi.inflow_indecies is not None else np.zeros(0, np.int64) - 1)
            if pr["exact"] is not None:
                has_diff = pr["data"].get("diffusion") is not None
                l2, dgn, h1 = dg.errors(exact_solution=pr["exact"],
                                        grad_exact_solution=pr["grad_exact"] if has_diff else None)
                out[key + "_errors"] = np.array([l2, dgn, np.nan if h1 is None else h1], dtype=np.float64)
            print(f"{mname}: {key}: N={dg.dim_system} nnz={dg.B.nnz}")
        out["cases"] = np.array(names)
        path = os.path.join(GOLDEN, mname + ".npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
