"""Shared helpers of the test-suite: golden fixtures -> geometry objects, CSR comparison in the parity metric."""
from __future__ import annotations

import os
import types

import numpy as np
from scipy.sparse import csr_matrix

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north-star tolerances
TOL_ENTRY = 1e-12      # stiffness / load entries, relative to the d x d block (or load vector) max
TOL_SOLUTION = 1e-9    # relative L2 against spsolve
TOL_ERRORS = 5e-4      # "3 significant digits"


class GoldenGeometry:
    """Object with the attribute names of the reference's DGFEMGeometry, rebuilt from a fixture (lists where the
    reference holds lists, so that the flattening code is exercised the same way)."""

    def __init__(self, z):
        off, idx = z["g_region_offsets"], z["g_region_indices"]
        regions = [idx[off[i]:off[i + 1]].copy() for i in range(off.shape[0] - 1)]
        self.mesh = types.SimpleNamespace(vertices=z["g_nodes"].copy(), filtered_regions=regions,
                                          filtered_points=z["g_filtered_points"].copy(), domain=None)
        self.nodes = self.mesh.vertices
        self.n_nodes = self.nodes.shape[0]
        self.n_elements = int(z["g_n_elements"])
        self.n_triangles = int(z["g_n_triangles"])
        self.subtriangulation = z["g_subtriangulation"].copy()
        self.triangle_to_polygon = list(z["g_triangle_to_polygon"])
        self.elem_bounding_boxes = [list(r) for r in z["g_elem_bounding_boxes"]]
        self.interior_edges = z["g_interior_edges"].astype(np.int32)
        self.interior_edges_to_element = z["g_interior_edges_to_element"].copy()
        self.interior_normals = z["g_interior_normals"].copy()
        self.boundary_edges = z["g_boundary_edges"].astype(np.int32)
        self.boundary_edges_to_element = z["g_boundary_edges_to_element"].copy()
        self.boundary_normals = z["g_boundary_normals"].copy()
        self.areas = z["g_areas"].copy()
        self.h_s = z["g_h_s"].copy()
        self.h = float(z["g_h"])


class ParityNote(UserWarning):
    """Carrier of per-case parity figures into pytest's warnings summary (visible in the -q log of a passing run)."""


def note(msg: str) -> None:
    import warnings
    warnings.warn(msg, ParityNote, stacklevel=2)


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, GoldenGeometry(z)


def golden_cases():
    out = []
    for m in golden_names():
        z = np.load(os.path.join(GOLDEN, m + ".npz"))
        for c in z["cases"]:
            out.append((m, str(c)))
    return out


def golden_system(z, key):
    n = z[key + "_L"].shape[0]
    B = csr_matrix((z[key + "_data"], z[key + "_indices"], z[key + "_indptr"]), shape=(n, n))
    return B, z[key + "_L"], z[key + "_solution"]


def split_case(key):
    name, p = key.rsplit("_p", 1)
    return name, int(p)


def compare_csr(B, Bref, d, tol=TOL_ENTRY):
    """Return (max block-relative value error, number of pattern mismatches that are NOT rounding noise,
    number of noise-level pattern mismatches).

    An entry is 'noise' when its magnitude is <= tol * max|block| in the matrix that stores it: the reference's own
    pattern keeps or drops such entries depending on the rounding of a mathematically zero sum (scipy prunes exact zeros
    only), so no independent implementation can reproduce them -- see DESIGN.md 'pattern parity'."""
    B = csr_matrix(B)
    Bref = csr_matrix(Bref)
    assert B.shape == Bref.shape
    nb = max(B.shape) // d + 1          # (row block, column block) -> one key; rectangular row slices allowed

    def blockmax(M):
        C = M.tocoo()
        key = (C.row // d).astype(np.int64) * nb + (C.col // d)
        uk, inv = np.unique(key, return_inverse=True)
        mx = np.zeros(uk.shape[0])
        np.maximum.at(mx, inv, np.abs(C.data))
        return dict(zip(uk.tolist(), mx.tolist()))

    bm = blockmax(Bref)
    bm2 = blockmax(B)
    D = (B - Bref).tocoo()
    key = (D.row // d).astype(np.int64) * nb + (D.col // d)
    scale = np.array([max(bm.get(k, 0.0), bm2.get(k, 0.0)) for k in key.tolist()])
    rel = np.abs(D.data) / np.where(scale > 0, scale, 1.0)
    max_rel = float(rel.max()) if rel.size else 0.0

    P, Pref = B.copy(), Bref.copy()
    P.data = np.ones_like(P.data)
    Pref.data = np.ones_like(Pref.data)
    X = (P - Pref).tocoo()
    bad = noise = 0
    for r, c, v in zip(X.row, X.col, X.data):
        if v == 0:
            continue
        k = int(r // d) * nb + int(c // d)
        val = abs(B[r, c]) if v > 0 else abs(Bref[r, c])
        s = max(bm.get(k, 0.0), bm2.get(k, 0.0))
        if s > 0 and val <= tol * s:
            noise += 1
        else:
            bad += 1
    return max_rel, bad, noise


def rel_inf(a, b):
    a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    s = np.max(np.abs(b)) if b.size else 1.0
    return float(np.max(np.abs(a - b)) / (s if s > 0 else 1.0)) if a.size else 0.0


def rel_l2(a, b):
    a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def oracle_rows(geometry, lo, hi, name, p, n_cols):
    """Block rows [lo, hi) of the oracle's B (global column numbering) and L, assembled on the ghost-inclusive sub-mesh cut
    by ``partition_geometry`` (how the oracle reaches workloads that are too large for one call)."""
    from oracle import dg_oracle
    from reyna_b200.partition import partition_geometry
    from reyna_b200.problems import PROBLEMS
    g = dg_oracle.flatten_partition(partition_geometry(geometry, lo, hi))
    Bp, Lp, info = dg_oracle.assemble(g, p, mirror_index_quirk=False, **PROBLEMS[name]["data"])
    d, R = info.dim_elem, g.n_rows
    coo = Bp[:R * d].tocoo()
    cols = g.elem_gid[coo.col // d] * d + coo.col % d
    B = csr_matrix((coo.data, (coo.row, cols)), shape=(R * d, n_cols))
    return B, np.asarray(Lp[:R * d].todense()).reshape(-1), d
