"""``DGFEMGeometry`` -- host-side geometry of a polygonal mesh, vectorised.

Drop-in for reyna/geometry/two_dimensional/DGFEM.py:15-269: same constructor, same attribute names, same array
conventions (edge rows ``[min, max]`` ordered by ``(max, min)`` vertex id as scipy.sparse.find returns them, :187-193;
``interior_edges_to_element`` rows ascending, :208-213; interior normals pointing from the first to the second element,
:240-255; boundary normals outward, :224-238; bounding boxes ``[xmin, xmax, ymin, ymax]``, :131-132; vertices rounded
in place by ``*= 1e8; /= 1e8``, :111-112).  Differences, all deliberate:

* no shapely: areas by the shoelace formula, ``h_s`` = longer side of the minimum-area enclosing rectangle by rotating
  calipers over the polygon's edges (what ``minimum_rotated_rectangle`` returns for the convex elements the reference
  supports, DGFEM.py:49-52) -- equal to the GEOS values up to rounding;
* the per-element Python loop (:129-157) and the edge dictionary (:197-220) are replaced by flat-array numpy, so a
  1M-element mesh takes seconds instead of minutes;
* ``subtriangulation='delaunay'`` (default) calls scipy/qhull per polygon exactly as the reference does (:142) and gives
  identical ``subtriangulation`` rows; ``'fan'`` is the fast vectorised alternative (same triangle count);
* ``elem_bounding_boxes`` is an ``(E,4)`` array and ``triangle_to_polygon`` an int array (the reference holds Python
  lists); indexing behaves the same.

Geometry generation is NOT part of the GPU hot path (north star); the flat int32/float64 arrays the kernels consume are
produced once by ``flat()`` and cached.
"""
from __future__ import annotations

import pickle
import time
import types
import typing

import numpy as np

from .mesh import flatten_regions


def _min_rect_long_side(X: np.ndarray, Y: np.ndarray, valid: np.ndarray) -> np.ndarray:
    """Longer side of the minimum-area enclosing rectangle of each (convex) polygon.

    ``X, Y`` are padded (E, M) vertex coordinates, ``valid`` the (E, M) mask; the candidate orientations are the
    polygon's own edges (rotating calipers)."""
    E, M = X.shape
    cnt = valid.sum(axis=1)
    best_area = np.full(E, np.inf)
    best_long = np.zeros(E)
    big = 1e300
    for k in range(M):
        live = k < cnt
        if not live.any():
            break
        k1 = np.where(k + 1 < cnt, k + 1, 0)
        ex = X[np.arange(E), k1] - X[:, k]
        ey = Y[np.arange(E), k1] - Y[:, k]
        ln = np.hypot(ex, ey)
        ok = live & (ln > 0)
        ln = np.where(ok, ln, 1.0)
        ux, uy = ex / ln, ey / ln
        s = X * ux[:, None] + Y * uy[:, None]
        t = -X * uy[:, None] + Y * ux[:, None]
        smax = np.where(valid, s, -big).max(axis=1)
        smin = np.where(valid, s, big).min(axis=1)
        tmax = np.where(valid, t, -big).max(axis=1)
        tmin = np.where(valid, t, big).min(axis=1)
        a = (smax - smin) * (tmax - tmin)
        better = ok & (a < best_area)
        best_area = np.where(better, a, best_area)
        best_long = np.where(better, np.maximum(smax - smin, tmax - tmin), best_long)
    return best_long


class DGFEMGeometry:
    """Geometry information a DGFEM method needs: normals, sub-triangulation, edge/element maps, sizes."""

    def __init__(self, poly_mesh, **kwargs):
        self.mesh = poly_mesh
        self.n_nodes = poly_mesh.vertices.shape[0]
        self.n_elements = len(poly_mesh.filtered_regions)
        self.nodes = poly_mesh.vertices

        self.elem_bounding_boxes = None
        self.boundary_edges = None
        self.boundary_edges_to_element = None
        self.interior_edges = None
        self.interior_edges_to_element = None
        self.boundary_normals = None
        self.interior_normals = None
        self.subtriangulation = None
        self.n_triangles = None
        self.triangle_to_polygon = None
        self.h: typing.Optional[float] = None
        self.h_s: typing.Optional[np.ndarray] = None
        self.areas: typing.Optional[np.ndarray] = None

        time_generation = kwargs.pop('time', False)
        self._subtriangulation_mode = kwargs.pop('subtriangulation', 'delaunay')
        if self._subtriangulation_mode not in ('delaunay', 'fan'):
            raise ValueError("subtriangulation must be 'delaunay' or 'fan'")
        self._flat = None
        _t = time.time()
        self._generate()
        if time_generation:
            print(f"Time taken to generate geometry: {time.time() - _t}s")

    # ------------------------------------------------------------------------------------------------------------
    def _generate(self) -> None:
        self.nodes *= 1e8
        self.nodes /= 1e8
        nodes = self.nodes
        off, idx = flatten_regions(self.mesh.filtered_regions)
        E = self.n_elements
        cnt = np.diff(off)
        M = int(cnt.max()) if E else 0
        elem_of = np.repeat(np.arange(E, dtype=np.int64), cnt)
        pos = np.arange(idx.shape[0], dtype=np.int64) - off[elem_of]

        # padded coordinates for the per-element reductions
        X = np.zeros((E, M))
        Y = np.zeros((E, M))
        valid = np.zeros((E, M), dtype=bool)
        X[elem_of, pos] = nodes[idx, 0]
        Y[elem_of, pos] = nodes[idx, 1]
        valid[elem_of, pos] = True
        big = 1e300
        self.elem_bounding_boxes = np.column_stack((
            np.where(valid, X, big).min(axis=1), np.where(valid, X, -big).max(axis=1),
            np.where(valid, Y, big).min(axis=1), np.where(valid, Y, -big).max(axis=1)))

        nxt = np.arange(idx.shape[0], dtype=np.int64) + 1
        nxt[off[1:] - 1] = off[:-1]
        px, py = nodes[idx, 0], nodes[idx, 1]
        qx, qy = nodes[idx[nxt], 0], nodes[idx[nxt], 1]
        self.areas = 0.5 * np.abs(np.add.reduceat(px * qy - qx * py, off[:-1])) if E else np.zeros(0)
        self.h_s = _min_rect_long_side(X, Y, valid)
        self.h = float(np.max(self.h_s)) if E else 0.0

        # ---- sub-triangulation (DGFEM.py:142-150) ----------------------------------------------------------------
        if self._subtriangulation_mode == 'delaunay':
            from scipy.spatial import Delaunay
            tris, t2p = [], []
            for i in range(E):
                elem = idx[off[i]:off[i + 1]]
                simp = Delaunay(nodes[elem, :]).simplices
                tris.append(elem[simp])
                t2p.append(np.full(simp.shape[0], i, dtype=np.int64))
            self.subtriangulation = np.concatenate(tris, axis=0).astype(np.int64) if tris else np.zeros((0, 3), np.int64)
            self.triangle_to_polygon = np.concatenate(t2p) if t2p else np.zeros(0, np.int64)
        else:
            ntri = cnt - 2
            t2p = np.repeat(np.arange(E, dtype=np.int64), ntri)
            toff = np.zeros(E + 1, dtype=np.int64)
            np.cumsum(ntri, out=toff[1:])
            k = np.arange(t2p.shape[0], dtype=np.int64) - toff[t2p]
            base = off[t2p]
            self.subtriangulation = np.column_stack((idx[base], idx[base + k + 1], idx[base + k + 2]))
            self.triangle_to_polygon = t2p
        self.n_triangles = self.subtriangulation.shape[0]

        # ---- edge classification (DGFEM.py:183-193) --------------------------------------------------------------
        a, b = idx, idx[nxt]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        key = hi * np.int64(self.n_nodes) + lo
        order = np.argsort(key, kind='stable')          # (max, min) order; ties keep ascending element index
        ks = key[order]
        first = np.ones(ks.shape[0], dtype=bool)
        first[1:] = ks[1:] != ks[:-1]
        start = np.flatnonzero(first)
        mult = np.diff(np.append(start, ks.shape[0]))
        u_lo, u_hi = lo[order][start], hi[order][start]
        owner = elem_of[order]
        is_b, is_i = mult == 1, mult == 2
        self.boundary_edges = np.column_stack((u_lo[is_b], u_hi[is_b])).astype(np.int32)
        self.interior_edges = np.column_stack((u_lo[is_i], u_hi[is_i])).astype(np.int32)
        # elements of each edge (DGFEM.py:195-220); rows ascending because the stable sort keeps element order
        self.boundary_edges_to_element = owner[start[is_b]].astype(np.int64)
        si = start[is_i]
        self.interior_edges_to_element = np.column_stack((owner[si], owner[si + 1])).astype(np.int64) \
            if si.size else np.zeros((0, 2), dtype=np.int64)

        # ---- unit normals (DGFEM.py:222-255) -----------------------------------------------------------------------
        fp = np.asarray(self.mesh.filtered_points, dtype=float)
        be = self.boundary_edges
        tan = nodes[be[:, 0], :] - nodes[be[:, 1], :]
        nrm_c = np.sqrt(tan[:, 0] ** 2 + tan[:, 1] ** 2)
        tan = np.roll(tan, -1, axis=1)
        tan[:, 1] *= -1
        nvec = np.divide(tan, nrm_c[:, np.newaxis])
        outward = nodes[be[:, 0], :] - fp[self.boundary_edges_to_element, :]
        flip = np.maximum(np.sum(nvec * outward, axis=1), 0) == 0
        nvec[flip, :] = -nvec[flip, :]
        self.boundary_normals = nvec

        if self.interior_edges.shape[0] > 1:
            ie = self.interior_edges
            tan = nodes[ie[:, 0], :] - nodes[ie[:, 1], :]
            nrm_c = np.sqrt(tan[:, 0] ** 2 + tan[:, 1] ** 2)
            tan = np.roll(tan, -1, axis=1)
            tan[:, 1] *= -1
            nvec = np.divide(tan, nrm_c[:, np.newaxis])
            outward = fp[self.interior_edges_to_element[:, 1], :] - fp[self.interior_edges_to_element[:, 0], :]
            flip = np.sum(nvec * outward, axis=1) < 0.0
            nvec[flip, :] = -nvec[flip, :]
            self.interior_normals = nvec

    # ------------------------------------------------------------------------------------------------------------
    def save_geometry(self, filepath: str, save_mesh: bool = False):
        """Pickle the attribute dictionary (DGFEM.py:257-269).  As in the reference, ``save_mesh=True`` EXCLUDES the
        mesh (the flag is inverted upstream); kept for drop-in compatibility."""
        with open(filepath, "wb") as file:
            pickle.dump({k: v for k, v in self.__dict__.items()
                         if k != ('mesh' if save_mesh else '') and not k.startswith('_')}, file)


_CACHE_ARRAYS = ('elem_bounding_boxes', 'boundary_edges', 'boundary_edges_to_element', 'interior_edges',
                 'interior_edges_to_element', 'boundary_normals', 'interior_normals', 'subtriangulation',
                 'triangle_to_polygon', 'h_s', 'areas')


def save_geometry_cache(geometry: "DGFEMGeometry", path: str) -> None:
    """Write mesh + geometry as ONE flat ``.npz`` (uncompressed: loading a 1M-element geometry takes about a second).

    This is the cache BASELINE config 5 asks for ("cached, generated once on host"); it replaces the reference's pickle of
    Python lists (``save_geometry``, geometry/two_dimensional/DGFEM.py:257-269), which has no loader."""
    off, idx = flatten_regions(geometry.mesh.filtered_regions)
    out = dict(vertices=np.asarray(geometry.nodes), region_offsets=off, region_indices=idx,
               filtered_points=np.asarray(geometry.mesh.filtered_points), h=np.float64(geometry.h))
    dom = getattr(geometry.mesh, 'domain', None)
    if dom is not None and hasattr(dom, 'bounding_box'):
        out['domain_box'] = np.asarray(dom.bounding_box, dtype=float)
    for name in _CACHE_ARRAYS:
        v = getattr(geometry, name)
        if v is not None:
            out[name] = np.asarray(v)
    tmp = path + '.tmp.npz'
    np.savez(tmp, **out)
    import os
    os.replace(tmp, path)


def load_geometry_cache(path: str) -> "DGFEMGeometry":
    """Rebuild a ``DGFEMGeometry`` from ``save_geometry_cache`` output without recomputing anything."""
    from .mesh import FlatRegions, PolyMesh, RectangleDomain
    z = np.load(path)
    dom = RectangleDomain(z['domain_box']) if 'domain_box' in z.files else None
    mesh = PolyMesh(z['vertices'], FlatRegions(z['region_offsets'], z['region_indices']), z['filtered_points'], dom)
    g = DGFEMGeometry.__new__(DGFEMGeometry)
    g.mesh = mesh
    g.nodes = mesh.vertices
    g.n_nodes = mesh.vertices.shape[0]
    g.n_elements = len(mesh.filtered_regions)
    for name in _CACHE_ARRAYS:
        setattr(g, name, z[name] if name in z.files else None)
    g.n_triangles = g.subtriangulation.shape[0]
    g.h = float(z['h'])
    g._subtriangulation_mode = 'cached'
    g._flat = None
    return g


def flat_geometry(geometry) -> types.SimpleNamespace:
    """Flat int32/float64 host arrays of any object with the reference's DGFEMGeometry attributes (cached on it)."""
    cached = getattr(geometry, '_rb_flat', None)
    if cached is not None:
        return cached
    off, idx = flatten_regions(geometry.mesh.filtered_regions)
    E = int(geometry.n_elements)
    t2p = np.asarray(geometry.triangle_to_polygon, dtype=np.int64).reshape(-1)
    if t2p.size and np.any(np.diff(t2p) < 0):
        raise ValueError('triangle_to_polygon must be non-decreasing')
    tri_off = np.zeros(E + 1, dtype=np.int64)
    np.cumsum(np.bincount(t2p, minlength=E), out=tri_off[1:])
    F = int(np.asarray(geometry.interior_edges).reshape(-1, 2).shape[0])
    inorm = geometry.interior_normals
    if inorm is None:
        if F > 0:   # the reference leaves interior_normals unset for a single interior edge (DGFEM.py:240)
            raise ValueError('geometry has interior edges but no interior_normals')
        inorm = np.zeros((0, 2))
    f = types.SimpleNamespace(
        n_nodes=int(np.asarray(geometry.nodes).shape[0]), n_elem=E, n_tri=int(t2p.shape[0]), n_iface=F,
        n_bface=int(np.asarray(geometry.boundary_edges).reshape(-1, 2).shape[0]),
        nodes=np.ascontiguousarray(geometry.nodes, dtype=np.float64),
        tri=np.ascontiguousarray(geometry.subtriangulation, dtype=np.int32).reshape(-1, 3),
        tri_elem=t2p.astype(np.int32), tri_off=tri_off.astype(np.int32),
        bbox=np.ascontiguousarray(np.asarray(geometry.elem_bounding_boxes, dtype=np.float64).reshape(-1, 4)),
        area=np.ascontiguousarray(geometry.areas, dtype=np.float64),
        poly_off=off.astype(np.int32), poly_vtx=idx.astype(np.int32),
        iedge=np.ascontiguousarray(geometry.interior_edges, dtype=np.int32).reshape(-1, 2),
        ielem=np.ascontiguousarray(geometry.interior_edges_to_element, dtype=np.int32).reshape(-1, 2),
        inorm=np.ascontiguousarray(inorm, dtype=np.float64).reshape(-1, 2),
        bedge=np.ascontiguousarray(geometry.boundary_edges, dtype=np.int32).reshape(-1, 2),
        belem=np.ascontiguousarray(geometry.boundary_edges_to_element, dtype=np.int32).reshape(-1),
        bnorm=np.ascontiguousarray(geometry.boundary_normals, dtype=np.float64).reshape(-1, 2),
    )
    try:
        geometry._rb_flat = f
    except Exception:   # pragma: no cover - exotic geometry objects
        pass
    return f
