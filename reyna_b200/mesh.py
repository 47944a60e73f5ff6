"""Host-side mesh container and a fast synthetic bounded-Voronoi generator for the benchmarks.

``PolyMesh`` is the data contract the reference passes from ``poly_mesher`` to ``DGFEMGeometry``
(reyna/polymesher/two_dimensional/_auxilliaries/abstraction.py:56-77): ``vertices (Nn,2)``, ``filtered_regions``
(list of int arrays, counter-clockwise), ``filtered_points (E,2)``, ``domain``.  Mesh GENERATION is out of scope of
the B200 port (north star: it stays host-side); ``voronoi_mesh`` exists because the reference's ``poly_mesher`` is not
present on the GPU box and takes minutes at 1M elements.  It follows the same recipe (reflect seeds near the
boundary, scipy/qhull Voronoi, Lloyd relaxation -- polymesher/two_dimensional/main.py:62-105) with the per-element
Python/numba loops replaced by flat-array numpy, for rectangular domains only.
"""
from __future__ import annotations

import itertools

import numpy as np
from scipy.spatial import Voronoi


class RectangleDomain:
    """Axis-aligned rectangle ``[[x0, x1], [y0, y1]]`` (cf. polymesher/two_dimensional/domains.py:35-50)."""

    def __init__(self, bounding_box, fixed_points=None):
        self.bounding_box = np.asarray(bounding_box, dtype=float)
        self.fixed_points = fixed_points

    def area(self) -> float:
        b = self.bounding_box
        return float((b[0, 1] - b[0, 0]) * (b[1, 1] - b[1, 0]))

    def pFix(self):
        return self.fixed_points


class PolyMesh:
    """vertices / filtered_regions / filtered_points / domain -- abstraction.py:73-77."""

    def __init__(self, vertices, filtered_regions, filtered_points, domain):
        self.vertices = vertices
        self.filtered_regions = filtered_regions
        self.filtered_points = filtered_points
        self.domain = domain


class FlatRegions:
    """Ragged polygon list stored flat (``offsets (E+1,)``, ``indices``); behaves like a list of int64 arrays."""

    def __init__(self, offsets: np.ndarray, indices: np.ndarray):
        self.offsets = np.asarray(offsets, dtype=np.int64)
        self.indices = np.asarray(indices, dtype=np.int64)

    def __len__(self) -> int:
        return self.offsets.shape[0] - 1

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return self.indices[self.offsets[i]:self.offsets[i + 1]]

    def __iter__(self):
        for i in range(len(self)):
            yield self.indices[self.offsets[i]:self.offsets[i + 1]]


def flatten_regions(regions):
    """(offsets int64 (E+1,), indices int64) of any list-of-index-lists; zero-copy for ``FlatRegions``."""
    if isinstance(regions, FlatRegions):
        return regions.offsets, regions.indices
    n = len(regions)
    lens = np.fromiter((len(r) for r in regions), dtype=np.int64, count=n)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    idx = np.fromiter(itertools.chain.from_iterable(regions), dtype=np.int64, count=int(off[-1]))
    return off, idx


def _reflect(points: np.ndarray, box: np.ndarray, alpha: float) -> np.ndarray:
    """Mirror images of the seeds closer than ``alpha`` to each side of the rectangle."""
    (x0, x1), (y0, y1) = box
    out = []
    m = points[:, 0] - x0 < alpha
    out.append(np.column_stack((2 * x0 - points[m, 0], points[m, 1])))
    m = x1 - points[:, 0] < alpha
    out.append(np.column_stack((2 * x1 - points[m, 0], points[m, 1])))
    m = points[:, 1] - y0 < alpha
    out.append(np.column_stack((points[m, 0], 2 * y0 - points[m, 1])))
    m = y1 - points[:, 1] < alpha
    out.append(np.column_stack((points[m, 0], 2 * y1 - points[m, 1])))
    return np.concatenate(out, axis=0)


def _cells(points: np.ndarray, box: np.ndarray):
    """Bounded Voronoi cells of ``points``: (vertex coords, offsets, flat CCW vertex indices)."""
    n = points.shape[0]
    area = float((box[0, 1] - box[0, 0]) * (box[1, 1] - box[1, 0]))
    alpha = 1.5 * np.sqrt(area / n) * 2.0
    allp = np.concatenate((points, _reflect(points, box, alpha)), axis=0)
    vor = Voronoi(allp, qhull_options='Qbb Qz')
    # ridges -> (cell, vertex) incidences for the real seeds only
    rp = vor.ridge_points
    rv = np.asarray(vor.ridge_vertices, dtype=np.int64)
    ok = (rv >= 0).all(axis=1)
    rp, rv = rp[ok], rv[ok]
    cell = np.concatenate((rp[:, 0], rp[:, 0], rp[:, 1], rp[:, 1]))
    vert = np.concatenate((rv[:, 0], rv[:, 1], rv[:, 0], rv[:, 1]))
    keep = cell < n
    cell, vert = cell[keep], vert[keep]
    # unique (cell, vertex) pairs, then counter-clockwise order about the seed
    key = np.unique(cell.astype(np.int64) * (vor.vertices.shape[0] + 1) + vert)
    cell = key // (vor.vertices.shape[0] + 1)
    vert = key % (vor.vertices.shape[0] + 1)
    d = vor.vertices[vert] - points[cell]
    ang = np.arctan2(d[:, 1], d[:, 0])
    order = np.lexsort((ang, cell))
    cell, vert = cell[order], vert[order]
    counts = np.bincount(cell, minlength=n)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    return vor.vertices, off, vert


def _centroids(verts: np.ndarray, off: np.ndarray, idx: np.ndarray):
    """Polygon centroids and areas from flat CCW polygons."""
    nxt = np.arange(idx.shape[0]) + 1
    last = off[1:] - 1
    nxt[last] = off[:-1]
    p, q = verts[idx], verts[idx[nxt]]
    cr = p[:, 0] * q[:, 1] - q[:, 0] * p[:, 1]
    a = 0.5 * np.add.reduceat(cr, off[:-1])
    cx = np.add.reduceat((p[:, 0] + q[:, 0]) * cr, off[:-1]) / (6.0 * a)
    cy = np.add.reduceat((p[:, 1] + q[:, 1]) * cr, off[:-1]) / (6.0 * a)
    return np.column_stack((cx, cy)), a


def morton_order(points: np.ndarray, bits: int = 16) -> np.ndarray:
    """Permutation sorting 2-D points along the Z-order curve (locality for the GPU kernels and the partitioner)."""
    lo, hi = points.min(axis=0), points.max(axis=0)
    q = ((points - lo) / np.maximum(hi - lo, 1e-300) * ((1 << bits) - 1)).astype(np.uint64)

    def spread(v):
        v = (v | (v << 16)) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << 8)) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << 4)) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << 2)) & np.uint64(0x3333333333333333)
        v = (v | (v << 1)) & np.uint64(0x5555555555555555)
        return v

    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))
    return np.argsort(code, kind='stable')


def voronoi_mesh(n_points: int, seed: int = 1337, lloyd_iterations: int = 3, bounding_box=((0.0, 1.0), (0.0, 1.0)),
                 morton: bool = True, jitter: float = 0.35) -> PolyMesh:
    """Synthetic bounded Voronoi mesh of a rectangle with ``n_points`` convex polygonal elements.

    Seeds start on a jittered lattice (so a few Lloyd sweeps already give well-shaped cells), are optionally sorted in
    Morton order (element index locality), and the result has compactly numbered vertices.  Deterministic in ``seed``."""
    box = np.asarray(bounding_box, dtype=float)
    rng = np.random.default_rng(seed)
    w, h = box[0, 1] - box[0, 0], box[1, 1] - box[1, 0]
    nx = max(1, int(round(np.sqrt(n_points * w / h))))
    ny = max(1, int(np.ceil(n_points / nx)))
    gx, gy = np.meshgrid((np.arange(nx) + 0.5) / nx, (np.arange(ny) + 0.5) / ny, indexing='ij')
    pts = np.column_stack((gx.ravel(), gy.ravel()))
    if pts.shape[0] > n_points:
        pts = pts[rng.permutation(pts.shape[0])[:n_points]]
    pts = pts + jitter * (rng.random(pts.shape) - 0.5) * 2.0 * np.array([1.0 / nx, 1.0 / ny])
    pts = np.clip(pts, 1e-6, 1 - 1e-6)
    pts = pts * np.array([w, h]) + box[:, 0]
    for _ in range(lloyd_iterations):
        verts, off, idx = _cells(pts, box)
        pts, _ = _centroids(verts, off, idx)
    if morton:
        pts = pts[morton_order(pts)]
    verts, off, idx = _cells(pts, box)
    used, inv = np.unique(idx, return_inverse=True)
    vertices = np.ascontiguousarray(verts[used])
    # snap boundary vertices onto the rectangle (the reflection puts them there up to rounding)
    regions = FlatRegions(off, inv.astype(np.int64))
    return PolyMesh(vertices, regions, np.ascontiguousarray(pts), RectangleDomain(box))


def tiled_voronoi_mesh(n_points: int, tiles: int = 4, seed: int = 1337, lloyd_iterations: int = 3,
                       bounding_box=((0.0, 1.0), (0.0, 1.0))) -> PolyMesh:
    """Large synthetic Voronoi mesh from ``tiles x tiles`` mirrored copies of one ``n_points / tiles^2``-element mesh.

    A bounded Voronoi mesh of a rectangle IS the restriction of the Voronoi diagram of the seeds plus their mirror images
    across the sides (that is how ``_cells`` and the reference's ``_poly_mesher_reflect`` bound it), so gluing alternately
    mirrored copies of a base mesh gives exactly the bounded Voronoi mesh of the tiled seed set: conforming across the
    tile sides, convex cells, same element statistics -- at 1/tiles^2 of the qhull cost (1M elements in seconds instead
    of minutes).  Tiles follow a Z-order curve and elements keep the base mesh's Morton order inside a tile, so the global
    element numbering is spatially local (contiguous index ranges are compact patches: the multi-GPU partition)."""
    k = int(tiles)
    if k < 1 or n_points % (k * k):
        raise ValueError("n_points must be a multiple of tiles^2")
    box = np.asarray(bounding_box, dtype=float)
    base = voronoi_mesh(n_points // (k * k), seed=seed, lloyd_iterations=lloyd_iterations,
                        bounding_box=((0.0, 1.0), (0.0, 1.0)), morton=True)
    off, idx = flatten_regions(base.filtered_regions)
    bv, bp = base.vertices, base.filtered_points
    nv, ne = bv.shape[0], bp.shape[0]
    cnt = np.diff(off)
    # reversed vertex order inside each polygon (a single mirror flips the orientation; keep every cell counter-clockwise)
    pos = np.arange(idx.shape[0], dtype=np.int64) - np.repeat(off[:-1], cnt)
    rev = np.repeat(off[:-1] + cnt - 1, cnt) - pos
    idx_rev = idx[rev]

    tx, ty = np.meshgrid(np.arange(k), np.arange(k), indexing='ij')
    tx, ty = tx.ravel(), ty.ravel()
    order = morton_order(np.column_stack((tx, ty)).astype(float) + 0.5, bits=8) if k > 1 else np.array([0])
    V = np.empty((k * k * nv, 2))
    P = np.empty((k * k * ne, 2))
    I = np.empty(k * k * idx.shape[0], dtype=np.int64)
    for slot, t in enumerate(order):
        i, j = int(tx[t]), int(ty[t])
        fx, fy = i % 2 == 1, j % 2 == 1
        vx = (1.0 - bv[:, 0]) if fx else bv[:, 0]
        vy = (1.0 - bv[:, 1]) if fy else bv[:, 1]
        V[slot * nv:(slot + 1) * nv, 0] = (i + vx) / k
        V[slot * nv:(slot + 1) * nv, 1] = (j + vy) / k
        px = (1.0 - bp[:, 0]) if fx else bp[:, 0]
        py = (1.0 - bp[:, 1]) if fy else bp[:, 1]
        P[slot * ne:(slot + 1) * ne, 0] = (i + px) / k
        P[slot * ne:(slot + 1) * ne, 1] = (j + py) / k
        I[slot * idx.shape[0]:(slot + 1) * idx.shape[0]] = (idx_rev if (fx != fy) else idx) + slot * nv
    # merge the duplicated vertices on the shared tile sides: only vertices lying on a side of the base tile can coincide
    eps = 1e-12
    on_side = (np.abs(bv[:, 0]) < eps) | (np.abs(bv[:, 0] - 1.0) < eps) | (np.abs(bv[:, 1]) < eps) | \
              (np.abs(bv[:, 1] - 1.0) < eps)
    side_ids = (np.flatnonzero(on_side)[None, :] + (np.arange(k * k) * nv)[:, None]).ravel()
    q = np.rint(V[side_ids] * float(1 << 44)).astype(np.int64)
    srt = np.lexsort((q[:, 1], q[:, 0]))
    qs = q[srt]
    newgrp = np.ones(qs.shape[0], dtype=bool)
    newgrp[1:] = np.any(qs[1:] != qs[:-1], axis=1)
    grp = np.cumsum(newgrp) - 1
    rep = side_ids[srt][np.flatnonzero(newgrp)]            # representative (first) vertex of every group
    remap = np.arange(V.shape[0], dtype=np.int64)
    remap[side_ids[srt]] = rep[grp]
    keep = np.zeros(V.shape[0], dtype=bool)
    keep[remap] = True
    newid = np.cumsum(keep) - 1
    inv = newid[remap]
    vertices = V[keep]
    offsets = np.concatenate([off[:-1] + s * idx.shape[0] for s in range(k * k)] + [[k * k * idx.shape[0]]]).astype(np.int64)
    w, h = box[0, 1] - box[0, 0], box[1, 1] - box[1, 0]
    vertices = vertices * np.array([w, h]) + box[:, 0]
    P = P * np.array([w, h]) + box[:, 0]
    return PolyMesh(np.ascontiguousarray(vertices), FlatRegions(offsets, inv[I].astype(np.int64)),
                    np.ascontiguousarray(P), RectangleDomain(box))
