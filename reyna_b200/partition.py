"""Element partitioning for the multi-GPU path (one process per GPU).

The reference is a single-process program; this module is new.  Elements are split into contiguous index ranges (the
benchmark meshes are numbered along a Morton curve, so a range is a compact patch; ``morton_partition_order`` gives the
permutation for arbitrary meshes).  Rank r owns the block rows of its elements:

* volume, boundary-face and interior-face terms of an owned element are computed by its owner only -- a face cut by the
  partition is visited from both sides, each rank producing just its own rows, so the ASSEMBLY needs no collective at all
  (the read-only geometry of the neighbouring "ghost" elements is replicated at partition time);
* the Krylov solver needs, per SpMV, the solution coefficients of the ghost elements: ``HaloPlan`` lists, per peer rank,
  which owned elements to send and where the received ghosts go (both sorted by global element id, so the plans of two
  ranks agree without any negotiation).
"""
from __future__ import annotations

import types
import typing

import numpy as np

from .geometry import flat_geometry


def partition_bounds(n_elements: int, world: int) -> np.ndarray:
    """Contiguous, balanced element ranges: rank r owns [bounds[r], bounds[r+1]).  Ranges start at multiples of 4 so that the
    groups of 4 consecutive elements the multigrid hierarchy agglomerates never straddle two ranks."""
    b = ((np.arange(world + 1, dtype=np.int64) * n_elements) // world) // 4 * 4
    b[-1] = n_elements
    return b


def morton_partition_order(geometry) -> np.ndarray:
    """Permutation (new -> old) that numbers elements along a Z-order curve of their bounding-box centres: the order in which
    contiguous index ranges are compact patches (partitions with short halos, multigrid aggregates of neighbouring
    elements).  The reference's mesher numbers elements in random seed order (polymesher/two_dimensional/main.py:111-165)."""
    from .mesh import morton_order
    bb = flat_geometry(geometry).bbox
    return morton_order(np.column_stack((0.5 * (bb[:, 0] + bb[:, 1]), 0.5 * (bb[:, 2] + bb[:, 3]))))


def locality_order(geometry, factor: float = 8.0) -> typing.Optional[np.ndarray]:
    """``None`` when the element numbering already has spatial locality (mean index distance of face neighbours at most
    ``factor * sqrt(E)``: row-major, Morton, Hilbert ... numberings), otherwise ``morton_partition_order(geometry)``.
    Cached on the geometry object."""
    cached = getattr(geometry, '_rb_locality', None)
    if cached is not None:
        return cached[0]
    fg = flat_geometry(geometry)
    perm = None
    if fg.n_iface > 0 and fg.n_elem >= 64:
        gap = float(np.abs(fg.ielem[:, 0].astype(np.int64) - fg.ielem[:, 1].astype(np.int64)).mean())
        if gap > factor * np.sqrt(fg.n_elem):
            perm = morton_partition_order(geometry)
    try:
        geometry._rb_locality = (perm,)
    except Exception:   # pragma: no cover - exotic geometry objects
        pass
    return perm


class RenumberedGeometry:
    """The same geometry with elements renumbered by ``perm`` (new -> old); faces, nodes and boundary edges keep their
    numbering.  Carries the attribute names ``DGFEM`` / ``BoundaryInformation`` read; ``old_of_new`` / ``new_of_old`` map
    element ids."""

    def __init__(self, parent, perm: np.ndarray):
        fg = flat_geometry(parent)
        E = fg.n_elem
        perm = np.asarray(perm, dtype=np.int64)
        if perm.shape != (E,) or not np.array_equal(np.sort(perm), np.arange(E)):
            raise ValueError('perm must be a permutation of the elements')
        inv = np.empty(E, dtype=np.int64)
        inv[perm] = np.arange(E)
        self.parent, self.old_of_new, self.new_of_old = parent, perm, inv

        def ragged(off, idx):
            cnt = np.diff(off.astype(np.int64))[perm]
            new_off = np.zeros(E + 1, dtype=np.int64)
            np.cumsum(cnt, out=new_off[1:])
            src = np.repeat(off[:-1].astype(np.int64)[perm] - new_off[:-1], cnt) + np.arange(int(new_off[-1]), dtype=np.int64)
            return new_off, idx[src], cnt

        tri_off, tri_sel, tcnt = ragged(fg.tri_off, np.arange(fg.n_tri, dtype=np.int64))
        poly_off, poly_vtx, _ = ragged(fg.poly_off, fg.poly_vtx)
        f = types.SimpleNamespace(
            n_nodes=fg.n_nodes, n_elem=E, n_tri=fg.n_tri, n_iface=fg.n_iface, n_bface=fg.n_bface, nodes=fg.nodes,
            tri=np.ascontiguousarray(fg.tri[tri_sel]), tri_elem=np.repeat(np.arange(E, dtype=np.int32), tcnt),
            tri_off=tri_off.astype(np.int32), bbox=np.ascontiguousarray(fg.bbox[perm]), area=np.ascontiguousarray(fg.area[perm]),
            poly_off=poly_off.astype(np.int32), poly_vtx=np.ascontiguousarray(poly_vtx),
            iedge=fg.iedge, ielem=np.ascontiguousarray(inv[fg.ielem.astype(np.int64)].astype(np.int32)), inorm=fg.inorm,
            bedge=fg.bedge, belem=inv[fg.belem.astype(np.int64)].astype(np.int32), bnorm=fg.bnorm,
        )
        self._rb_flat = f
        self.nodes = fg.nodes
        self.n_elements = E
        self.h = parent.h
        self.boundary_edges = f.bedge
        self.boundary_edges_to_element = f.belem
        self.boundary_normals = f.bnorm
        self.interior_edges = f.iedge
        self.interior_edges_to_element = f.ielem
        self.interior_normals = f.inorm
        self.elem_bounding_boxes = f.bbox
        self.areas = f.area
        self.mesh = getattr(parent, "mesh", None)


class PartitionedGeometry:
    """The slice of a DGFEMGeometry that one rank needs: owned elements [0, n_rows) followed by ghost elements
    [n_rows, n_elem) in ascending global order.  Carries the attribute names ``DGFEM`` reads."""

    def __init__(self, parent, lo: int, hi: int):
        fg = flat_geometry(parent)
        E = fg.n_elem
        if not (0 <= lo <= hi <= E):
            raise ValueError("invalid element range")
        self.parent = parent
        self.lo, self.hi = int(lo), int(hi)
        self.n_global_elements = E
        R = hi - lo
        k1, k2 = fg.ielem[:, 0].astype(np.int64), fg.ielem[:, 1].astype(np.int64)
        own1 = (k1 >= lo) & (k1 < hi)
        own2 = (k2 >= lo) & (k2 < hi)
        fsel = np.flatnonzero(own1 | own2)                       # faces touching an owned element, global order kept
        other = np.concatenate((k2[fsel][own1[fsel] & ~own2[fsel]], k1[fsel][own2[fsel] & ~own1[fsel]]))
        ghosts = np.unique(other)                                # ascending global ids
        gid = np.concatenate((np.arange(lo, hi, dtype=np.int64), ghosts))
        g2l = np.full(E, -1, dtype=np.int64)
        g2l[gid] = np.arange(gid.shape[0])
        self.ghost_gid = ghosts
        self.elem_gid = gid

        tsel = slice(int(fg.tri_off[lo]), int(fg.tri_off[hi]))
        bsel = np.flatnonzero((fg.belem >= lo) & (fg.belem < hi))
        poly_cnt = np.diff(fg.poly_off)[gid]
        poly_off = np.zeros(gid.shape[0] + 1, dtype=np.int64)
        np.cumsum(poly_cnt, out=poly_off[1:])
        starts = fg.poly_off[gid].astype(np.int64)
        poly_idx = np.repeat(starts - poly_off[:-1], poly_cnt) + np.arange(int(poly_off[-1]), dtype=np.int64)

        f = types.SimpleNamespace(
            n_nodes=fg.n_nodes, n_elem=int(gid.shape[0]), n_rows=int(R), n_tri=int(tsel.stop - tsel.start),
            n_iface=int(fsel.shape[0]), n_bface=int(bsel.shape[0]),
            nodes=fg.nodes,
            tri=np.ascontiguousarray(fg.tri[tsel]),
            tri_elem=(fg.tri_elem[tsel].astype(np.int64) - lo).astype(np.int32),
            tri_off=(fg.tri_off[lo:hi + 1].astype(np.int64) - int(fg.tri_off[lo])).astype(np.int32),
            bbox=np.ascontiguousarray(fg.bbox[gid]), area=np.ascontiguousarray(fg.area[gid]),
            poly_off=poly_off.astype(np.int32), poly_vtx=np.ascontiguousarray(fg.poly_vtx[poly_idx]),
            iedge=np.ascontiguousarray(fg.iedge[fsel]),
            ielem=np.ascontiguousarray(np.column_stack((g2l[k1[fsel]], g2l[k2[fsel]])).astype(np.int32)),
            inorm=np.ascontiguousarray(fg.inorm[fsel]),
            bedge=np.ascontiguousarray(fg.bedge[bsel]),
            belem=(fg.belem[bsel].astype(np.int64) - lo).astype(np.int32),
            bnorm=np.ascontiguousarray(fg.bnorm[bsel]),
            elem_gid=gid.astype(np.int32),
        )
        self._rb_flat = f
        self.face_gid = fsel
        self.bface_gid = bsel
        # attribute names DGFEM / BoundaryInformation read
        self.nodes = fg.nodes
        self.n_elements = int(R)
        self.h = parent.h
        self.boundary_edges = f.bedge
        self.boundary_edges_to_element = f.belem
        self.boundary_normals = f.bnorm
        self.interior_edges = f.iedge
        self.mesh = getattr(parent, "mesh", None)


def partition_geometry(geometry, lo: int, hi: int) -> PartitionedGeometry:
    return PartitionedGeometry(geometry, lo, hi)


class HaloPlan:
    """Who sends which owned elements to whom, and where received ghosts land (element granularity; the solver moves
    ``d`` coefficients per element).  ``send[q]`` = local indices of owned elements rank q needs, ascending global id;
    ``recv[q]`` = local ghost indices (>= n_rows) filled by rank q, ascending global id."""

    def __init__(self, geometry, bounds: np.ndarray, rank: int):
        fg = flat_geometry(geometry)
        bounds = np.asarray(bounds, dtype=np.int64)
        world = bounds.shape[0] - 1
        self.rank, self.world = int(rank), int(world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        k1, k2 = fg.ielem[:, 0].astype(np.int64), fg.ielem[:, 1].astype(np.int64)
        o1 = np.searchsorted(bounds, k1, side="right") - 1          # owner rank of each face side
        o2 = np.searchsorted(bounds, k2, side="right") - 1
        cut = o1 != o2
        a = np.concatenate((k1[cut], k2[cut]))                       # element ...
        oa = np.concatenate((o1[cut], o2[cut]))                      # ... its owner ...
        ob = np.concatenate((o2[cut], o1[cut]))                      # ... and the rank that needs it as a ghost
        self.send: typing.Dict[int, np.ndarray] = {}
        self.recv: typing.Dict[int, np.ndarray] = {}
        mine = oa == rank
        for q in np.unique(ob[mine]):
            self.send[int(q)] = np.unique(a[mine & (ob == q)]) - lo
        need = ob == rank
        ghosts = np.unique(a[need])
        n_rows = hi - lo
        for q in np.unique(oa[need]):
            g = np.unique(a[need & (oa == q)])
            self.recv[int(q)] = n_rows + np.searchsorted(ghosts, g)
        self.n_rows, self.n_ghost = n_rows, int(ghosts.shape[0])
        self.ghost_gid = ghosts

    def peers(self):
        return sorted(set(self.send) | set(self.recv))

    def exchange_torch(self, x, d: int, group=None):
        """Refresh the ghost part of ``x`` ((n_rows + n_ghost) * d values, torch tensor on any device) with
        ``torch.distributed`` point-to-point operations (gloo on CPU, NCCL on GPU)."""
        import torch
        import torch.distributed as dist
        xv = x.view(-1, d)
        ops, bufs = [], []
        for q in self.peers():
            if q in self.send:
                sb = xv[torch.as_tensor(self.send[q], device=x.device)].contiguous()
                ops.append(dist.P2POp(dist.isend, sb, q, group=group))
            if q in self.recv:
                rb = torch.empty((self.recv[q].shape[0], d), dtype=x.dtype, device=x.device)
                ops.append(dist.P2POp(dist.irecv, rb, q, group=group))
                bufs.append((q, rb))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for q, rb in bufs:
            xv[torch.as_tensor(self.recv[q], device=x.device)] = rb
        return x
