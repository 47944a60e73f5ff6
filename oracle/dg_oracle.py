"""CPU oracle: a vectorised numpy/scipy restatement of reyna's DGFEM assemble-and-solve path.

TEST INFRASTRUCTURE ONLY -- the checker, never the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; ``reyna_b200/`` never does.

Parity status: PINNED.  The reference publishes no golden vectors (its tests build a fresh random mesh each time),
so this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the build container through
``oracle/ref_harness.py``: ``tests/test_oracle_vs_reference.py`` (live, when /root/reference exists) and the committed
fixtures under ``tests/golden/`` (written by ``oracle/make_golden.py``) which travel to the GPU box.

What is restated (all paths relative to /root/reference/reyna/DGFEM/two_dimensional/):

* ``quadrature``          <- main.py:233-257            (DGFEM._intialise_quadrature)
* ``basis_orders``        <- _auxilliaries/polygonal_basis_utils.py:7-26   (Basis_index2D)
* ``jacobi_orthonormal``  <- _auxilliaries/polygonal_basis_utils.py:29-78  (tensor_LegendreP)
* ``bbox_legendre``       <- _auxilliaries/assembly_aux.py:32-66           (tensor_shift_leg)
* ``basis_and_grad``      <- _auxilliaries/assembly_aux.py:69-133          (tensor_tensor_leg / tensor_gradtensor_leg)
* ``volume_terms``        <- main.py:275-327 + _auxilliaries/assembly/full_assembly.py:9-77   (localstiff)
* ``interior_face_terms`` <- main.py:329-374 + _auxilliaries/assembly/full_assembly.py:80-214 (int_localstiff)
* ``dirichlet_face_terms``<- main.py:376-427 + _auxilliaries/assembly/boundary_assembly.py:59-137
* ``inflow_face_terms``   <- main.py:429-474 + _auxilliaries/assembly/boundary_assembly.py:9-56
* ``split_boundaries``    <- _auxilliaries/boundary_information.py:54-93
* ``assemble`` / ``solve``<- main.py:181-231  (scipy COO->CSR, ``+``/``-`` with exact-zero pruning, spsolve)
* ``errors``              <- main.py:476-615 + _auxilliaries/error_assembly/*.py

The third-party pieces on the path (scipy.sparse construction and binops, scipy spsolve / SuperLU, numpy leggauss,
scipy roots_jacobi, numba JIT of the user callables) are not restated: the oracle calls the same installed
libraries the reference calls, so their semantics (duplicate summation, zero pruning, int32 indices) are inherited.
"""
from __future__ import annotations

import math
import types
import typing

import numpy as np
from numba import njit, f8
from numpy.polynomial.legendre import leggauss
from scipy.sparse import csr_matrix
from scipy.sparse.linalg import spsolve
from scipy.special import roots_jacobi

CHUNK = 1 << 15  # items (triangles / faces) per vectorised batch; bounds temporary memory at p=3


# --------------------------------------------------------------------------------------------------------------
# problem data
# --------------------------------------------------------------------------------------------------------------

def wrap_callables(advection=None, diffusion=None, reaction=None, forcing=None, dirichlet_bcs=None):
    """numba-compile the user callables with the signatures the reference uses (main.py:161-165)."""
    return types.SimpleNamespace(
        advection=njit(f8[:, :](f8[:, :]))(advection) if advection is not None else None,
        diffusion=njit(f8[:, :, :](f8[:, :]))(diffusion) if diffusion is not None else None,
        reaction=njit(f8[:](f8[:, :]))(reaction) if reaction is not None else None,
        forcing=njit(f8[:](f8[:, :]))(forcing) if forcing is not None else None,
        dirichlet_bcs=njit(f8[:](f8[:, :]))(dirichlet_bcs) if dirichlet_bcs is not None else None,
    )


def flatten_geometry(geometry) -> types.SimpleNamespace:
    """Turn a (reference or reyna_b200) DGFEMGeometry into flat arrays (geometry/two_dimensional/DGFEM.py:64-86)."""
    regions = [np.asarray(r, dtype=np.int64) for r in geometry.mesh.filtered_regions]
    poly_off = np.zeros(len(regions) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in regions], out=poly_off[1:])
    poly_vtx = np.concatenate(regions) if regions else np.zeros(0, dtype=np.int64)
    g = types.SimpleNamespace(
        nodes=np.ascontiguousarray(geometry.nodes, dtype=np.float64),
        n_elements=int(geometry.n_elements),
        tri=np.ascontiguousarray(geometry.subtriangulation, dtype=np.int64),
        tri_elem=np.asarray(geometry.triangle_to_polygon, dtype=np.int64),
        bbox=np.asarray(geometry.elem_bounding_boxes, dtype=np.float64).reshape(-1, 4),
        area=np.asarray(geometry.areas, dtype=np.float64),
        poly_off=poly_off, poly_vtx=poly_vtx,
        iedge=np.asarray(geometry.interior_edges, dtype=np.int64).reshape(-1, 2),
        ielem=np.asarray(geometry.interior_edges_to_element, dtype=np.int64).reshape(-1, 2),
        inorm=(np.asarray(geometry.interior_normals, dtype=np.float64).reshape(-1, 2)
               if geometry.interior_normals is not None else np.zeros((0, 2))),
        bedge=np.asarray(geometry.boundary_edges, dtype=np.int64).reshape(-1, 2),
        belem=np.asarray(geometry.boundary_edges_to_element, dtype=np.int64).reshape(-1),
        bnorm=np.asarray(geometry.boundary_normals, dtype=np.float64).reshape(-1, 2),
    )
    return g


def flatten_partition(part) -> types.SimpleNamespace:
    """Oracle geometry of a ``reyna_b200.partition.PartitionedGeometry``: its owned elements [0, n_rows) followed by the ghost
    elements across the cut faces, as a stand-alone mesh.  ``assemble`` on it yields, in block rows [0, n_rows), exactly the
    rows of the global matrix (local column numbering; ghost rows are incomplete and must be ignored): the test-suite and
    bench.py use it to check / time real slices of workloads that are too large for one oracle call."""
    f = part._rb_flat
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    return types.SimpleNamespace(
        nodes=np.ascontiguousarray(f.nodes, dtype=np.float64), n_elements=int(f.n_elem), n_rows=int(f.n_rows),
        tri=i64(f.tri), tri_elem=i64(f.tri_elem), bbox=np.asarray(f.bbox, dtype=np.float64).reshape(-1, 4),
        area=np.asarray(f.area, dtype=np.float64), poly_off=i64(f.poly_off), poly_vtx=i64(f.poly_vtx),
        iedge=i64(f.iedge).reshape(-1, 2), ielem=i64(f.ielem).reshape(-1, 2),
        inorm=np.asarray(f.inorm, dtype=np.float64).reshape(-1, 2), bedge=i64(f.bedge).reshape(-1, 2), belem=i64(f.belem),
        bnorm=np.asarray(f.bnorm, dtype=np.float64).reshape(-1, 2), elem_gid=i64(f.elem_gid))


# --------------------------------------------------------------------------------------------------------------
# quadrature and basis
# --------------------------------------------------------------------------------------------------------------

def quadrature(p: int):
    """(w_tri (Qv,), ref_tri (Qv,2), w_edge (Qe,), x_edge (Qe,)) -- main.py:233-257.

    Gauss-Legendre in the first direction, Gauss-Jacobi(1,0) in the second, x-major tensor grid, Duffy collapse
    onto the unit triangle; the weights still sum to 4 (the 1/8 is applied in the volume kernel)."""
    q = p + 1
    x, wx = leggauss(q)
    y, wy = roots_jacobi(q, 1, 0)
    qx = np.repeat(x, q)
    qy = np.tile(y, q)
    w = (wx[:, None] * wy).reshape(-1)
    shift = np.stack((0.5 * (1.0 + qx) * (1.0 - qy) - 1.0, qy), axis=1)
    ref = 0.5 * shift + 0.5
    return w, ref, wx.copy(), x.copy()


def basis_orders(p: int) -> np.ndarray:
    """(d,2) index pairs (i,j), i+j<=p, row-major in (i,j) -- polygonal_basis_utils.py:7-26."""
    if p < 0:
        raise ValueError('Input "polydegree" must be non-negative (>=0).')
    return np.array([(i, j) for i in range(p + 1) for j in range(p + 1) if i + j <= p], dtype=np.int64).reshape(-1, 2)


def jacobi_orthonormal(y: np.ndarray, alpha: int, nmax: int) -> np.ndarray:
    """Orthonormal Jacobi P^(alpha,alpha)_n(y), n=0..nmax, shape (nmax+1, M) -- polygonal_basis_utils.py:29-78."""
    a = float(alpha)
    out = np.empty((nmax + 1,) + y.shape, dtype=np.float64)
    gamma0 = 2.0 ** (2.0 * a + 1.0) / (2.0 * a + 1.0) * (math.gamma(a + 1.0) ** 2) / math.gamma(2.0 * a + 1.0)
    out[0] = 1.0 / np.sqrt(gamma0)
    if nmax == 0:
        return out
    gamma1 = ((a + 1.0) ** 2 / (2.0 * a + 3.0)) * gamma0
    out[1] = ((a + 1.0) / np.sqrt(gamma1)) * y
    a_old = 1.0 / (a + 1.0) * np.sqrt((a + 1.0) ** 2 / (2.0 * a + 3.0))
    for j in range(1, nmax):
        h1 = 2.0 * (j + a)
        a_new = 1.0 / (j + a + 1.0) * np.sqrt(
            (j + 1.0) * (j + 1.0 + 2.0 * a) * (j + 1.0 + a) ** 2 / ((h1 + 1.0) * (h1 + 3.0)))
        out[j + 1] = (y * out[j] - a_old * out[j - 1]) / a_new
        a_old = a_new
    return out


_CLAMP = 1.0 - 2.220446049250313e-16


def bbox_legendre(x: np.ndarray, m: np.ndarray, h: np.ndarray, p: int):
    """1-D bbox-scaled orthonormal Legendre values and derivatives, each (p+1, M) -- assembly_aux.py:32-66.

    ``x``, ``m``, ``h`` are broadcast-compatible (M,).  The argument is clamped to +-(1-eps) when it leaves [-1,1]."""
    y = (x - m) / h
    mask = np.abs(y) > 1.0
    if mask.any():
        y = np.where(mask, _CLAMP * np.sign(y), y)
    val = h ** (-0.5) * jacobi_orthonormal(y, 0, p)
    der = np.zeros_like(val)
    if p >= 1:
        corr = np.array([np.sqrt((n + 1.0) * n) for n in range(1, p + 1)])
        der[1:] = h ** (-1.5) * jacobi_orthonormal(y, 1, p - 1) * corr.reshape((-1,) + (1,) * y.ndim)
    return val, der


def basis_and_grad(X: np.ndarray, bbox: np.ndarray, orders: np.ndarray, need_grad: bool = True):
    """phi (d, M) and grad (d, M, 2) at points X (M,2) for per-point bounding boxes bbox (M,4) -- assembly_aux.py:69-133."""
    p = int(orders.max()) if orders.size else 0
    hx = 0.5 * (bbox[:, 1] - bbox[:, 0])
    hy = 0.5 * (bbox[:, 3] - bbox[:, 2])
    mx = 0.5 * (bbox[:, 1] + bbox[:, 0])
    my = 0.5 * (bbox[:, 3] + bbox[:, 2])
    vx, dx = bbox_legendre(X[:, 0], mx, hx, p)
    vy, dy = bbox_legendre(X[:, 1], my, hy, p)
    ox, oy = orders[:, 0], orders[:, 1]
    phi = vx[ox] * vy[oy]
    if not need_grad:
        return phi, None
    grad = np.empty(phi.shape + (2,), dtype=np.float64)
    grad[..., 0] = dx[ox] * vy[oy]
    grad[..., 1] = vx[ox] * dy[oy]
    return phi, grad


# --------------------------------------------------------------------------------------------------------------
# geometry of quadrature points
# --------------------------------------------------------------------------------------------------------------

def triangle_points(g, ref: np.ndarray, sl: slice):
    """Physical points (n, Qv, 2) and |det(0.5*[V1-V0; V2-V0])| (n,) -- full_assembly.py:40-45, assembly_aux.py:7-29."""
    V = g.nodes[g.tri[sl]]                                   # (n,3,2)
    bary = np.stack((1.0 - ref[:, 0] - ref[:, 1], ref[:, 0], ref[:, 1]), axis=1)   # (Qv,3)
    X = np.einsum('qk,nkc->nqc', bary, V)
    e1 = 0.5 * (V[:, 1] - V[:, 0])
    e2 = 0.5 * (V[:, 2] - V[:, 0])
    det = np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
    return X, det


def edge_points(nodes: np.ndarray, edges: np.ndarray, x1: np.ndarray):
    """mid (n,2), tanvec (n,2), X (n,Qe,2), De (n,) -- full_assembly.py:118-122."""
    P0, P1 = nodes[edges[:, 0]], nodes[edges[:, 1]]
    mid = (P0 + P1) / 2.0
    tan = 0.5 * (P1 - P0)
    X = x1[None, :, None] * tan[:, None, :] + mid[:, None, :]
    De = np.sqrt(tan[:, 0] ** 2 + tan[:, 1] ** 2)
    return mid, tan, X, De


def max_subtriangle_area(g, edges: np.ndarray, elems: np.ndarray) -> np.ndarray:
    """|k_b|_K = max_v 0.5*|(P1-P0) x (v-P0)| over the vertices v of element K -- full_assembly.py:142-143."""
    P0, P1 = g.nodes[edges[:, 0]], g.nodes[edges[:, 1]]
    e = P1 - P0
    out = np.zeros(edges.shape[0])
    cnt = g.poly_off[elems + 1] - g.poly_off[elems]
    for k in range(int(cnt.max()) if cnt.size else 0):
        live = k < cnt
        v = g.nodes[g.poly_vtx[np.where(live, g.poly_off[elems] + k, 0)]]
        w = v - P0
        cr = 0.5 * np.abs(e[:, 0] * w[:, 1] - e[:, 1] * w[:, 0])
        out = np.where(live, np.maximum(out, cr), out)
    return out


def _chunks(n: int):
    for s in range(0, n, CHUNK):
        yield slice(s, min(n, s + CHUNK))


# --------------------------------------------------------------------------------------------------------------
# the four assemblers
# --------------------------------------------------------------------------------------------------------------

def volume_terms(g, p: int, co):
    """Diagonal blocks Z (E,d,d) with Z[e,i,j] as in full_assembly.py:56-77 (already summed over the element's
    sub-triangles in triangle order, main.py:315) and the load (E,d).  B[(e,j),(e,i)] = Z[e,i,j]."""
    orders = basis_orders(p)
    d = orders.shape[0]
    w, ref, _, _ = quadrature(p)
    E = g.n_elements
    Z = np.zeros((E, d, d))
    Lf = np.zeros((E, d))
    for sl in _chunks(g.tri.shape[0]):
        X, det = triangle_points(g, ref, sl)
        n, Q = X.shape[0], X.shape[1]
        Xf = np.ascontiguousarray(X.reshape(-1, 2))
        el = g.tri_elem[sl]
        phi, grad = basis_and_grad(Xf, np.repeat(g.bbox[el], Q, axis=0), orders)
        phi = phi.reshape(d, n, Q)
        grad = grad.reshape(d, n, Q, 2)
        W = det[:, None] * w[None, :]                        # (n,Q)
        z = np.zeros((n, d, d))
        if co.diffusion is not None:
            a = co.diffusion(Xf).reshape(n, Q, 2, 2)
            ga = np.einsum('inqk,nqkl->inql', grad, a)
            z += np.einsum('inql,jnql->nij', ga, W[None, :, :, None] * grad)
        if co.advection is not None:
            b = co.advection(Xf).reshape(n, Q, 2)
            bg = np.sum(b[None] * grad, axis=-1)             # (d,n,Q)
            z += np.einsum('inq,jnq->nij', bg, phi * W[None])
        if co.reaction is not None:
            c = co.reaction(Xf).reshape(n, Q)
            z += np.einsum('inq,jnq->nij', c[None] * phi, W[None] * phi)
        z *= 0.5
        # sequential accumulation in triangle order, as main.py:315 does
        np.add.at(Z, el, z)
        if co.forcing is not None:
            f = co.forcing(Xf).reshape(n, Q)
            zf = 0.5 * np.sum(f[None] * phi * W[None], axis=-1)   # (d,n)
            np.add.at(Lf, el, zf.T)
    return Z, Lf


def interior_face_terms(g, p: int, co, sigma_D: float):
    """Per interior face the 2d x 2d matrix M (F,2d,2d) in B-orientation: M[f, r, c] is added to
    B[ind[r], ind[c]] with ind = [dofs(K1), dofs(K2)] -- full_assembly.py:80-214 + main.py:363-369."""
    orders = basis_orders(p)
    d = orders.shape[0]
    _, _, w1, x1 = quadrature(p)
    F = g.iedge.shape[0]
    M = np.zeros((F, 2 * d, 2 * d))
    for sl in _chunks(F):
        edges, el, nrm = g.iedge[sl], g.ielem[sl], g.inorm[sl]
        mid, tan, X, De = edge_points(g.nodes, edges, x1)
        n, Q = X.shape[0], X.shape[1]
        Xf = np.ascontiguousarray(X.reshape(-1, 2))
        phi1, g1 = basis_and_grad(Xf, np.repeat(g.bbox[el[:, 0]], Q, axis=0), orders)
        phi2, g2 = basis_and_grad(Xf, np.repeat(g.bbox[el[:, 1]], Q, axis=0), orders)
        phi1, phi2 = phi1.reshape(d, n, Q), phi2.reshape(d, n, Q)
        g1, g2 = g1.reshape(d, n, Q, 2), g2.reshape(d, n, Q, 2)
        Wq = w1[None, :] * np.ones((n, 1))                   # reference weights; De applied at the end (:214)
        blk = np.zeros((n, 2 * d, 2 * d))
        if co.diffusion is not None:
            a_mid = co.diffusion(np.ascontiguousarray(mid))
            lam = np.einsum('ni,nij,nj->n', nrm, a_mid, nrm)
            a1, a2 = g.area[el[:, 0]], g.area[el[:, 1]]
            kb1 = max_subtriangle_area(g, edges, el[:, 0])
            kb2 = max_subtriangle_area(g, edges, el[:, 1])
            ci1 = np.minimum(a1 / kb1, float(p * p))
            ci2 = np.minimum(a2 / kb2, float(p * p))
            sig = sigma_D * lam * p ** 2 * (2 * De) * np.maximum(ci1 / a1, ci2 / a2)
            a = co.diffusion(Xf).reshape(n, Q, 2, 2)
            F1 = np.einsum('nqjk,inqk,nj->inq', a, g1, nrm)  # n . (a grad phi)
            F2 = np.einsum('nqjk,inqk,nj->inq', a, g2, nrm)

            def mm(A, Bm):
                return np.einsum('rnq,cnq,nq->nrc', A, Bm, Wq)
            s = sig[:, None, None]
            blk[:, :d, :d] += s * mm(phi1, phi1) - 0.5 * mm(phi1, F1) - 0.5 * mm(F1, phi1)
            blk[:, :d, d:] += -s * mm(phi1, phi2) - 0.5 * mm(phi1, F2) + 0.5 * mm(F1, phi2)
            blk[:, d:, :d] += -s * mm(phi2, phi1) + 0.5 * mm(phi2, F1) - 0.5 * mm(F2, phi1)
            blk[:, d:, d:] += s * mm(phi2, phi2) + 0.5 * mm(phi2, F2) + 0.5 * mm(F2, phi2)
        if co.advection is not None:
            b_mid = co.advection(np.ascontiguousarray(mid))
            k1_upwind = np.sum(b_mid * nrm, axis=1) >= 1e-12          # full_assembly.py:194
            beta = np.sum(co.advection(Xf).reshape(n, Q, 2) * nrm[:, None, :], axis=2)   # (n,Q)
            Wb = Wq * beta
            m22 = np.einsum('rnq,cnq,nq->nrc', phi2, phi2, Wb)
            m21 = np.einsum('rnq,cnq,nq->nrc', phi2, phi1, Wb)
            m11 = np.einsum('rnq,cnq,nq->nrc', phi1, phi1, Wb)
            m12 = np.einsum('rnq,cnq,nq->nrc', phi1, phi2, Wb)
            up = k1_upwind[:, None, None]
            blk[:, d:, d:] += np.where(up, m22, 0.0)
            blk[:, d:, :d] += np.where(up, -m21, 0.0)
            blk[:, :d, :d] += np.where(up, 0.0, -m11)
            blk[:, :d, d:] += np.where(up, 0.0, m12)
        M[sl] = De[:, None, None] * blk
    return M


def dirichlet_face_terms(g, p: int, co, sigma_D: float, idx: np.ndarray):
    """Zmat (n,d,d), Zf (n,d) for the elliptic boundary faces ``idx`` (to be SUBTRACTED) -- boundary_assembly.py:59-137."""
    orders = basis_orders(p)
    d = orders.shape[0]
    _, _, w1, x1 = quadrature(p)
    idx = np.asarray(idx, dtype=np.int64)
    edges, el, nrm = g.bedge[idx], g.belem[idx], g.bnorm[idx]
    mid, tan, X, De = edge_points(g.nodes, edges, x1)
    n, Q = X.shape[0], X.shape[1]
    Xf = np.ascontiguousarray(X.reshape(-1, 2))
    W = De[:, None] * w1[None, :]
    lam = np.einsum('ni,nij,nj->n', nrm, co.diffusion(np.ascontiguousarray(mid)), nrm)
    area = g.area[el]
    kb = max_subtriangle_area(g, edges, el)
    cinv = np.minimum(area / kb, float(p * p))
    sig = sigma_D * lam * p ** 2 * (2 * De) * cinv / area
    gD = co.dirichlet_bcs(Xf).reshape(n, Q)
    a = co.diffusion(Xf).reshape(n, Q, 2, 2)
    an = np.einsum('nqjk,nk->nqj', a, nrm)                   # coe = a . n
    phi, grad = basis_and_grad(Xf, np.repeat(g.bbox[el], Q, axis=0), orders)
    phi, grad = phi.reshape(d, n, Q), grad.reshape(d, n, Q, 2)
    G = np.einsum('inqk,nqk->inq', grad, an)
    Z = (np.einsum('rnq,cnq,nq->nrc', G, phi, W) + np.einsum('rnq,cnq,nq->nrc', phi, G, W)
         - sig[:, None, None] * np.einsum('rnq,cnq,nq->nrc', phi, phi, W))
    Zf = np.einsum('nq,inq,nq->ni', gD, G - sig[None, :, None] * phi, W)
    return Z, Zf


def inflow_face_terms(g, p: int, co, idx: np.ndarray):
    """Zmat (n,d,d), Zf (n,d) for the inflow boundary faces ``idx`` (to be SUBTRACTED) -- boundary_assembly.py:9-56."""
    orders = basis_orders(p)
    d = orders.shape[0]
    _, _, w1, x1 = quadrature(p)
    idx = np.asarray(idx, dtype=np.int64)
    edges, el, nrm = g.bedge[idx], g.belem[idx], g.bnorm[idx]
    mid, tan, X, De = edge_points(g.nodes, edges, x1)
    n, Q = X.shape[0], X.shape[1]
    Xf = np.ascontiguousarray(X.reshape(-1, 2))
    beta = np.sum(co.advection(Xf).reshape(n, Q, 2) * nrm[:, None, :], axis=2)
    gD = co.dirichlet_bcs(Xf).reshape(n, Q)
    phi, _ = basis_and_grad(Xf, np.repeat(g.bbox[el], Q, axis=0), orders, need_grad=False)
    phi = phi.reshape(d, n, Q)
    Z = De[:, None, None] * np.einsum('rnq,cnq,nq->nrc', phi, phi, beta * w1[None, :])
    Zf = De[:, None] * np.einsum('inq,nq->ni', phi, beta * gD * w1[None, :])
    return Z, Zf


def split_boundaries(g, co, tol: float = 1e-12):
    """(elliptical_indecies | None, inflow_indecies | None) -- boundary_information.py:54-93."""
    mids = np.ascontiguousarray(0.5 * (g.nodes[g.bedge[:, 0]] + g.nodes[g.bedge[:, 1]]))
    ell = inf = None
    if co.diffusion is not None:
        hyp = np.einsum('ni,nij,nj->n', g.bnorm, co.diffusion(mids), g.bnorm) <= tol
        ell = np.flatnonzero(~hyp)
    if co.advection is not None:
        inf = np.flatnonzero(np.sum(co.advection(mids) * g.bnorm, axis=1) <= tol)
    return ell, inf


# --------------------------------------------------------------------------------------------------------------
# global assembly / solve (main.py:181-231)
# --------------------------------------------------------------------------------------------------------------

def _diag_block_coo(d: int, elems: np.ndarray):
    """COO (row, col) index arrays for d x d diagonal blocks stored as Z[e, i, j] -> B[(e,j),(e,i)]."""
    ii, jj = np.meshgrid(np.arange(d), np.arange(d), indexing='ij')      # Z[i,j]
    rows = elems[:, None, None] * d + jj[None]
    cols = elems[:, None, None] * d + ii[None]
    return rows.reshape(-1), cols.reshape(-1)


def assemble(geometry, p: int, advection=None, diffusion=None, reaction=None, forcing=None, dirichlet_bcs=None,
             sigma_D: float = 10.0, tol: float = 1e-12, mirror_index_quirk: bool = True, precompiled=None):
    """Return (B csr N x N, L csr N x 1, info) exactly as DGFEM.dgfem(solve=False) leaves them (main.py:203-225)."""
    co = precompiled if precompiled is not None else wrap_callables(advection, diffusion, reaction, forcing,
                                                                      dirichlet_bcs)
    g = geometry if hasattr(geometry, 'poly_off') else flatten_geometry(geometry)
    orders = basis_orders(p)
    d = orders.shape[0]
    E = g.n_elements
    N = d * E
    ell, inf = split_boundaries(g, co, tol)

    # volume: csr_matrix keeps all E*d*d entries (explicit zeros included) -- main.py:319-325
    Z, Lf = volume_terms(g, p, co)
    r, c = _diag_block_coo(d, np.arange(E))
    B = csr_matrix((Z.reshape(-1), (r, c)), shape=(N, N))
    L = csr_matrix((Lf.reshape(-1), (np.arange(N), np.zeros(N, dtype=int))), shape=(N, 1))

    # interior faces -- main.py:329-374 ; "+=" on csr is csr_plus_csr (prunes exact zeros)
    if g.iedge.shape[0] > 0:
        M = interior_face_terms(g, p, co, sigma_D)
        ind = np.concatenate((g.ielem[:, 0, None] * d + np.arange(d)[None], g.ielem[:, 1, None] * d + np.arange(d)[None]),
                             axis=1)                         # (F,2d)
        rows = np.repeat(ind[:, :, None], 2 * d, axis=2)
        cols = np.repeat(ind[:, None, :], 2 * d, axis=1)
        Bint = csr_matrix((M.reshape(-1), (rows.reshape(-1), cols.reshape(-1))), shape=(N, N))
    else:
        Bint = csr_matrix((N, N))
    B = B + Bint

    if co.diffusion is not None:
        Zd, Zfd = dirichlet_face_terms(g, p, co, sigma_D, ell)
        el = g.belem[ell]
        acc = np.zeros((E, d, d))
        np.add.at(acc, el, Zd)
        accf = np.zeros((E, d))
        np.add.at(accf, el, Zfd)
        # main.py:397-400 sets the COO indices only for the elements of the FIRST n_elliptic boundary edges
        covered = np.zeros(E, dtype=bool)
        covered[g.belem[:len(ell)]] = True
        if not mirror_index_quirk:
            covered[el] = True
        r, c = _diag_block_coo(d, np.arange(E))
        keep = np.repeat(covered, d * d)
        r, c = np.where(keep, r, 0), np.where(keep, c, 0)
        Bd = csr_matrix((acc.reshape(-1), (r, c)), shape=(N, N))
        B = B - Bd
        L = L - csr_matrix(accf.reshape(-1, 1))

    if co.advection is not None:
        Za, Zfa = inflow_face_terms(g, p, co, inf)
        el = g.belem[inf]
        r, c = _diag_block_coo(d, el)
        Ba = csr_matrix((Za.reshape(-1), (r, c)), shape=(N, N))
        accf = np.zeros((E, d))
        np.add.at(accf, el, Zfa)
        B = B - Ba
        L = L - csr_matrix(accf.reshape(-1, 1))

    info = types.SimpleNamespace(orders=orders, dim_elem=d, dim_system=N, elliptical_indecies=ell, inflow_indecies=inf)
    return B, L, info


def solve(B: csr_matrix, L: csr_matrix) -> np.ndarray:
    """main.py:231 -- scipy.sparse.linalg.spsolve (SuperLU, COLAMD)."""
    return spsolve(B, L)


# --------------------------------------------------------------------------------------------------------------
# errors (main.py:476-615)
# --------------------------------------------------------------------------------------------------------------

def errors(geometry, p: int, solution: np.ndarray, exact_solution, grad_exact_solution=None, div_advection=None,
           advection=None, diffusion=None, reaction=None, sigma_D: float = 10.0, tol: float = 1e-12,
           precompiled=None):
    """(l2, dg, h1 | None) -- main.py:476-615 with error_assembly/full_assembly.py and boundary_assembly.py."""
    co = precompiled if precompiled is not None else wrap_callables(advection, diffusion, reaction, None, None)
    g = geometry if hasattr(geometry, 'poly_off') else flatten_geometry(geometry)
    if (co.diffusion is None) != (grad_exact_solution is None):
        raise ValueError('Need to input both or neither "diffusion" and "grad_u_exact".')
    reaction_f = co.reaction if co.reaction is not None else (lambda x: np.zeros(x.shape[0]))
    if co.advection is None or div_advection is None:
        div_advection = lambda x: np.zeros(x.shape[0])      # noqa: E731
    orders = basis_orders(p)
    d = orders.shape[0]
    w, ref, w1, x1 = quadrature(p)
    U = np.asarray(solution, dtype=np.float64).reshape(g.n_elements, d)
    l2 = dg = h1 = 0.0

    for sl in _chunks(g.tri.shape[0]):
        X, det = triangle_points(g, ref, sl)
        n, Q = X.shape[0], X.shape[1]
        Xf = np.ascontiguousarray(X.reshape(-1, 2))
        el = g.tri_elem[sl]
        phi, grad = basis_and_grad(Xf, np.repeat(g.bbox[el], Q, axis=0), orders, need_grad=co.diffusion is not None)
        phi = phi.reshape(d, n, Q)
        W = det[:, None] * w[None, :]
        u = exact_solution(Xf).reshape(n, Q)
        c0 = (reaction_f(Xf) - 0.5 * div_advection(Xf)).reshape(n, Q)
        uh = np.einsum('inq,ni->nq', phi, U[el])
        t1 = (u - uh) ** 2
        t2 = 0.0
        if co.diffusion is not None:
            grad = grad.reshape(d, n, Q, 2)
            a = co.diffusion(Xf).reshape(n, Q, 2, 2)
            gu = grad_exact_solution(Xf).reshape(n, Q, 2)
            guh = np.einsum('inqk,ni->nqk', grad, U[el])
            e = gu - guh
            t4 = np.sum(e ** 2, axis=2)
            t2 = np.einsum('nqi,nqij,nqj->nq', e, a, e)
            h1 += 0.5 * np.sum(t4 * W)
        l2 += 0.5 * np.sum(t1 * W)
        dg += 0.5 * np.sum((t2 + c0 * t1) * W)

    for sl in _chunks(g.iedge.shape[0]):
        edges, el, nrm = g.iedge[sl], g.ielem[sl], g.inorm[sl]
        mid, tan, X, De = edge_points(g.nodes, edges, x1)
        n, Q = X.shape[0], X.shape[1]
        Xf = np.ascontiguousarray(X.reshape(-1, 2))
        phi1, _ = basis_and_grad(Xf, np.repeat(g.bbox[el[:, 0]], Q, axis=0), orders, need_grad=False)
        phi2, _ = basis_and_grad(Xf, np.repeat(g.bbox[el[:, 1]], Q, axis=0), orders, need_grad=False)
        u1 = np.einsum('inq,ni->nq', phi1.reshape(d, n, Q), U[el[:, 0]])
        u2 = np.einsum('inq,ni->nq', phi2.reshape(d, n, Q), U[el[:, 1]])
        t = (u1 - u2) ** 2
        W = De[:, None] * w1[None, :]
        if co.diffusion is not None:
            lam = np.einsum('ni,nij,nj->n', nrm, co.diffusion(np.ascontiguousarray(mid)), nrm)
            a1, a2 = g.area[el[:, 0]], g.area[el[:, 1]]
            ci1 = np.minimum(a1 / max_subtriangle_area(g, edges, el[:, 0]), float(p * p))
            ci2 = np.minimum(a2 / max_subtriangle_area(g, edges, el[:, 1]), float(p * p))
            sig = sigma_D * lam * p ** 2 * (2 * De) * np.maximum(ci1 / a1, ci2 / a2)
            dg += np.sum(sig[:, None] * t * W)
        if co.advection is not None:
            bn = np.abs(np.sum(co.advection(Xf).reshape(n, Q, 2) * nrm[:, None, :], axis=2))
            dg += 0.5 * np.sum(bn * t * W)

    ell, _ = split_boundaries(g, co, tol)
    if co.diffusion is not None and len(ell) > 0:
        edges, el, nrm = g.bedge[ell], g.belem[ell], g.bnorm[ell]
        mid, tan, X, De = edge_points(g.nodes, edges, x1)
        n, Q = X.shape[0], X.shape[1]
        Xf = np.ascontiguousarray(X.reshape(-1, 2))
        lam = np.einsum('ni,nij,nj->n', nrm, co.diffusion(np.ascontiguousarray(mid)), nrm)
        area = g.area[el]
        cinv = np.minimum(area / max_subtriangle_area(g, edges, el), float(p * p))
        sig = sigma_D * lam * p ** 2 * (2 * De) * cinv / area
        phi, _ = basis_and_grad(Xf, np.repeat(g.bbox[el], Q, axis=0), orders, need_grad=False)
        uh = np.einsum('inq,ni->nq', phi.reshape(d, n, Q), U[el])
        u = exact_solution(Xf).reshape(n, Q)
        dg += np.sum(sig[:, None] * (u - uh) ** 2 * De[:, None] * w1[None, :])
    if co.advection is not None and g.bedge.shape[0] > 0:
        edges, el, nrm = g.bedge, g.belem, g.bnorm
        mid, tan, X, De = edge_points(g.nodes, edges, x1)
        n, Q = X.shape[0], X.shape[1]
        Xf = np.ascontiguousarray(X.reshape(-1, 2))
        bn = np.abs(np.sum(co.advection(Xf).reshape(n, Q, 2) * nrm[:, None, :], axis=2))
        phi, _ = basis_and_grad(Xf, np.repeat(g.bbox[el], Q, axis=0), orders, need_grad=False)
        uh = np.einsum('inq,ni->nq', phi.reshape(d, n, Q), U[el])
        u = exact_solution(Xf).reshape(n, Q)
        dg += np.sum(0.5 * bn * (u - uh) ** 2 * De[:, None] * w1[None, :])

    if co.diffusion is None:
        return float(np.sqrt(l2)), float(np.sqrt(dg)), None
    return float(np.sqrt(l2)), float(np.sqrt(dg)), float(np.sqrt(h1))
