"""Post-processing of a DG solution on the device.

``evaluate`` is the reference's element-wise point evaluation (reyna/DGFEM/two_dimensional/tools.py:117-153, same
signature and semantics); ``evaluate_batch`` evaluates many (point, element) pairs in one launch, which is what plotting or
sampling a solution on a 1M-element mesh needs (the reference loops over elements in Python).  ``plot_data`` prepares
everything ``plot_DG`` (tools.py:15-114) draws -- the lifted polygons, their face colours and the axis limits -- in one batched
evaluation; ``plot_DG`` itself hands that to matplotlib when it is installed (drawing is outside the hot path).
"""
from __future__ import annotations

import ctypes as C
import typing

import numpy as np

from . import _lib
from .quadrature import basis_index_2d


def _degree_from_dim(dim_elem: int) -> int:
    # tools.py:143-144: p from d = (p+1)(p+2)/2
    return int(round(0.5 * (-3.0 + np.sqrt(1.0 + 8.0 * dim_elem))))


def evaluate_batch(points: np.ndarray, elements: np.ndarray, solution: np.ndarray, bounding_boxes, polynomial_degree: int,
                   device=None) -> np.ndarray:
    """u_h(points[i]) on element ``elements[i]`` for all i; ``bounding_boxes`` is ``geometry.elem_bounding_boxes``."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
    el = np.ascontiguousarray(elements, dtype=np.int32).reshape(-1)
    if pts.shape[0] != el.shape[0]:
        raise ValueError('points and elements must have the same length')
    bbox = np.ascontiguousarray(np.asarray(bounding_boxes, dtype=np.float64).reshape(-1, 4))
    if el.size and (el.min() < 0 or el.max() >= bbox.shape[0]):
        raise ValueError('element index out of range')
    with torch.cuda.device(dev):
        t = {k: torch.from_numpy(v).to(dev) for k, v in dict(pts=pts, el=el, bbox=bbox,
             sol=np.ascontiguousarray(solution, dtype=np.float64).reshape(-1)).items()}
        out = torch.empty(pts.shape[0], dtype=torch.float64, device=dev)
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.reyna_b200_evaluate(int(polynomial_degree), pts.shape[0], t['bbox'].data_ptr(), t['pts'].data_ptr(),
                                         t['el'].data_ptr(), t['sol'].data_ptr(), out.data_ptr(), sp), 'reyna_b200_evaluate')
        return out.cpu().numpy()


def evaluate(x: np.ndarray, dg_coefficients: np.ndarray, bounding_box: np.ndarray,
             polynomial_degree: typing.Optional[int] = None, orders: typing.Optional[np.ndarray] = None) -> np.ndarray:
    """Element-wise evaluation with the reference's signature (tools.py:117-153): ``x`` (M,2) points of ONE element,
    ``dg_coefficients`` its d coefficients, ``bounding_box`` its [xmin, xmax, ymin, ymax]."""
    coefs = np.asarray(dg_coefficients, dtype=np.float64).reshape(-1)
    if orders is not None:
        polynomial_degree = int(np.max(np.sum(np.asarray(orders), axis=1)))
        if np.asarray(orders).shape != basis_index_2d(polynomial_degree).shape or \
                not np.array_equal(np.asarray(orders), basis_index_2d(polynomial_degree)):
            raise ValueError('orders must be the total-degree index set of Basis_index2D')
    elif polynomial_degree is None:
        polynomial_degree = _degree_from_dim(coefs.shape[0])
    x = np.asarray(x, dtype=np.float64).reshape(-1, 2)
    return evaluate_batch(x, np.zeros(x.shape[0], dtype=np.int32), coefs, np.asarray(bounding_box).reshape(1, 4),
                          int(polynomial_degree))


def plot_data(numerical_solution: np.ndarray, geometry, poly_degree: int, device=None) -> dict:
    """The data ``plot_DG`` (tools.py:15-114) draws, for all elements at once: ``offsets`` [E+1] / ``xyz`` [sum of polygon
    sizes, 3] = every polygon lifted to its solution values at its vertices (tools.py:66-89), ``facecolor`` [E,3] =
    (|m|/s, |m|/s, 1-|m|/s) with m the mean vertex value of the element and s = max(|U_max|, |U_min|) (tools.py:91-99), and the
    axis limits ``xlim``, ``ylim``, ``zlim`` (tools.py:101-103).  p = 0 follows tools.py:38-59 (one value per element, colour
    from the global mean)."""
    from .mesh import flatten_regions
    sol = np.asarray(numerical_solution, dtype=np.float64).reshape(-1)
    off, idx = flatten_regions(geometry.mesh.filtered_regions)
    off = np.asarray(off, dtype=np.int64)
    idx = np.asarray(idx, dtype=np.int64)
    nodes = np.asarray(geometry.nodes, dtype=np.float64)
    E = int(geometry.n_elements)
    cnt = np.diff(off)
    elem = np.repeat(np.arange(E, dtype=np.int64), cnt)
    xy = nodes[idx]
    if poly_degree == 0:
        u = sol[elem]
        u_max = float(np.max(sol))
        u_mean = float(np.mean(sol))
        c = abs(u_mean / u_max)
        face = np.tile(np.array([c, c, 1.0 - c]), (E, 1))
        zlim = (None, None)                                   # tools.py:59: auto-scaled
    else:
        u = evaluate_batch(xy, elem, sol, geometry.elem_bounding_boxes, poly_degree, device=device)
        u_min, u_max = float(np.min(u)), float(np.max(u))
        mean = np.add.reduceat(u, off[:-1]) / cnt
        scaling = max(abs(u_max), abs(u_min))
        c = np.abs(mean / scaling)
        face = np.column_stack((c, c, 1.0 - c))
        zlim = (u_min, u_max)
    return dict(offsets=off, xyz=np.column_stack((xy, u)), facecolor=face,
                xlim=(float(xy[:, 0].min()), float(xy[:, 0].max())), ylim=(float(xy[:, 1].min()), float(xy[:, 1].max())),
                zlim=zlim)


def plot_DG(numerical_solution: np.ndarray, geometry, poly_degree: int) -> None:
    """The reference's plot (tools.py:15-114) from ``plot_data``; needs matplotlib (not a dependency of the hot path)."""
    try:
        import matplotlib.pyplot as plt
        from mpl_toolkits.mplot3d.art3d import Poly3DCollection
    except ImportError as exc:      # pragma: no cover - matplotlib is absent from the build image
        raise ImportError('plot_DG needs matplotlib; reyna_b200.tools.plot_data returns the prepared arrays') from exc
    d = plot_data(numerical_solution, geometry, poly_degree)   # pragma: no cover
    fig, ax = plt.subplots(subplot_kw={'projection': '3d'})     # pragma: no cover
    off, xyz = d['offsets'], d['xyz']                           # pragma: no cover
    for t in range(off.shape[0] - 1):                           # pragma: no cover
        poly = Poly3DCollection([xyz[off[t]:off[t + 1]]])
        poly.set_facecolor(d['facecolor'][t])
        poly.set_edgecolor('k')
        ax.add_collection3d(poly)
    ax.set_xlim(*d['xlim']); ax.set_ylim(*d['ylim']); ax.set_zlim(*d['zlim'])   # pragma: no cover
    ax.set_xlabel(r'$x$'); ax.set_ylabel(r'$y$'); ax.set_zlabel(r'$u$')         # pragma: no cover
    plt.show()                                                                  # pragma: no cover
