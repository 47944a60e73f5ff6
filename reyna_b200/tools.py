"""Post-processing of a DG solution on the device.

``evaluate`` is the reference's element-wise point evaluation (reyna/DGFEM/two_dimensional/tools.py:117-153, same
signature and semantics); ``evaluate_batch`` evaluates many (point, element) pairs in one launch, which is what plotting or
sampling a solution on a 1M-element mesh needs (the reference loops over elements in Python).  ``plot_DG`` (matplotlib) is
out of scope.
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
