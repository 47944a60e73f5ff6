"""Reference quadrature and basis index set (host side, tiny).

Mirrors reyna/DGFEM/two_dimensional/main.py:233-257 (``DGFEM._intialise_quadrature``) and
_auxilliaries/polygonal_basis_utils.py:7-26 (``Basis_index2D``).  The same numpy/scipy routines the reference calls
(``leggauss``, ``roots_jacobi``) are used, so the nodes and weights handed to the kernels are bit-identical.
"""
from __future__ import annotations

import numpy as np
from numpy.polynomial.legendre import leggauss
from scipy.special import roots_jacobi


def basis_index_2d(polydegree: int) -> np.ndarray:
    """(d, 2) int64 array of (i, j) with i + j <= p, row-major in (i, j)."""
    if polydegree < 0:
        raise ValueError('Input "polydegree" must be non-negative (>=0).')
    grid = np.indices((polydegree + 1, polydegree + 1)).reshape(2, -1).T
    return grid[grid.sum(axis=1) <= polydegree]


def reference_quadrature(polydegree: int):
    """Return ``(element_rule, edge_rule)`` in the reference's shapes:
    ``element_rule = (weights (Qv,1), points (Qv,2))`` on the unit triangle (weights sum to 4, Duffy-collapsed
    Gauss-Legendre x Gauss-Jacobi(1,0), x-major) and ``edge_rule = (weights (Qe,1), nodes (Qe,1))`` on [-1, 1]."""
    q = polydegree + 1
    x, w_x = leggauss(q)
    y, w_y = roots_jacobi(q, 1, 0)
    quad_x = np.repeat(x, q)[:, None]
    quad_y = np.tile(y, q)[:, None]
    weights = (w_x[:, None] * w_y).reshape(-1, 1)
    shifted = np.hstack((0.5 * (1.0 + quad_x) * (1.0 - quad_y) - 1.0, quad_y))
    ref_points = 0.5 * shifted + 0.5
    return (weights, ref_points), (w_x[:, None], x[:, None])
