"""Host side of the coefficient hand-off: physical quadrature points and ONE batched evaluation of each user callable.

The reference calls every coefficient callable once per sub-triangle / face on (p+1)^2 or p+1 points
(full_assembly.py:57,65,69,73,140,159,194-195; boundary_assembly.py:44-45,103,111-112).  Here the points of all items
are generated with exactly the reference's formulas (assembly_aux.py:7-29 for triangles, ``mid + x_q * tanvec`` for
edges, full_assembly.py:118-121), laid out q-major (``[q][item]``, the layout the kernels read coalesced), and each
callable is evaluated once over the whole array -- in row chunks on a thread pool, because the numba-compiled callables
are ``nogil`` and element-wise, so chunking does not change a single bit of the result.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_CHUNK = 1 << 17          # points per task
_pool = None
_pool_size = 0


def host_threads() -> int:
    env = os.environ.get("REYNA_B200_HOST_THREADS")
    if env:
        return max(1, int(env))
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:   # pragma: no cover
        n = os.cpu_count() or 1
    return max(1, min(n, 32))


def _get_pool(n: int) -> ThreadPoolExecutor:
    global _pool, _pool_size
    if _pool is None or _pool_size != n:
        if _pool is not None:
            _pool.shutdown(wait=False)
        _pool = ThreadPoolExecutor(max_workers=n, thread_name_prefix="rb-sample")
        _pool_size = n
    return _pool


def evaluate_batched(fn, X: np.ndarray, tail_shape: tuple, out: np.ndarray = None) -> np.ndarray:
    """``fn(X)`` for a C-contiguous (M,2) array, evaluated in chunks on the host thread pool into ``out`` (M,*tail)."""
    M = X.shape[0]
    if out is None:
        out = np.empty((M,) + tuple(tail_shape), dtype=np.float64)
    if M == 0:
        return out
    nthreads = host_threads()
    if M <= _CHUNK or nthreads == 1:
        out[...] = np.asarray(fn(X), dtype=np.float64).reshape(out.shape)
        return out
    bounds = list(range(0, M, _CHUNK)) + [M]

    def work(i):
        lo, hi = bounds[i], bounds[i + 1]
        out[lo:hi] = np.asarray(fn(X[lo:hi]), dtype=np.float64).reshape((hi - lo,) + tuple(tail_shape))

    list(_get_pool(nthreads).map(work, range(len(bounds) - 1)))
    return out


def volume_points(fg, ref2: np.ndarray) -> np.ndarray:
    """(Qv*T, 2) physical volume points, q-major: X[q,t] = (1-r0-r1) V0 + r0 V1 + r1 V2 (assembly_aux.py:7-29)."""
    V = fg.nodes[fg.tri]                                     # (T,3,2)
    ref2 = np.asarray(ref2)
    l0 = 1.0 - ref2[:, 0] - ref2[:, 1]
    Q, T = ref2.shape[0], V.shape[0]
    X = np.empty((Q, T, 2))
    for q in range(Q):          # same operation order as the matrix product [l0, r0, r1] @ V of the reference
        X[q] = l0[q] * V[:, 0, :] + ref2[q, 0] * V[:, 1, :] + ref2[q, 1] * V[:, 2, :]
    return X.reshape(-1, 2)


def edge_points(nodes: np.ndarray, edges: np.ndarray, x1: np.ndarray):
    """(mid (F,2), X (Qe*F,2) q-major) with X[q,f] = mid + x_q * tanvec, tanvec = (P1-P0)/2 (full_assembly.py:118-121)."""
    P0, P1 = nodes[edges[:, 0]], nodes[edges[:, 1]]
    mid = np.ascontiguousarray((P0 + P1) / 2.0)
    tan = 0.5 * (P1 - P0)
    x1 = np.asarray(x1).reshape(-1)
    X = np.empty((x1.shape[0], edges.shape[0], 2))
    for q in range(x1.shape[0]):
        X[q] = x1[q] * tan + mid
    return mid, X.reshape(-1, 2)
