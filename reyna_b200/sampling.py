"""Host side of the coefficient hand-off: physical quadrature points and ONE batched evaluation of each user callable.

The reference calls every coefficient callable once per sub-triangle / face on (p+1)^2 or p+1 points
(full_assembly.py:57,65,69,73,140,159,194-195; boundary_assembly.py:44-45,103,111-112).  Here the points of all items
are generated with exactly the reference's formulas (assembly_aux.py:7-29 for triangles, ``mid + x_q * tanvec`` for
edges, full_assembly.py:118-121), laid out q-major (``[q][item]``; volume points in the round order of ``reyna_b200.stream``), and each
callable is evaluated once over the whole array -- in row chunks on a thread pool, because the numba-compiled callables
are ``nogil`` and element-wise, so chunking does not change a single bit of the result.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from numba import njit, prange

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
    # one process per GPU (torchrun): the ranks of a node evaluate their callables at the same time and share its cores
    try:
        n //= max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:   # pragma: no cover
        pass
    return max(1, min(n, 32))


def _get_pool(n: int) -> ThreadPoolExecutor:
    global _pool, _pool_size
    if _pool is None or _pool_size != n:
        if _pool is not None:
            _pool.shutdown(wait=False)
        _pool = ThreadPoolExecutor(max_workers=n, thread_name_prefix="rb-sample")
        _pool_size = n
    return _pool


_malloc_tuned = False


def _tune_malloc() -> None:
    """numba returns a fresh array from every call of a compiled callable; a chunk's result is above glibc's mmap threshold, so
    every call would mmap, page-fault and munmap its result.  ``REYNA_B200_MALLOPT=1`` raises the threshold so that the
    per-thread arenas stay warm (12 % less evaluation time on an 8-core container, within the run-to-run noise on the 16-core
    GPU box: opt-in, because it changes the allocator of the whole process).  Results are not affected."""
    global _malloc_tuned
    if _malloc_tuned:
        return
    _malloc_tuned = True
    if os.environ.get("REYNA_B200_MALLOPT", "0") == "0":
        return
    try:
        import ctypes
        libc = ctypes.CDLL("libc.so.6")
        libc.mallopt(-3, 64 * 1024 * 1024)     # M_MMAP_THRESHOLD
        libc.mallopt(-1, 1 << 30)              # M_TRIM_THRESHOLD
    except Exception:   # pragma: no cover  (not glibc)
        pass


def evaluate_batched(fn, X: np.ndarray, tail_shape: tuple, out: np.ndarray = None) -> np.ndarray:
    """``fn(X)`` for a C-contiguous (M,2) array, evaluated in chunks on the host thread pool into ``out`` (M,*tail)."""
    M = X.shape[0]
    if out is None:
        out = np.empty((M,) + tuple(tail_shape), dtype=np.float64)
    if M == 0:
        return out
    nthreads = host_threads()
    _tune_malloc()
    if M <= _CHUNK or nthreads == 1:
        out[...] = np.asarray(fn(X), dtype=np.float64).reshape(out.shape)
        return out
    bounds = list(range(0, M, _CHUNK)) + [M]

    def work(i):
        lo, hi = bounds[i], bounds[i + 1]
        out[lo:hi] = np.asarray(fn(X[lo:hi]), dtype=np.float64).reshape((hi - lo,) + tuple(tail_shape))

    list(_get_pool(nthreads).map(work, range(len(bounds) - 1)))
    return out


class EvalJob:
    """Pending evaluation of some callables over a range of points (``submit_terms``): ``wait()`` returns when every task has
    written its rows and re-raises the first exception a callable threw."""

    def __init__(self, futures):
        self._futures = futures

    def wait(self) -> None:
        for f in self._futures:
            f.result()
        self._futures = ()

    def abandon(self) -> None:
        """Drop the job after a failure elsewhere: tasks that have not started are cancelled, running ones are waited for (they
        write into buffers the next call reuses); their own exceptions are swallowed."""
        for f in self._futures:
            f.cancel()
        for f in self._futures:
            try:
                f.result()
            except BaseException:      # noqa: BLE001 - includes CancelledError
                pass
        self._futures = ()


def submit_terms(terms, X: np.ndarray, lo: int = 0, hi: int = None) -> EvalJob:
    """Queue ``out[lo:hi] = fn(X[lo:hi])`` for every ``(fn, tail_shape, out)`` of ``terms`` on the host thread pool and return at
    once.  The range is cut into tasks of at most ``_CHUNK`` points; a task applies ALL callables to its points while they are
    cache-resident.  Jobs are served in submission order, so a caller that submits the jobs of a whole assembly up front and
    then waits for them one by one keeps every host thread busy from the first face sample to the last volume sample -- there
    is no barrier at which threads idle (bit-identical to ``evaluate_batched``: element-wise functions)."""
    hi = X.shape[0] if hi is None else hi
    M = hi - lo
    if M <= 0 or not terms:
        return EvalJob(())
    nthreads = host_threads()
    _tune_malloc()

    def work(a, b):
        Xc = X[a:b]
        for fn, tail, out in terms:
            out[a:b] = np.asarray(fn(Xc), dtype=np.float64).reshape((b - a,) + tuple(tail))

    if nthreads == 1 or M <= _CHUNK // 4:
        work(lo, hi)                      # not worth a hand-off
        return EvalJob(())
    ntask = -(-M // _CHUNK)
    step = -(-M // ntask)
    pool = _get_pool(nthreads)
    return EvalJob([pool.submit(work, a, min(a + step, hi)) for a in range(lo, hi, step)])


def evaluate_many(terms, X: np.ndarray) -> None:
    """``out[...] = fn(X)`` for every ``(fn, tail_shape, out)`` of ``terms``, chunk by chunk on the host thread pool with ALL
    callables applied to a chunk of points while it is cache-resident (the point array is read from memory once instead of
    once per callable).  Bit-identical to ``evaluate_batched`` per callable (element-wise functions)."""
    M = X.shape[0]
    if M == 0 or not terms:
        return
    nthreads = host_threads()
    _tune_malloc()
    bounds = list(range(0, M, _CHUNK)) + [M]

    def work(i):
        lo, hi = bounds[i], bounds[i + 1]
        Xc = X[lo:hi]
        for fn, tail, out in terms:
            out[lo:hi] = np.asarray(fn(Xc), dtype=np.float64).reshape((hi - lo,) + tuple(tail))

    if len(bounds) == 2 or nthreads == 1:
        for i in range(len(bounds) - 1):
            work(i)
        return
    list(_get_pool(nthreads).map(work, range(len(bounds) - 1)))


@njit(parallel=True, cache=True, nogil=True)
def _volume_points_kernel(nodes, tri, ref2, out):
    Q, T = ref2.shape[0], tri.shape[0]
    for t in prange(T):
        i0, i1, i2 = tri[t, 0], tri[t, 1], tri[t, 2]
        x0, y0 = nodes[i0, 0], nodes[i0, 1]
        x1, y1 = nodes[i1, 0], nodes[i1, 1]
        x2, y2 = nodes[i2, 0], nodes[i2, 1]
        for q in range(Q):
            r0, r1 = ref2[q, 0], ref2[q, 1]
            l0 = 1.0 - r0 - r1
            out[q, t, 0] = l0 * x0 + r0 * x1 + r1 * x2
            out[q, t, 1] = l0 * y0 + r0 * y1 + r1 * y2


@njit(parallel=True, cache=True, nogil=True)
def _edge_points_kernel(nodes, edges, x1, mid, out):
    Q, F = x1.shape[0], edges.shape[0]
    for f in prange(F):
        ax, ay = nodes[edges[f, 0], 0], nodes[edges[f, 0], 1]
        bx, by = nodes[edges[f, 1], 0], nodes[edges[f, 1], 1]
        mx, my = (ax + bx) / 2.0, (ay + by) / 2.0
        tx, ty = 0.5 * (bx - ax), 0.5 * (by - ay)
        mid[f, 0] = mx
        mid[f, 1] = my
        for q in range(Q):
            out[q, f, 0] = x1[q] * tx + mx
            out[q, f, 1] = x1[q] * ty + my


@njit(parallel=True, cache=True, nogil=True)
def _volume_stream_points_kernel(nodes, tri, ref2, tri_t0, tri_n, out):
    Q, T = ref2.shape[0], tri.shape[0]
    for t in prange(T):
        i0, i1, i2 = tri[t, 0], tri[t, 1], tri[t, 2]
        x0, y0 = nodes[i0, 0], nodes[i0, 1]
        x1, y1 = nodes[i1, 0], nodes[i1, 1]
        x2, y2 = nodes[i2, 0], nodes[i2, 1]
        n = tri_n[t]
        base = tri_t0[t] * Q + (t - tri_t0[t])
        for q in range(Q):
            r0, r1 = ref2[q, 0], ref2[q, 1]
            l0 = 1.0 - r0 - r1
            out[base + q * n, 0] = l0 * x0 + r0 * x1 + r1 * x2
            out[base + q * n, 1] = l0 * y0 + r0 * y1 + r1 * y2


def volume_stream_points(fg, vs, ref2: np.ndarray) -> np.ndarray:
    """(Qv*T, 2) physical volume points in the round order of ``reyna_b200.stream`` (the order of the volume samples of
    ``rb_coefs``); same formula as ``volume_points``: X = (1-r0-r1) V0 + r0 V1 + r1 V2 (assembly_aux.py:7-29)."""
    ref2 = np.ascontiguousarray(ref2, dtype=np.float64)
    Q = ref2.shape[0]
    out = np.empty((fg.n_tri * Q, 2))
    _set_numba_threads()
    _volume_stream_points_kernel(fg.nodes, fg.tri, ref2, vs.tri_t0, vs.tri_n, out)
    return out


def volume_points(fg, ref2: np.ndarray) -> np.ndarray:
    """(Qv*T, 2) physical volume points, q-major: X[q,t] = (1-r0-r1) V0 + r0 V1 + r1 V2 (assembly_aux.py:7-29)."""
    ref2 = np.ascontiguousarray(ref2, dtype=np.float64)
    out = np.empty((ref2.shape[0], fg.tri.shape[0], 2))
    _set_numba_threads()
    _volume_points_kernel(fg.nodes, fg.tri, ref2, out)
    return out.reshape(-1, 2)


def edge_points(nodes: np.ndarray, edges: np.ndarray, x1: np.ndarray):
    """(mid (F,2), X (Qe*F,2) q-major) with X[q,f] = mid + x_q * tanvec, tanvec = (P1-P0)/2 (full_assembly.py:118-121)."""
    x1 = np.ascontiguousarray(np.asarray(x1).reshape(-1), dtype=np.float64)
    mid = np.empty((edges.shape[0], 2))
    out = np.empty((x1.shape[0], edges.shape[0], 2))
    _set_numba_threads()
    _edge_points_kernel(nodes, np.ascontiguousarray(edges), x1, mid, out)
    return mid, out.reshape(-1, 2)


def _set_numba_threads() -> None:
    import numba
    n = min(host_threads(), numba.config.NUMBA_NUM_THREADS)
    if numba.get_num_threads() != n:
        numba.set_num_threads(n)


def volume_stream_of(fg):
    """The volume-round tables of a flat geometry (``reyna_b200.stream.volume_stream``), cached on it."""
    vs = getattr(fg, '_vs', None)
    if vs is None:
        from .stream import volume_stream
        vs = volume_stream(fg)
        fg._vs = vs
    return vs


class _LazyPoints(dict):
    """Point cache of one quadrature rule; the q-major volume points ``Xv`` (only ``errors()`` reads them) are generated on
    first use."""

    def __init__(self, fg, ref2):
        super().__init__()
        self._fg, self._ref2 = fg, ref2

    def __missing__(self, key):
        if key != 'Xv':
            raise KeyError(key)
        self['Xv'] = volume_points(self._fg, self._ref2)
        return self['Xv']


def cached_points(geometry, fg, element_rule, edge_rule) -> dict:
    """Physical quadrature points of a geometry for one quadrature rule: volume ``Xs`` (round order, what the assembly
    samples) and ``Xv`` (q-major, lazily, for ``errors()``), interior faces ``(mid_i, Xi)``,
    boundary faces ``(mid_b, Xb)``.  They depend on the mesh and the polynomial degree only, so they are computed once
    and kept on the geometry object (repeated assemblies with changing coefficients pay for the callables only)."""
    ref2 = np.asarray(element_rule[1])
    x1 = np.asarray(edge_rule[1]).reshape(-1)
    key = (ref2.shape[0], x1.shape[0])
    cache = getattr(geometry, '_rb_points', None)
    if cache is None:
        cache = {}
        try:
            geometry._rb_points = cache
        except Exception:   # pragma: no cover
            pass
    pts = cache.get(key)
    if pts is None:
        pts = _LazyPoints(fg, ref2)
        pts['Xs'] = volume_stream_points(fg, volume_stream_of(fg), ref2)
        pts['mid_i'], pts['Xi'] = edge_points(fg.nodes, fg.iedge, x1) if fg.n_iface else (np.zeros((0, 2)), np.zeros((0, 2)))
        pts['mid_b'], pts['Xb'] = edge_points(fg.nodes, fg.bedge, x1) if fg.n_bface else (np.zeros((0, 2)), np.zeros((0, 2)))
        cache[key] = pts
    return pts
