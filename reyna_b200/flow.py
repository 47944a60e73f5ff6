"""Flow ordering of the elements for advection-reaction problems without diffusion (host side of the solver's block
Gauss-Seidel sweep, ``rb_solve_opts.sweep_*``).

Without diffusion an interior face only contributes rows to its DOWNWIND element (full_assembly.py:194-212: the upwind half of
the face matrix is exactly zero and scipy prunes it), so the block row of an element couples to its upwind neighbours only:
in a topological order of the upwind graph ``B`` is block lower triangular (hard part 1).  Levels:
``level(e) = 1 + max(level of the upwind neighbours)``; all elements of one level are independent.  Elements on cycles of
the flow (closed streamlines) never become free: the cycle is broken at the element with the fewest unresolved upwind
neighbours (those count as zero in the sweep), and the sweep is then a preconditioner only.
"""
from __future__ import annotations

import numpy as np
from numba import njit


@njit(cache=True)
def _kahn_levels(n_elem, up, down, level):
    """``up[f] -> down[f]`` edges.  Returns (number of levels, number of cycle breaks)."""
    F = up.shape[0]
    indeg = np.zeros(n_elem, dtype=np.int64)
    cnt = np.zeros(n_elem + 1, dtype=np.int64)
    for f in range(F):
        indeg[down[f]] += 1
        cnt[up[f] + 1] += 1
    off = np.cumsum(cnt)
    adj = np.empty(F, dtype=np.int64)
    pos = off[:-1].copy()
    for f in range(F):
        adj[pos[up[f]]] = down[f]
        pos[up[f]] += 1
    frontier = np.empty(n_elem, dtype=np.int64)
    nxt = np.empty(n_elem, dtype=np.int64)
    nf = 0
    for e in range(n_elem):
        level[e] = -1
        if indeg[e] == 0:
            frontier[nf] = e
            nf += 1
    lev = 0
    done = 0
    breaks = 0
    while True:
        if nf == 0:
            # every remaining element waits for another one: a cycle of the flow.  Break it at the remaining element with the
            # fewest unresolved upwind neighbours (its missing neighbours count as zero in the sweep) and go on.
            if done == n_elem or breaks >= 8192:
                break
            best = -1
            for e in range(n_elem):
                if level[e] < 0 and indeg[e] > 0 and (best < 0 or indeg[e] < indeg[best]):
                    best = e
                    if indeg[e] == 1:
                        break
            if best < 0:
                break
            indeg[best] = 0
            frontier[0] = best
            nf = 1
            breaks += 1
        nn = 0
        for i in range(nf):
            e = frontier[i]
            level[e] = lev
            for k in range(off[e], off[e + 1]):
                o = adj[k]
                if indeg[o] <= 0:
                    continue                     # released early (cycle break) or already placed
                indeg[o] -= 1
                if indeg[o] == 0:
                    nxt[nn] = o
                    nn += 1
        done += nf
        frontier, nxt = nxt, frontier
        nf = nn
        lev += 1
    return lev, breaks


def flow_levels(ielem: np.ndarray, upwind: np.ndarray, n_elem: int):
    """(elements sorted by level [n_elem] int32, level offsets [n_levels+1] int32, number of cycles that had to be broken).
    ``upwind[f] != 0`` <=> K1 = ielem[f,0] is the upwind element of interior face f (rb_coefs.if_upwind)."""
    ielem = np.asarray(ielem, dtype=np.int64).reshape(-1, 2)
    u = np.asarray(upwind).reshape(-1) != 0
    keep = (ielem[:, 0] < n_elem) & (ielem[:, 1] < n_elem)
    up = np.where(u, ielem[:, 0], ielem[:, 1])[keep]
    down = np.where(u, ielem[:, 1], ielem[:, 0])[keep]
    level = np.empty(n_elem, dtype=np.int64)
    nl, n_cyclic = _kahn_levels(n_elem, np.ascontiguousarray(up), np.ascontiguousarray(down), level)
    nl, n_cyclic = int(nl), int(n_cyclic)
    if np.any(level < 0):                     # more cycles than the break budget: the rest forms a last level
        level[level < 0] = nl
        nl += 1
    order = np.argsort(level, kind='stable').astype(np.int32)
    off = np.zeros(nl + 1, dtype=np.int32)
    np.cumsum(np.bincount(level, minlength=nl), out=off[1:])
    return order, off, n_cyclic
