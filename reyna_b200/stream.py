"""Volume stream of a flat geometry (host side of ``csrc/assemble_items.cu::dg_volume_kernel``).

The volume kernel runs one lane per sub-triangle.  Sub-triangles are packed, in their given order, into "volume rounds"
of WHOLE elements with at most 32 triangles (one warp); an element with more than 32 triangles fills rounds of its own and
is cut into "segments" of at most 32 triangles whose sums the face kernel adds up in order.  The coefficient samples of a
round -- its dominant input, 64 bytes per quadrature point at full ADR -- are ONE contiguous range of every sample array,
q-major inside the round: sample ``q`` of the ``l``-th triangle of a round with ``n`` triangles starting at triangle ``t0``
sits at index ``Qv*t0 + q*n + l``.  The kernel therefore fetches a round with one ``cp.async.bulk`` (TMA) copy per array and
the lanes of the warp read consecutive words of shared memory.

The order in which the user's callables see the points is internal (the reference calls them once per sub-triangle,
full_assembly.py:57-73; a batched call on a permuted point array returns the same values bit for bit).
"""
from __future__ import annotations

import types

import numpy as np
from numba import njit

WARP = 32
MAX_SEGMENTS = 8      # segments of one element the face kernel can add up (csrc/assemble_items.cu: kMaxSeg)


@njit(cache=True)
def _pack_rounds(tri_off, n_rows, round_tri0):
    """Greedy packing of whole elements into rounds of at most 32 triangles; ``round_tri0[r]`` = first triangle of round r."""
    nr = 0
    fill = WARP                               # forces a new round at the first element
    for e in range(n_rows):
        t0, t1 = tri_off[e], tri_off[e + 1]
        cnt = t1 - t0
        if cnt <= WARP:
            if fill + cnt > WARP:
                round_tri0[nr] = t0
                nr += 1
                fill = 0
            fill += cnt
        else:                                 # rounds of its own, 32 triangles each
            t = t0
            while t < t1:
                round_tri0[nr] = t
                nr += 1
                t += WARP
            fill = WARP
    return nr


def volume_stream(fg) -> types.SimpleNamespace:
    """Round tables of a flat geometry: ``vround_tri0`` [n_rounds+1], ``tri_n`` / ``tri_t0`` [n_tri] (width and first triangle of
    the round of every triangle), ``tri_xy`` [n_tri,3,2], and -- only when an element has more than 32 triangles --
    ``tri_seg`` [n_tri] and ``elem_seg0`` [n_rows+1] (segments = pieces of at most 32 triangles of one element)."""
    R = int(getattr(fg, 'n_rows', fg.n_elem))
    T = int(fg.n_tri)
    tri_off = np.ascontiguousarray(fg.tri_off[:R + 1], dtype=np.int64)
    cnt = np.diff(tri_off)
    if R and int(cnt.min()) <= 0:
        raise ValueError('every element needs at least one sub-triangle')
    if R and int(cnt.max()) > MAX_SEGMENTS * WARP:
        raise ValueError(f'an element has {int(cnt.max())} sub-triangles; the volume kernel handles at most {MAX_SEGMENTS * WARP}')
    round_tri0 = np.zeros(int(np.sum((cnt + WARP - 1) // WARP)) + 2, dtype=np.int64)
    nr = int(_pack_rounds(tri_off, R, round_tri0))
    round_tri0 = round_tri0[:nr + 1]
    round_tri0[nr] = T
    widths = np.diff(round_tri0)
    if nr and (int(widths.max()) > WARP or int(widths.min()) <= 0):      # pragma: no cover - guards the packing invariant
        raise AssertionError('invalid volume round')
    tri_round = np.repeat(np.arange(nr, dtype=np.int64), widths)
    out = types.SimpleNamespace(n_rounds=nr, vround_tri0=round_tri0.astype(np.int32), tri_n=widths[tri_round],
                                tri_t0=round_tri0[:-1][tri_round],
                                tri_xy=np.ascontiguousarray(fg.nodes[fg.tri[:T].astype(np.int64)], dtype=np.float64),
                                tri_seg=None, elem_seg0=None, n_segs=R)
    if cnt.size and int(cnt.max()) > WARP:
        elem_seg0 = np.zeros(R + 1, dtype=np.int64)
        np.cumsum((cnt + WARP - 1) // WARP, out=elem_seg0[1:])
        te = fg.tri_elem[:T].astype(np.int64)
        out.tri_seg = (elem_seg0[te] + (np.arange(T, dtype=np.int64) - tri_off[te]) // WARP).astype(np.int32)
        out.elem_seg0 = elem_seg0.astype(np.int32)
        out.n_segs = int(elem_seg0[-1])
    return out


def stream_positions(vs, Q: int) -> np.ndarray:
    """[n_tri, Q] index of every (sub-triangle, quadrature point) sample in the volume sample arrays."""
    T = vs.tri_n.shape[0]
    lane = np.arange(T, dtype=np.int64) - vs.tri_t0
    q = np.arange(Q, dtype=np.int64)[None, :]
    return vs.tri_t0[:, None] * Q + q * vs.tri_n[:, None] + lane[:, None]
