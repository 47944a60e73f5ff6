"""``DGFEM.errors`` on the device -- L2, DG and H1 (semi-)norms of ``u - u_h`` (main.py:476-615).

The exact solution, its gradient, ``div(advection)`` and the PDE coefficients are evaluated once on the batched
quadrature points (host, as in ``dgfem``); the integrals are computed by ``reyna_b200_errors`` (csrc/errors.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, sampling
from .geometry import flat_geometry


def dg_errors(dg, exact_solution, grad_exact_solution=None, div_advection=None):
    torch = _lib.require_cuda()
    L = _lib.lib()
    fg = flat_geometry(dg.geometry)
    p = dg.polydegree
    dev = torch.device(dg.device) if dg.device is not None else torch.device('cuda', torch.cuda.current_device())

    # main.py:512-518 -- the reference replaces a missing reaction by zeros (and keeps it: a documented side effect)
    if dg.reaction is None:
        dg.reaction = lambda x: np.zeros(x.shape[0])
    if dg.advection is None or div_advection is None:
        div_advection = None

    (w2, ref2), (w1, x1) = dg.element_reference_quadrature, dg.edge_reference_quadrature
    host = {}
    pts = sampling.cached_points(dg.geometry, fg, dg.element_reference_quadrature, dg.edge_reference_quadrature)
    Xv = pts['Xv']
    host['vol_u'] = sampling.evaluate_batched(exact_solution, Xv, ())
    c0 = sampling.evaluate_batched(dg.reaction, Xv, ())
    if div_advection is not None:
        c0 = c0 - 0.5 * sampling.evaluate_batched(div_advection, Xv, ())
    host['vol_c0'] = c0
    if dg.diffusion is not None:
        host['vol_gu'] = sampling.evaluate_batched(grad_exact_solution, Xv, (2,))
        host['vol_a'] = sampling.evaluate_batched(dg.diffusion, Xv, (2, 2))
    if fg.n_iface:
        mid, Xf = pts['mid_i'], pts['Xi']
        if dg.diffusion is not None:
            host['if_a_mid'] = sampling.evaluate_batched(dg.diffusion, mid, (2, 2))
        if dg.advection is not None:
            host['if_b'] = sampling.evaluate_batched(dg.advection, Xf, (2,))
    if fg.n_bface:
        mid, Xb = pts['mid_b'], pts['Xb']
        host['bf_u'] = sampling.evaluate_batched(exact_solution, Xb, ())
        if dg.diffusion is not None:
            host['bf_a_mid'] = sampling.evaluate_batched(dg.diffusion, mid, (2, 2))
            flags = np.zeros(fg.n_bface, dtype=np.uint8)
            ell = dg.boundary_information.elliptical_indecies
            if ell is not None and len(ell):
                flags[np.asarray(ell, dtype=np.int64)] = 1
            host['bf_flags'] = flags
        if dg.advection is not None:
            host['bf_b'] = sampling.evaluate_batched(dg.advection, Xb, (2,))

    with torch.cuda.device(dev):
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        gd = dg._device_geometry(torch, dev, fg)
        dt = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in host.items()}
        g = _lib.RbGeom(n_nodes=fg.n_nodes, n_elem=fg.n_elem, n_rows=fg.n_elem, n_tri=fg.n_tri, n_iface=fg.n_iface,
                        n_bface=fg.n_bface, **{k: (v.data_ptr() if v is not None else None) for k, v in gd.items()})
        es = _lib.RbErrSamples(**{k: (dt[k].data_ptr() if k in dt else None)
                                  for k in ('vol_u', 'vol_gu', 'vol_a', 'vol_c0', 'if_a_mid', 'if_b', 'bf_u', 'bf_a_mid',
                                            'bf_b', 'bf_flags')})
        q = dg._quad_struct()
        sol = torch.from_numpy(np.ascontiguousarray(dg.solution, dtype=np.float64)).to(dev)
        out = torch.zeros(3, dtype=torch.float64, device=dev)
        ws_bytes = L.reyna_b200_errors_workspace()
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(L.reyna_b200_errors(C.byref(g), C.byref(q), C.byref(es), float(dg.sigma_D), sol.data_ptr(),
                                       out.data_ptr(), ws.data_ptr(), ws_bytes, sp), 'reyna_b200_errors')
        l2, dgn, h1 = (float(v) for v in out.cpu().numpy())
    if dg.diffusion is None:
        return float(np.sqrt(l2)), float(np.sqrt(dgn)), None
    return float(np.sqrt(l2)), float(np.sqrt(dgn)), float(np.sqrt(h1))
