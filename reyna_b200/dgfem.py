"""``DGFEM`` -- the reference's solver class (reyna/DGFEM/two_dimensional/main.py:29-636) with the assemble-and-solve
hot path executed on a B200 through ``libreyna_b200.so``.

Same public surface: ``DGFEM(geometry, polynomial_degree)``, ``add_data(advection, diffusion, reaction, forcing,
dirichlet_bcs)``, ``dgfem(solve, verbose)``, ``errors(...)``; same attributes (``B``, ``L``, ``solution``, ``orders``,
``dim_elem``, ``dim_system``, ``h``, ``sigma_D``, ``boundary_information``, ``element_reference_quadrature``,
``edge_reference_quadrature``); same exceptions and warnings.  ``geometry`` may be the reference's own
``DGFEMGeometry`` or ``reyna_b200.DGFEMGeometry`` -- only the attribute names matter.

What runs where
  host   quadrature rule, boundary classification, physical quadrature points, ONE batched evaluation of each user
         callable (numba-compiled with the reference's signatures, main.py:161-165), pinned upload;
  B200   adjacency/segmented sort, fused volume+face assembly straight into CSR, boundary terms, zero pruning,
         block-Jacobi Krylov solve.
``B`` and ``L`` are materialised as scipy matrices lazily (first attribute access) -- ``dgfem(solve=True)`` itself never
needs the matrix on the host.  There is no CPU fallback: without the CUDA library or a GPU, ``dgfem`` raises.
"""
from __future__ import annotations

import ctypes as C
import time
import typing
import warnings

import numpy as np
from numba import njit, f8
from scipy.sparse import csr_matrix

from . import _lib
from .boundary_information import BoundaryInformation
from . import sampling
from .geometry import flat_geometry
from .quadrature import basis_index_2d, reference_quadrature


def _ptr(t) -> typing.Optional[int]:
    return None if t is None else t.data_ptr()


class DGFEM:

    def __init__(self, geometry, polynomial_degree: int = 1):
        self.geometry = geometry
        self.h: float = geometry.h
        self.polydegree: int = polynomial_degree

        self.advection = None
        self.diffusion = None
        self.reaction = None
        self.forcing = None
        self.dirichlet_bcs = None

        self.boundary_information: typing.Optional[BoundaryInformation] = None
        self.sigma_D = 10.0

        self.orders = None
        self.dim_elem: typing.Optional[int] = None
        self.dim_system: typing.Optional[int] = None
        self.solution: typing.Optional[np.ndarray] = None

        self._B = None
        self._L = None
        self._dev = None           # device-resident system of the last dgfem() call

        self.element_reference_quadrature = None
        self.edge_reference_quadrature = None

        # Krylov replacement of spsolve (main.py:231)
        # 'preconditioner': 'auto' = multigrid ('mg') when diffusion is present, block Jacobi ('bj') otherwise
        # 'mg_hierarchy': 'agglomerated' (P1 spaces on groups of 4 elements; defaults mg_nu 2, mg_omega 0.8, mg_alpha 2.2) or
        # 'aggregation' (piecewise constants; defaults 3, 0.7, 1.8) -- set 'mg_nu' / 'mg_omega' / 'mg_alpha' to override
        self.solver_options = {'method': 'auto', 'rtol': 1e-14, 'max_iter': 200000, 'check_every': 10,
                               'preconditioner': 'auto', 'mg_hierarchy': 'agglomerated'}
        self.solve_info: typing.Optional[dict] = None
        self.timings: dict = {}
        self.device = None         # torch device; default: current CUDA device
        self.keep_device_inputs = False   # bench.py: keep the uploaded samples for kernel-only timing
        self._inputs = None
        self._side_streams = {}
        self._pinned = {}          # cached pinned staging buffers (host side of the coefficient upload)

    # ------------------------------------------------------------------------------------------------------------
    # data
    # ------------------------------------------------------------------------------------------------------------
    def add_data(self, advection=None, diffusion=None, reaction=None, forcing=None, dirichlet_bcs=None):
        """Register the PDE coefficients (main.py:115-179).  Callables take an (N,2) array and return (N,2), (N,2,2) or
        (N,) arrays and must be numba-compatible; they are compiled with the reference's signatures (plus ``nogil`` so
        that the batched evaluation can use host threads)."""
        opts = dict(nogil=True)
        self.advection = njit(f8[:, :](f8[:, :]), **opts)(advection) if advection is not None else None
        self.diffusion = njit(f8[:, :, :](f8[:, :]), **opts)(diffusion) if diffusion is not None else None
        self.reaction = njit(f8[:](f8[:, :]), **opts)(reaction) if reaction is not None else None
        self.forcing = njit(f8[:](f8[:, :]), **opts)(forcing) if forcing is not None else None
        self.dirichlet_bcs = njit(f8[:](f8[:, :]), **opts)(dirichlet_bcs) if dirichlet_bcs is not None else None

        if self.diffusion is not None and self.polydegree == 0:
            warnings.warn("The polynomial degree given is too low to run diffusion simulations -- adjusted to the "
                          "minimum viable option (p=1)")
            self.polydegree += 1

        if self.advection is None and self.diffusion is None:
            raise ValueError('No advection or diffusion specified.')

        if self.dirichlet_bcs is None:
            raise ValueError('Must have Dirichlet boundary conditions.')

        self._intialise_quadrature()
        self.boundary_information = self._define_boundary_information()

    def _intialise_quadrature(self) -> None:
        self.element_reference_quadrature, self.edge_reference_quadrature = reference_quadrature(self.polydegree)

    def _define_boundary_information(self, **kwargs) -> BoundaryInformation:
        info = BoundaryInformation(self.geometry, **kwargs)
        info.split_boundaries(self.advection, self.diffusion)
        return info

    # ------------------------------------------------------------------------------------------------------------
    # lazily materialised host matrices
    # ------------------------------------------------------------------------------------------------------------
    @property
    def B(self):
        if self._B is None and self._dev is not None:
            self._materialise()
        return self._B

    @B.setter
    def B(self, value):
        self._B = value

    @property
    def L(self):
        if self._L is None and self._dev is not None:
            self._materialise()
        return self._L

    @L.setter
    def L(self, value):
        self._L = value

    def _materialise(self) -> None:
        dv = self._dev
        N = self.dim_system
        t0 = time.perf_counter()
        indptr = dv['indptr'].cpu().numpy()
        indices = dv['indices'].cpu().numpy()
        data = dv['data'].cpu().numpy()
        load = dv['load'].cpu().numpy()
        self._B = csr_matrix((data, indices, indptr), shape=(N, dv.get('n_cols', N)))
        self._L = csr_matrix(load.reshape(-1, 1))       # zero rows pruned, as main.py:218/225 leave it
        self.timings['d2h_s'] = time.perf_counter() - t0

    # ------------------------------------------------------------------------------------------------------------
    # the hot path
    # ------------------------------------------------------------------------------------------------------------
    def dgfem(self, solve: bool = True, verbose: int = 0):
        """Assemble ``B`` and ``L`` on the GPU and (optionally) solve -- main.py:181-231."""
        assert 0 <= verbose <= 1, '"verbose" must be either 0 or 1.'

        self.orders = basis_index_2d(self.polydegree)
        self.dim_elem = int(np.shape(self.orders)[0])
        self.dim_system = self.dim_elem * self.geometry.n_elements

        _time = time.time()
        self._B = self._L = None
        self._assemble_device()
        if verbose == 1:
            print(f"Assembly: {time.time() - _time}")

        if solve:
            self.solution = self._solve_device()

    # -- host: coefficient sampling -----------------------------------------------------------------------------------
    def _staging(self, torch, key, shape):
        """Pinned host buffer (cached per DGFEM object) so that the upload is a true asynchronous DMA."""
        buf = self._pinned.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape):
            buf = torch.empty(shape, dtype=torch.float64, pin_memory=True)
            self._pinned[key] = buf
        return buf

    def _sample_coefficients(self, torch, dev, fg):
        """Evaluate every callable once on the batched quadrature points (q-major) and upload: each array is copied to
        the device asynchronously from pinned memory as soon as it is evaluated, so the DMA of one coefficient overlaps
        the host evaluation of the next.  Returns (device tensors, bytes uploaded)."""
        (w2, ref2), (w1, x1) = self.element_reference_quadrature, self.edge_reference_quadrature
        dev_t = {}
        nbytes = 0

        def put(key, fn, X, tail):
            nonlocal nbytes
            if X.shape[0] == 0:
                return
            M = X.shape[0]
            # one spare row: the volume kernel's bulk copies move 16-byte units (rb_coefs, include/reyna_b200.h)
            buf = self._staging(torch, key, (M + 1,) + tail)
            sampling.evaluate_batched(fn, X, tail, out=buf.numpy()[:M])
            buf.numpy()[M] = 0.0
            dev_t[key] = buf.to(dev, non_blocking=True)[:M]
            nbytes += M * int(np.prod(tail, dtype=np.int64)) * 8

        pts = sampling.cached_points(self.geometry, fg, self.element_reference_quadrature, self.edge_reference_quadrature)
        Xv = pts['Xs']                 # round order of reyna_b200.stream (what dg_volume_kernel reads)
        if self.diffusion is not None:
            put('vol_a', self.diffusion, Xv, (2, 2))
        if self.advection is not None:
            put('vol_b', self.advection, Xv, (2,))
        if self.reaction is not None:
            put('vol_c', self.reaction, Xv, ())
        if self.forcing is not None:
            put('vol_f', self.forcing, Xv, ())

        if fg.n_iface:
            mid, Xf = pts['mid_i'], pts['Xi']
            if self.diffusion is not None:
                put('if_a', self.diffusion, Xf, (2, 2))
                put('if_a_mid', self.diffusion, mid, (2, 2))
            if self.advection is not None:
                put('if_b', self.advection, Xf, (2,))
                b_mid = sampling.evaluate_batched(self.advection, mid, (2,))
                # b(mid) . n >= 1e-12 (full_assembly.py:194); written out per component: np.sum(.., axis=1) over two columns
                # costs 25 ns per face on one core (75 ms at 3 M faces), the same two products and one add cost 3 ns
                upw = ((b_mid[:, 0] * fg.inorm[:, 0] + b_mid[:, 1] * fg.inorm[:, 1]) >= 1e-12).astype(np.uint8)
                dev_t['if_upwind'] = torch.from_numpy(upw).to(dev)
                nbytes += upw.nbytes

        if fg.n_bface:
            mid, Xb = pts['mid_b'], pts['Xb']
            put('bf_g', self.dirichlet_bcs, Xb, ())
            if self.diffusion is not None:
                put('bf_a', self.diffusion, Xb, (2, 2))
                put('bf_a_mid', self.diffusion, mid, (2, 2))
            if self.advection is not None:
                put('bf_b', self.advection, Xb, (2,))
        return dev_t, nbytes

    def _boundary_lists(self, fg):
        """Per-boundary-face flags and the (element -> boundary faces) lists the boundary kernel walks."""
        nb = fg.n_bface
        flags = np.zeros(nb, dtype=np.uint8)
        ell = self.boundary_information.elliptical_indecies
        inf = self.boundary_information.inflow_indecies
        if self.diffusion is not None and ell is not None and len(ell):
            ell = np.asarray(ell, dtype=np.int64)
            flags[ell] |= 1
            # The reference sets the COO indices of the Dirichlet blocks only for the elements of the FIRST n_elliptic
            # boundary edges (main.py:397-400); blocks of other elements are summed into B[0,0].  Mirrored (bit 2).
            covered = np.zeros(fg.n_elem, dtype=bool)
            covered[fg.belem[:len(ell)]] = True
            spill = ell[~covered[fg.belem[ell]]]
            flags[spill] |= 4
        if self.advection is not None and inf is not None and len(inf):
            flags[np.asarray(inf, dtype=np.int64)] |= 2
        live = np.flatnonzero(flags & 3)
        order = live[np.argsort(fg.belem[live], kind='stable')]     # by element, faces ascending within an element
        elems = fg.belem[order]
        first = np.ones(elems.shape[0], dtype=bool)
        first[1:] = elems[1:] != elems[:-1]
        bnd_elem = elems[first].astype(np.int32)
        bnd_off = np.append(np.flatnonzero(first), elems.shape[0]).astype(np.int32)
        return flags, bnd_elem, bnd_off, order.astype(np.int32)

    # -- device ---------------------------------------------------------------------------------------------------------
    def _device_geometry(self, torch, dev, fg):
        cache = getattr(self.geometry, '_rb_dev', None)
        if cache is None:
            cache = {}
            try:
                self.geometry._rb_dev = cache
            except Exception:   # pragma: no cover
                pass
        key = str(dev)
        if key not in cache:
            names = ('nodes', 'tri', 'tri_elem', 'tri_off', 'bbox', 'area', 'poly_off', 'poly_vtx', 'iedge', 'ielem',
                     'inorm', 'bedge', 'belem', 'bnorm')
            cache[key] = {n: torch.from_numpy(getattr(fg, n)).to(dev) for n in names}
            gid = getattr(fg, 'elem_gid', None)      # partitions only: local -> global element id
            cache[key]['elem_gid'] = torch.from_numpy(gid).to(dev) if gid is not None else None
            vs = sampling.volume_stream_of(fg)       # volume rounds of dg_volume_kernel
            for n in ('tri_xy', 'vround_tri0', 'tri_seg', 'elem_seg0'):
                a = getattr(vs, n)
                cache[key][n] = torch.from_numpy(a).to(dev) if a is not None else None
        return cache[key]

    def _quad_struct(self) -> "_lib.RbQuad":
        q = _lib.RbQuad()
        q.p = self.polydegree
        (w2, ref2), (w1, x1) = self.element_reference_quadrature, self.edge_reference_quadrature
        for i in range(ref2.shape[0]):
            q.wv[i] = float(w2[i, 0])
            q.rv[i][0] = float(ref2[i, 0])
            q.rv[i][1] = float(ref2[i, 1])
        for i in range(x1.shape[0]):
            q.we[i] = float(w1[i, 0])
            q.xe[i] = float(x1[i, 0])
        return q

    def _assemble_device(self) -> None:
        torch = _lib.require_cuda()
        L = _lib.lib()
        dev = torch.device(self.device) if self.device is not None else torch.device('cuda', torch.cuda.current_device())
        p = self.polydegree
        if not 1 <= p <= 3:
            raise _lib.ReynaB200Error(f'polynomial degree {p} is outside the supported range 1..3')
        tm = self.timings = {}
        t0 = time.perf_counter()
        fg = flat_geometry(self.geometry)
        E, d = fg.n_elem, self.dim_elem
        N = E * d
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream()
            sp = C.c_void_p(stream.cuda_stream)
            gd = self._device_geometry(torch, dev, fg)
            tm['geometry_upload_s'] = time.perf_counter() - t0

            t1 = time.perf_counter()
            cd, nbytes = self._sample_coefficients(torch, dev, fg)
            flags, bnd_elem, bnd_off, bnd_face = self._boundary_lists(fg)
            bd = {k: torch.from_numpy(v).to(dev) for k, v in
                  dict(flags=flags, bnd_elem=bnd_elem, bnd_off=bnd_off, bnd_face=bnd_face).items()}
            bd['has_spill'] = bool((flags & 4).any())
            tm['h2d_bytes'] = int(nbytes)
            tm['coefficient_eval_upload_s'] = time.perf_counter() - t1

            self._dev = self._assemble_on_device(torch, L, dev, sp, fg, gd, cd, bd, tm)
            self._inputs = (fg, gd, cd, bd) if self.keep_device_inputs else None
            stream.synchronize()
        tm['assemble_total_s'] = time.perf_counter() - t0

    def _assemble_on_device(self, torch, L, dev, sp, fg, gd, cd, bd, tm):
        """Everything from the uploaded samples to the device-resident CSR system; also used by bench.py for the
        kernel-only timing (inputs already resident in HBM)."""
        p, d = self.polydegree, self.dim_elem
        E = fg.n_elem                              # owned + ghost elements
        R = getattr(fg, 'n_rows', E)               # owned elements = block rows assembled here
        N = R * d
        i32 = dict(dtype=torch.int32, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)

        vs = sampling.volume_stream_of(fg)
        g = _lib.RbGeom(n_nodes=fg.n_nodes, n_elem=E, n_rows=R, n_tri=fg.n_tri, n_iface=fg.n_iface, n_bface=fg.n_bface,
                        n_vrounds=vs.n_rounds, n_vsegs=vs.n_segs, **{k: _ptr(v) for k, v in gd.items()})
        co = _lib.RbCoefs(**{k: _ptr(cd.get(k)) for k in ('vol_a', 'vol_b', 'vol_c', 'vol_f', 'if_a', 'if_b', 'if_a_mid',
                                                          'if_upwind', 'bf_a', 'bf_b', 'bf_g', 'bf_a_mid')})
        q = self._quad_struct()

        has_diff = self.diffusion is not None
        # -- structure
        st = _lib.RbStructure()
        adj_off = torch.empty(R + 1, **i32)
        adj_item = torch.empty(max(1, 2 * fg.n_iface), **i32)
        adj_nb = torch.empty(max(1, 2 * fg.n_iface), **i32)
        grp_off = torch.empty(R + 1, **i32)
        counts = torch.empty(4, **i32)
        st.adj_off, st.adj_item, st.adj_nb, st.grp_off, st.counts = (_ptr(adj_off), _ptr(adj_item), _ptr(adj_nb), _ptr(grp_off),
                                                                    _ptr(counts))
        ws_bytes = L.reyna_b200_structure_workspace(R, fg.n_iface)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(L.reyna_b200_structure(C.byref(g), int(has_diff), _ptr(cd.get('if_upwind')), C.byref(st), _ptr(ws),
                                          ws_bytes, sp), 'reyna_b200_structure')
        nnz = d * d * (R + st.n_groups)
        if nnz >= 2 ** 31:
            raise _lib.ReynaB200Error(f'nnz = {nnz} does not fit the int32 CSR indices')

        # -- CSR arrays
        indptr = torch.empty(N + 1, **i32)
        indices = torch.empty(nnz, **i32)
        data = torch.empty(nnz, **f64)
        load = torch.empty(N, **f64)
        zero_count = torch.zeros(1, dtype=torch.int64, device=dev)
        # the index arrays only depend on the structure: they are written on a side stream while the (occupancy-limited,
        # 8 warps per SM) assembly kernel runs, and joined before the step ends
        main_stream = torch.cuda.current_stream()
        side = self._side_streams.get(str(dev))
        if side is None:
            side = self._side_streams[str(dev)] = torch.cuda.Stream(device=dev)
        side.wait_stream(main_stream)
        _lib.check(L.reyna_b200_csr_pattern(C.byref(g), C.byref(st), p, _ptr(indptr), _ptr(indices),
                                            C.c_void_p(side.cuda_stream)), 'reyna_b200_csr_pattern')
        indptr.record_stream(side)
        indices.record_stream(side)
        aws_bytes = L.reyna_b200_assemble_workspace(E, fg.n_iface, vs.n_segs, p)
        aws = torch.empty(aws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(L.reyna_b200_assemble(C.byref(g), C.byref(st), C.byref(co), C.byref(q), float(self.sigma_D),
                                         _ptr(data), _ptr(load), _ptr(zero_count), _ptr(aws), aws_bytes, sp),
                   'reyna_b200_assemble')
        n_bnd = int(bd['bnd_elem'].shape[0])
        spill = torch.zeros(max(1, n_bnd), **f64)
        if n_bnd:
            _lib.check(L.reyna_b200_boundary(C.byref(g), C.byref(st), C.byref(co), C.byref(q), float(self.sigma_D),
                                             n_bnd, _ptr(bd['bnd_elem']), _ptr(bd['bnd_off']), _ptr(bd['bnd_face']),
                                             _ptr(bd['flags']), _ptr(data), _ptr(load), _ptr(spill),
                                             _ptr(zero_count), sp), 'reyna_b200_boundary')
        if bd.get('has_spill', False):
            if getattr(fg, 'elem_gid', None) is not None:
                raise _lib.ReynaB200Error('partitioned assembly does not mirror the reference index quirk of main.py:397-400 '
                                          '(boundary edges that are only partly elliptic)')
            data[0] -= spill.sum()          # the reference's B[0,0] quirk (main.py:397-400)
            zero_count += (data[0] == 0).to(torch.int64)
        main_stream.wait_stream(side)
        blocked = True
        structural_data = data
        structural = (indptr, indices, data)        # the solver (and its multigrid hierarchy) works on the block pattern
        nz = int(zero_count.item())         # synchronises
        tm['structural_nnz'] = int(nnz)
        tm['exact_zeros'] = nz
        if nz > 0:
            # scipy's `+`/`-` drop exact zeros (main.py:211-225): compact to the reference's pattern
            cws_bytes = L.reyna_b200_compact_workspace(N)
            cws = torch.empty(cws_bytes, dtype=torch.uint8, device=dev)
            indptr2 = torch.empty(N + 1, **i32)
            _lib.check(L.reyna_b200_compact_count(N, _ptr(indptr), _ptr(data), _ptr(indptr2), _ptr(cws), cws_bytes, sp),
                       'reyna_b200_compact_count')
            nnz2 = int(indptr2[-1].item())
            indices2 = torch.empty(nnz2, **i32)
            data2 = torch.empty(nnz2, **f64)
            _lib.check(L.reyna_b200_compact_fill(N, _ptr(indptr), _ptr(indices), _ptr(data), _ptr(indptr2),
                                                 _ptr(indices2), _ptr(data2), sp), 'reyna_b200_compact_fill')
            indptr, indices, data, nnz = indptr2, indices2, data2, nnz2
            blocked = False
        tm['nnz'] = int(nnz)
        return dict(indptr=indptr, indices=indices, data=data, load=load, blocked=blocked, n=N, d=d, structural=structural,
                    bbox=gd['bbox'],
                    n_cols=int(getattr(self.geometry, 'n_global_elements', R)) * d,
                    structure=dict(adj_off=adj_off, adj_item=adj_item, adj_nb=adj_nb, grp_off=grp_off, n_items=st.n_items,
                                   n_groups=st.n_groups, n_dup_items=st.n_dup_items),
                    # ctypes argument blocks of the fused assembly launch (bench.py re-launches it alone for the roofline)
                    _assemble_call=(g, st, co, q, structural_data, load, zero_count, aws, aws_bytes))

    def _relaunch_assemble(self, dv, stream_ptr) -> None:
        """Launch the fused volume+face assembly kernel once more on the system of the last dgfem() (measurement aid)."""
        g, st, co, q, data, load, zero_count, aws, aws_bytes = dv['_assemble_call']
        _lib.check(_lib.lib().reyna_b200_assemble(C.byref(g), C.byref(st), C.byref(co), C.byref(q), float(self.sigma_D),
                                                  _ptr(data), _ptr(load), _ptr(zero_count), _ptr(aws), aws_bytes,
                                                  stream_ptr), 'reyna_b200_assemble')

    def _solve_device(self) -> np.ndarray:
        torch = _lib.require_cuda()
        L = _lib.lib()
        dv = self._dev
        N, d = dv['n'], dv['d']
        if dv.get('n_cols', N) != N:
            raise _lib.ReynaB200Error('a partitioned system is solved through reyna_b200.distributed (halo exchange), '
                                      'not through DGFEM.dgfem(solve=True) on one rank')
        dev = dv['data'].device
        so = self.solver_options
        method = so.get('method', 'auto')
        if method == 'auto':
            method = 'cg' if self.advection is None else 'bicgstab'
        opts = _lib.RbSolveOpts(method=0 if method == 'cg' else 1, max_iter=int(so.get('max_iter', 200000)),
                                rtol=float(so.get('rtol', 1e-14)), check_every=int(so.get('check_every', 10)))
        prec = so.get('preconditioner', 'auto')
        if prec == 'auto':
            prec = 'mg' if self.diffusion is not None else 'bj'
        info = _lib.RbSolveInfo()
        t0 = time.perf_counter()
        s_indptr, s_indices, s_data = dv['structural']
        mg = C.c_void_p(None)
        mg_levels = None
        with torch.cuda.device(dev):
            # a side stream: the library replays the multigrid application as a CUDA graph, which the legacy default
            # stream cannot capture
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            torch.cuda.set_stream(side)
            sp = C.c_void_p(side.cuda_stream)
            try:
                if prec == 'mg':
                    # 'agglomerated': P1 spaces on groups of 4 elements below the P1 level; 'aggregation': the
                    # piecewise-constant chain (the only one the distributed hierarchy has)
                    agglo = so.get('mg_hierarchy', 'agglomerated') == 'agglomerated'
                    mo = _lib.RbMgOpts(nu=int(so.get('mg_nu', 2 if agglo else 3)),
                                       omega=float(so.get('mg_omega', 0.8 if agglo else 0.7)),
                                       alpha=float(so.get('mg_alpha', 2.2 if agglo else 1.8)),
                                       elem_bbox=_ptr(dv['bbox']) if agglo else None)
                    _lib.check(L.reyna_b200_mg_create(N, d, self.polydegree, _ptr(s_indptr), _ptr(s_indices), _ptr(s_data),
                                                      C.byref(mo), None, C.byref(mg), sp), 'reyna_b200_mg_create')
                    nl = C.c_int32(0)
                    rows = (C.c_int32 * 16)()
                    L.reyna_b200_mg_info(mg, C.byref(nl), rows, None)
                    mg_levels = [int(rows[i]) for i in range(nl.value)]
                setup_s = time.perf_counter() - t0
                x = torch.zeros(N, dtype=torch.float64, device=dev)
                ws_bytes = L.reyna_b200_solve_workspace(N, 0, d, int(s_data.numel()))
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                _lib.check(L.reyna_b200_solve(N, 0, d, 1, _ptr(s_indptr), _ptr(s_indices), _ptr(s_data), _ptr(dv['load']),
                                              _ptr(x), C.byref(opts), C.byref(info), None, None, None, mg, _ptr(ws),
                                              ws_bytes, sp), 'reyna_b200_solve')
                side.synchronize()
                sol = x.cpu().numpy()
            finally:
                torch.cuda.set_stream(torch.cuda.default_stream(dev))
                if mg.value:
                    L.reyna_b200_mg_destroy(mg)
        self.solve_info = dict(method=method, iterations=int(info.iterations), rel_residual=float(info.rel_residual),
                               converged=bool(info.converged), restarts=int(info.restarts),
                               preconditioner=prec, mg_levels=mg_levels, setup_seconds=setup_s,
                               stagnated=bool(info.stagnated), seconds=time.perf_counter() - t0)
        self.timings['solve_s'] = self.solve_info['seconds']
        if not info.converged and not info.stagnated:
            warnings.warn(f"Krylov solver stopped after {info.iterations} iterations with relative residual "
                          f"{info.rel_residual:.3e}")
        return sol

    # ------------------------------------------------------------------------------------------------------------
    def errors(self, exact_solution, grad_exact_solution=None, div_advection=None):
        """(l2, dg, h1 | None) of the computed solution -- main.py:476-615."""
        if self.solution is None:
            raise ValueError('Need to run the .dgfem() method to generate a solution before calculating an error.')
        if (self.diffusion is None) != (grad_exact_solution is None):
            raise ValueError('Need to input both or neither "diffusion" and "grad_u_exact".')
        from .errors import dg_errors
        return dg_errors(self, exact_solution, grad_exact_solution, div_advection)

    def evaluate(self, points: np.ndarray, elements: np.ndarray) -> np.ndarray:
        """u_h at ``points[i]`` (which must lie in element ``elements[i]``), batched on the device -- the many-element form
        of tools.evaluate (tools.py:117-153)."""
        if self.solution is None:
            raise ValueError('Need to run the .dgfem() method to generate a solution before evaluating it.')
        from .tools import evaluate_batch
        return evaluate_batch(points, elements, self.solution, self.geometry.elem_bounding_boxes, self.polydegree,
                              device=self.device)

    def plot_DG(self):
        if self.solution is None:
            raise ValueError("Need to provide a solution using either the method '.dgfem' or an external function.")
        raise NotImplementedError('plotting is outside the reyna_b200 hot path (needs matplotlib)')
