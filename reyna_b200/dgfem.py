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
import os
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
        self._geo = geometry       # geometry the device path works on: `geometry`, or its locality renumbering (solve=True)
        self._perm = None          # element permutation (new -> old) of `_geo`, None = identity
        self.renumber = 'auto'     # 'auto': solve on a Morton renumbering when the element numbering has no locality; False: never

        self.element_reference_quadrature = None
        self.edge_reference_quadrature = None

        # Krylov replacement of spsolve (main.py:231)
        # 'preconditioner': 'auto' = multigrid ('mg') when diffusion is present, otherwise the flow-ordered block Gauss-Seidel sweep
        # ('sweep'; 'bj' = plain block Jacobi)
        # 'mg_hierarchy': 'agglomerated' (P1 spaces on groups of 4 elements; defaults mg_nu 2, mg_omega 0.8, mg_alpha 2.2) or
        # 'aggregation' (piecewise constants; defaults 3, 0.7, 1.8) -- set 'mg_nu' / 'mg_omega' / 'mg_alpha' to override
        # 'mg_fp32': True = the smoothing sweeps read single-precision copies of the level matrices (measured: same solve time at
        # config 5, the sweeps are bound by the gathers of x, not by the matrix values -- off by default)
        # 'rtol' is below the rounding floor on purpose: like spsolve, the iteration runs until the TRUE residual stops improving;
        # at or below 'accept_rtol' that floor counts as convergence, above it the solve warns (solve_info['converged'] False)
        self.solver_options = {'method': 'auto', 'rtol': 1e-14, 'accept_rtol': 1e-7, 'max_iter': 200000, 'check_every': 10,
                               'preconditioner': 'auto', 'mg_hierarchy': 'agglomerated'}
        self.solve_info: typing.Optional[dict] = None
        self.timings: dict = {}
        self.device = None         # torch device; default: current CUDA device
        self.keep_device_inputs = False   # bench.py: keep the uploaded samples for kernel-only timing
        self.download_result = False      # True: dgfem() also brings B and L to (pinned) host memory, pipelined with the assembly
        # elements per evaluate/upload/assemble/download pipeline stage
        self.pipeline_chunk_elements = int(os.environ.get('REYNA_B200_PIPELINE_CHUNK', 65536))
        self._host = None                 # host copy of the last system (download_result)
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
        if self._host is not None:       # already on the host (pinned buffers that the next dgfem() reuses: copy out)
            indptr, indices, data, load = (np.array(self._host[k]) for k in ('indptr', 'indices', 'data', 'load'))
        else:
            indptr = dv['indptr'].cpu().numpy()
            indices = dv['indices'].cpu().numpy()
            data = dv['data'].cpu().numpy()
            load = dv['load'].cpu().numpy()
        B = csr_matrix((data, indices, indptr), shape=(N, dv.get('n_cols', N)))
        if self._perm is not None:
            # assembled in the locality renumbering (solve=True on a geometry without locality): back to the caller's numbering
            d = self.dim_elem
            dof = (np.asarray(self._geo.new_of_old, dtype=np.int64)[:, None] * d + np.arange(d)).reshape(-1)
            B = B[dof][:, dof].tocsr()
            B.sort_indices()
            load = load[dof]
        self._B = B
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
        # The multigrid hierarchy aggregates consecutive elements and the Krylov kernels gather neighbours: both need an
        # element numbering with spatial locality.  The reference numbers elements in random seed order
        # (polymesher/two_dimensional/main.py:111-165): such a geometry is solved on its Morton renumbering and the solution
        # is handed back in the caller's numbering (B and L too, lazily).  Assembly alone (solve=False) keeps the numbering.
        self._geo, self._perm = self.geometry, None
        if solve and self.renumber and not hasattr(self.geometry, 'elem_gid') and \
                getattr(flat_geometry(self.geometry), 'elem_gid', None) is None:
            from .partition import locality_order, RenumberedGeometry
            perm = locality_order(self.geometry)
            if perm is not None:
                work = getattr(self.geometry, '_rb_renumbered', None)
                if work is None:
                    work = RenumberedGeometry(self.geometry, perm)
                    try:
                        self.geometry._rb_renumbered = work
                    except Exception:   # pragma: no cover
                        pass
                self._geo, self._perm = work, work.old_of_new
        self._assemble_device()
        if verbose == 1:
            print(f"Assembly: {time.time() - _time}")

        if solve:
            sol = self._solve_device()
            if self._perm is not None:      # rows of the work numbering -> the caller's element numbering
                d = self.dim_elem
                out = np.empty_like(sol)
                out.reshape(-1, d)[self._perm] = sol.reshape(-1, d)
                sol = out
            self.solution = sol

    # -- host: coefficient sampling -----------------------------------------------------------------------------------
    def _staging(self, torch, key, shape, dtype=None):
        """Pinned host buffer (cached per DGFEM object) so that uploads / downloads are true asynchronous DMAs."""
        dtype = dtype or torch.float64
        buf = self._pinned.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._pinned[key] = buf
        return buf

    def _put(self, torch, pending, key, fn, X, tail) -> None:
        """Queue the evaluation of ``fn`` on all of ``X`` into pinned memory (host thread pool); ``_upload_pending`` waits for it
        and starts the upload."""
        M = X.shape[0]
        if M == 0:
            return
        # one spare row: the volume kernel's bulk copies move 16-byte units (rb_coefs, include/reyna_b200.h)
        buf = self._staging(torch, key, (M + 1,) + tail)
        job = sampling.submit_terms([(fn, tail, buf.numpy()[:M])], X)
        pending.append((key, buf, M, tail, job))

    @staticmethod
    def _upload_pending(dev, dev_t, pending) -> int:
        """Wait for the queued evaluations in order and start their uploads; returns the bytes uploaded."""
        nbytes = 0
        for key, buf, M, tail, job in pending:
            job.wait()
            buf.numpy()[M] = 0.0
            dev_t[key] = buf.to(dev, non_blocking=True)[:M]
            nbytes += M * int(np.prod(tail, dtype=np.int64)) * 8
        return nbytes

    def _submit_face_coefficients(self, torch, fg, pts):
        """Queue the evaluation of the interior-face and boundary-face samples (q-major) on the host thread pool."""
        pending = []
        b_mid = b_job = None
        if fg.n_iface:
            mid, Xf = pts['mid_i'], pts['Xi']
            if self.advection is not None:
                # first in the queue: the upwind flags (structure) are the first thing the device needs
                b_mid = np.empty((mid.shape[0], 2), dtype=np.float64)
                b_job = sampling.submit_terms([(self.advection, (2,), b_mid)], mid)
            if self.diffusion is not None:
                self._put(torch, pending, 'if_a', self.diffusion, Xf, (2, 2))
                self._put(torch, pending, 'if_a_mid', self.diffusion, mid, (2, 2))
            if self.advection is not None:
                self._put(torch, pending, 'if_b', self.advection, Xf, (2,))
        if fg.n_bface:
            mid, Xb = pts['mid_b'], pts['Xb']
            self._put(torch, pending, 'bf_g', self.dirichlet_bcs, Xb, ())
            if self.diffusion is not None:
                self._put(torch, pending, 'bf_a', self.diffusion, Xb, (2, 2))
                self._put(torch, pending, 'bf_a_mid', self.diffusion, mid, (2, 2))
            if self.advection is not None:
                self._put(torch, pending, 'bf_b', self.advection, Xb, (2,))
        return pending, b_mid, b_job

    def _sample_face_coefficients(self, torch, dev, fg, pts, submitted=None):
        """Interior-face and boundary-face samples (q-major) + upwind flags: everything the structure and the prepass need."""
        pending, b_mid, b_job = submitted if submitted is not None else self._submit_face_coefficients(torch, fg, pts)
        dev_t, nbytes = {}, 0
        self._upwind_host = None
        if fg.n_iface:
            if self.advection is not None:
                b_job.wait()
                # b(mid) . n >= 1e-12 (full_assembly.py:194); written out per component: np.sum(.., axis=1) over two columns
                # costs 25 ns per face on one core (75 ms at 3 M faces), the same two products and one add cost 3 ns
                upw = ((b_mid[:, 0] * fg.inorm[:, 0] + b_mid[:, 1] * fg.inorm[:, 1]) >= 1e-12).astype(np.uint8)
                dev_t['if_upwind'] = torch.from_numpy(upw).to(dev)
                nbytes += upw.nbytes
                self._upwind_host = upw
        elif self.advection is not None:
            dev_t['if_upwind'] = torch.zeros(1, dtype=torch.uint8, device=dev)     # no interior face: nothing to flag
            self._upwind_host = np.zeros(0, dtype=np.uint8)
        nbytes += self._upload_pending(dev, dev_t, pending)
        return dev_t, nbytes

    def _volume_terms(self):
        """(key, callable, tail shape) of the volume sample arrays that are present."""
        terms = []
        if self.diffusion is not None:
            terms.append(('vol_a', self.diffusion, (2, 2)))
        if self.advection is not None:
            terms.append(('vol_b', self.advection, (2,)))
        if self.reaction is not None:
            terms.append(('vol_c', self.reaction, ()))
        if self.forcing is not None:
            terms.append(('vol_f', self.forcing, ()))
        return terms

    def _chunks(self, fg, vs):
        """Element chunks [(elem_lo, elem_hi, vround_lo, vround_hi)] of WHOLE volume rounds: the unit of the evaluate -> upload ->
        assemble -> download pipeline (``self.pipeline_chunk_elements`` elements each; one chunk for small meshes)."""
        R = int(getattr(fg, 'n_rows', fg.n_elem))
        target = max(1, int(self.pipeline_chunk_elements))
        K = max(1, min(64, (R + target - 1) // target))
        if K == 1 or vs.n_rounds == 0:
            return [(0, R, 0, vs.n_rounds)]
        round_elem0 = fg.tri_elem[vs.vround_tri0[:-1]].astype(np.int64)      # first element of every round (ascending)
        cuts = [0]
        for k in range(1, K):
            r = int(np.searchsorted(round_elem0, (k * R) // K, side='left'))
            # a round that starts inside an element with more than 32 triangles belongs to that element: move to its first round
            while r < vs.n_rounds and r > 0 and round_elem0[r] == round_elem0[r - 1]:
                r += 1
            if cuts[-1] < r < vs.n_rounds:
                cuts.append(r)
        cuts.append(vs.n_rounds)
        out = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            e_lo = int(round_elem0[a])
            e_hi = int(round_elem0[b]) if b < vs.n_rounds else R
            out.append((e_lo, e_hi, a, b))
        out[0] = (0,) + out[0][1:]
        return out

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
        cache = getattr(fg, '_rb_dev', None)      # cached on the flat geometry (itself cached on its geometry object)
        if cache is None:
            cache = {}
            try:
                fg._rb_dev = cache
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
            # geometry-only tables of the assembly (basis scalings, flattened faces, |k_b|_K): once per geometry and device
            L = _lib.lib()
            E = fg.n_elem
            g = _lib.RbGeom(n_nodes=fg.n_nodes, n_elem=E, n_rows=getattr(fg, 'n_rows', E), n_tri=fg.n_tri, n_iface=fg.n_iface,
                            n_bface=fg.n_bface, n_vrounds=vs.n_rounds, n_vsegs=vs.n_segs,
                            **{k: _ptr(v) for k, v in cache[key].items()})
            nbytes = L.reyna_b200_geometry_tables_bytes(E, fg.n_iface)
            tables = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(L.reyna_b200_geometry_tables(C.byref(g), _ptr(tables), nbytes,
                                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                           'reyna_b200_geometry_tables')
            cache[key]['tables'] = tables
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
        """Host side of one assembly: face samples -> structure / pattern / prepass, then the volume samples chunk by chunk:
        evaluate (host threads) | upload | assemble | download run as a pipeline on three streams, so the step costs about what
        the evaluation of the user's callables costs (the span of main.py:208-228)."""
        torch = _lib.require_cuda()
        L = _lib.lib()
        dev = torch.device(self.device) if self.device is not None else torch.device('cuda', torch.cuda.current_device())
        p = self.polydegree
        if not 1 <= p <= 3:
            raise _lib.ReynaB200Error(f'polynomial degree {p} is outside the supported range 1..3')
        tm = self.timings = {}
        t0 = time.perf_counter()
        fg = flat_geometry(self._geo)
        self._host = None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream()
            sp = C.c_void_p(stream.cuda_stream)
            gd = self._device_geometry(torch, dev, fg)
            vs = sampling.volume_stream_of(fg)
            tm['geometry_upload_s'] = time.perf_counter() - t0

            t1 = time.perf_counter()
            pts = sampling.cached_points(self._geo, fg, self.element_reference_quadrature, self.edge_reference_quadrature)
            # ALL evaluations of the call are queued on the host thread pool now -- face samples first, then the volume samples
            # chunk by chunk -- and consumed in that order below: the host threads never idle at a barrier between the stages
            face_jobs = self._submit_face_coefficients(torch, fg, pts)
            Xs = pts['Xs']
            M = Xs.shape[0]
            terms = self._volume_terms()
            Q = int(np.asarray(self.element_reference_quadrature[1]).shape[0])
            chunks = self._chunks(fg, vs)
            bufs = [self._staging(torch, key, (M + 1,) + tail) for key, _, tail in terms]
            vol_jobs = []
            for (_, _, r_lo, r_hi) in chunks:
                lo, hi = Q * int(vs.vround_tri0[r_lo]), Q * int(vs.vround_tri0[r_hi])
                vol_jobs.append((lo, hi, sampling.submit_terms([(fn, tail, buf.numpy()) for (_, fn, tail), buf in zip(terms, bufs)],
                                                               Xs, lo, hi)))
            try:
                cd, nbytes = self._sample_face_coefficients(torch, dev, fg, pts, submitted=face_jobs)
                flags, bnd_elem, bnd_off, bnd_face = self._boundary_lists(fg)
                bd = {k: torch.from_numpy(v).to(dev) for k, v in
                      dict(flags=flags, bnd_elem=bnd_elem, bnd_off=bnd_off, bnd_face=bnd_face).items()}
                bd['has_spill'] = bool((flags & 4).any())
                bd['bnd_elem_host'] = bnd_elem
                tm['face_eval_upload_s'] = time.perf_counter() - t1

                # volume samples: device arrays now (the kernels get their final addresses), contents chunk by chunk below
                for key, _, tail in terms:
                    cd[key] = torch.empty((M + 1,) + tail, dtype=torch.float64, device=dev)[:M]
                    nbytes += M * int(np.prod(tail, dtype=np.int64)) * 8
                tm['h2d_bytes'] = int(nbytes)

                ctx = self._begin_on_device(torch, L, dev, sp, fg, gd, cd, tm)
                self._start_index_download(torch, ctx)
                s_in = self._stream(torch, dev, 'h2d')
                t2 = time.perf_counter()
                for (e_lo, e_hi, r_lo, r_hi), (lo, hi, job) in zip(chunks, vol_jobs):
                    job.wait()
                    with torch.cuda.stream(s_in):
                        for (key, _, _), buf in zip(terms, bufs):
                            cd[key][lo:hi].copy_(buf[lo:hi], non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(s_in)
                    stream.wait_event(ready)
                    self._assemble_chunk(torch, L, sp, ctx, bd, e_lo, e_hi, r_lo, r_hi)
                    self._start_rows_download(torch, ctx, e_lo, e_hi)
                tm['volume_eval_pipeline_s'] = time.perf_counter() - t2
            except BaseException:
                # a callable or a launch failed: no queued task may still be writing into the (cached) staging buffers
                for job in [face_jobs[2]] + [j for _, _, _, _, j in face_jobs[0]] + [j for _, _, j in vol_jobs]:
                    if job is not None:
                        job.abandon()
                raise
            t3 = time.perf_counter()
            self._dev = self._finish_on_device(torch, L, dev, sp, fg, gd, ctx, bd, tm)
            self._inputs = (fg, gd, cd, bd) if self.keep_device_inputs else None
            stream.synchronize()
            tm['device_tail_s'] = time.perf_counter() - t3          # last chunk's upload + kernels after its evaluation
            t4 = time.perf_counter()
            self._finish_download(torch, ctx, tm)
            tm['download_tail_s'] = time.perf_counter() - t4        # what is left of the D2H when the device is done
        tm['assemble_total_s'] = time.perf_counter() - t0

    def _stream(self, torch, dev, name):
        key = (str(dev), name)
        st = self._side_streams.get(key)
        if st is None:
            st = self._side_streams[key] = torch.cuda.Stream(device=dev)
        return st

    # -- device ---------------------------------------------------------------------------------------------------------
    def _begin_on_device(self, torch, L, dev, sp, fg, gd, cd, tm):
        """Structure, CSR arrays, index pattern (side stream) and the assembly prepass: everything that needs the FACE samples
        only.  Returns the context the chunk launches and ``_finish_on_device`` work on."""
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
        # with diffusion the structure depends on the geometry only: after the first assembly on a geometry its four counters
        # are known on the host and the call does not have to synchronise to return them
        known = getattr(self._geo, '_rb_counts', None) if has_diff else None
        key = (R, int(fg.n_iface))
        main_stream = torch.cuda.current_stream()
        side = self._stream(torch, dev, 'pattern')
        structure_ready = None
        if known is not None and known[0] == key:
            # the counters are known: adjacency + index pattern are built on the side stream while the main stream already runs
            # the penalties prepass and the volume kernel (neither needs the structure); the face kernel waits for the event
            st.n_items, st.n_groups, st.n_dup_items, st.max_items = known[1]
            side.wait_stream(main_stream)
            _lib.check(L.reyna_b200_structure_known(C.byref(g), 1, None, C.byref(st), _ptr(ws), ws_bytes,
                                                    C.c_void_p(side.cuda_stream)), 'reyna_b200_structure_known')
            structure_ready = torch.cuda.Event()
            structure_ready.record(side)
            for t in (adj_off, adj_item, adj_nb, grp_off, counts, ws):
                t.record_stream(side)
        else:
            _lib.check(L.reyna_b200_structure(C.byref(g), int(has_diff), _ptr(cd.get('if_upwind')), C.byref(st), _ptr(ws),
                                              ws_bytes, sp), 'reyna_b200_structure')
            if has_diff:
                try:
                    self._geo._rb_counts = (key, (st.n_items, st.n_groups, st.n_dup_items, st.max_items))
                except Exception:   # pragma: no cover - exotic geometry objects
                    pass
            side.wait_stream(main_stream)
        nnz = d * d * (R + st.n_groups)
        if nnz >= 2 ** 31:
            raise _lib.ReynaB200Error(f'nnz = {nnz} does not fit the int32 CSR indices')
        # -- CSR arrays
        indptr = torch.empty(N + 1, **i32)
        indices = torch.empty(nnz, **i32)
        data = torch.empty(nnz, **f64)
        load = torch.empty(N, **f64)
        zero_count = torch.zeros(1, dtype=torch.int64, device=dev)
        # the index arrays only depend on the structure: they are written on the side stream while the assembly runs, and joined
        # before the step ends
        if structure_ready is not None:
            side.wait_stream(main_stream)            # (allocations above)
        _lib.check(L.reyna_b200_csr_pattern(C.byref(g), C.byref(st), p, _ptr(indptr), _ptr(indices),
                                            C.c_void_p(side.cuda_stream)), 'reyna_b200_csr_pattern')
        indptr.record_stream(side)
        indices.record_stream(side)
        aws_bytes = L.reyna_b200_assemble_workspace(E, fg.n_iface, vs.n_segs, p)
        aws = torch.empty(aws_bytes, dtype=torch.uint8, device=dev)
        return dict(g=g, st=st, co=co, q=q, R=R, N=N, d=d, nnz=int(nnz), indptr=indptr, indices=indices, data=data, load=load,
                    zero_count=zero_count, aws=aws, aws_bytes=aws_bytes, side=side, main=main_stream, dev=dev, prepared=False,
                    structure_ready=structure_ready,
                    structure=dict(adj_off=adj_off, adj_item=adj_item, adj_nb=adj_nb, grp_off=grp_off, n_items=st.n_items,
                                   n_groups=st.n_groups, n_dup_items=st.n_dup_items), spill=None, _keep=(ws, counts))

    def _assemble_chunk(self, torch, L, sp, ctx, bd, e_lo, e_hi, r_lo, r_hi) -> None:
        """Block rows of elements [e_lo, e_hi): prepass (first chunk), volume + face kernels, boundary terms of the chunk."""
        g, st, co, q = ctx['g'], ctx['st'], ctx['co'], ctx['q']
        if not ctx['prepared']:
            _lib.check(L.reyna_b200_assemble_prepare(C.byref(g), C.byref(st), C.byref(co), C.byref(q), float(self.sigma_D),
                                                     _ptr(ctx['aws']), ctx['aws_bytes'], sp), 'reyna_b200_assemble_prepare')
            ctx['prepared'] = True
        args = (C.byref(g), C.byref(st), C.byref(co), C.byref(q), float(self.sigma_D), _ptr(ctx['data']), _ptr(ctx['load']),
                _ptr(ctx['zero_count']), _ptr(ctx['aws']), ctx['aws_bytes'], int(e_lo), int(e_hi), int(r_lo), int(r_hi), sp)
        if ctx['structure_ready'] is not None:
            # the structure is still being built on the side stream: volume kernel first, the face kernel after the join
            _lib.check(L.reyna_b200_assemble_volume_range(*args), 'reyna_b200_assemble_volume_range')
            ctx['main'].wait_event(ctx['structure_ready'])
            ctx['structure_ready'] = None
            _lib.check(L.reyna_b200_assemble_faces_range(*args), 'reyna_b200_assemble_faces_range')
        else:
            _lib.check(L.reyna_b200_assemble_range(*args), 'reyna_b200_assemble_range')
        be = bd['bnd_elem_host']
        n_all = int(be.shape[0])
        if n_all:
            if ctx['spill'] is None:
                ctx['spill'] = torch.zeros(n_all, dtype=torch.float64, device=ctx['dev'])
            i0, i1 = int(np.searchsorted(be, e_lo, side='left')), int(np.searchsorted(be, e_hi, side='left'))
            if i1 > i0:
                _lib.check(L.reyna_b200_boundary(C.byref(g), C.byref(st), C.byref(co), C.byref(q), float(self.sigma_D),
                                                 i1 - i0, bd['bnd_elem'].data_ptr() + 4 * i0, bd['bnd_off'].data_ptr() + 4 * i0,
                                                 _ptr(bd['bnd_face']), _ptr(bd['flags']), _ptr(ctx['data']), _ptr(ctx['load']),
                                                 ctx['spill'].data_ptr() + 8 * i0, _ptr(ctx['zero_count']), sp),
                           'reyna_b200_boundary')

    def _finish_on_device(self, torch, L, dev, sp, fg, gd, ctx, bd, tm, resolve=True):
        """The reference's B[0,0] quirk, join of the pattern stream, exact-zero pruning; returns the device-resident system.
        ``resolve=False`` leaves the count of exact zeros on the device (no host synchronisation; ``_resolve_zeros`` later)."""
        p, d, R, N, nnz = self.polydegree, ctx['d'], ctx['R'], ctx['N'], ctx['nnz']
        indptr, indices, data, load, zero_count = ctx['indptr'], ctx['indices'], ctx['data'], ctx['load'], ctx['zero_count']
        i32 = dict(dtype=torch.int32, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)
        if bd.get('has_spill', False):
            if getattr(fg, 'elem_gid', None) is not None:
                raise _lib.ReynaB200Error('partitioned assembly does not mirror the reference index quirk of main.py:397-400 '
                                          '(boundary edges that are only partly elliptic)')
            data[0] -= ctx['spill'].sum()          # the reference's B[0,0] quirk (main.py:397-400)
            zero_count += (data[0] == 0).to(torch.int64)
            ctx['spill_applied'] = True
        ctx['main'].wait_stream(ctx['side'])
        blocked = True
        structural = (indptr, indices, data)        # the solver (and its multigrid hierarchy) works on the block pattern
        tm['structural_nnz'] = int(nnz)
        if not resolve:
            return dict(indptr=indptr, indices=indices, data=data, load=load, blocked=True, n=N, d=d, structural=structural,
                        bbox=gd['bbox'], n_cols=int(getattr(self._geo, 'n_global_elements', R)) * d,
                        structure=ctx['structure'], zero_count=zero_count, unresolved=True,
                        _assemble_call=(ctx['g'], ctx['st'], ctx['co'], ctx['q'], data, load, zero_count, ctx['aws'],
                                        ctx['aws_bytes']))
        nz = int(zero_count.item())         # synchronises
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
            ctx['compacted'] = True
        tm['nnz'] = int(nnz)
        return dict(indptr=indptr, indices=indices, data=data, load=load, blocked=blocked, n=N, d=d, structural=structural,
                    bbox=gd['bbox'],
                    n_cols=int(getattr(self._geo, 'n_global_elements', R)) * d,
                    structure=ctx['structure'],
                    # ctypes argument blocks of the assembly launch (bench.py re-launches it alone for the roofline)
                    _assemble_call=(ctx['g'], ctx['st'], ctx['co'], ctx['q'], structural[2], load, zero_count, ctx['aws'],
                                    ctx['aws_bytes']))

    def _assemble_on_device(self, torch, L, dev, sp, fg, gd, cd, bd, tm, resolve=True):
        """Everything from the uploaded samples to the device-resident CSR system in one go (inputs already resident in HBM:
        bench.py's kernel-only timing, tools).  With ``resolve=False`` the step ends without reading the exact-zero counter
        back (every kernel has run; the caller checks ``int(dv['zero_count'])`` when it consumes the result)."""
        ctx = self._begin_on_device(torch, L, dev, sp, fg, gd, cd, tm)
        vs = sampling.volume_stream_of(fg)
        self._assemble_chunk(torch, L, sp, ctx, bd, 0, ctx['R'], 0, vs.n_rounds)
        return self._finish_on_device(torch, L, dev, sp, fg, gd, ctx, bd, tm, resolve=resolve)

    # -- pipelined download of the result (``download_result``) ------------------------------------------------------------------
    def _start_index_download(self, torch, ctx) -> None:
        if not self.download_result:
            return
        s_out = self._stream(torch, ctx['dev'], 'd2h')
        host = ctx['host'] = {}
        for k in ('indptr', 'indices', 'data', 'load'):
            host[k] = self._staging(torch, 'out_' + k, tuple(ctx[k].shape), ctx[k].dtype)
        ctx['grp_host'] = ctx['structure']['grp_off'].cpu().numpy().astype(np.int64)   # row offsets of the chunks (4 bytes per element)
        s_out.wait_stream(ctx['side'])               # the pattern kernel has written the index arrays
        with torch.cuda.stream(s_out):
            host['indptr'].copy_(ctx['indptr'], non_blocking=True)
            host['indices'].copy_(ctx['indices'], non_blocking=True)

    def _start_rows_download(self, torch, ctx, e_lo, e_hi) -> None:
        """Values and load of the block rows of elements [e_lo, e_hi): contiguous ranges of the CSR value array."""
        if not self.download_result:
            return
        s_out = self._stream(torch, ctx['dev'], 'd2h')
        d = ctx['d']
        gh = ctx['grp_host']
        lo, hi = d * d * (int(gh[e_lo]) + e_lo), d * d * (int(gh[e_hi]) + e_hi)
        s_out.wait_stream(ctx['main'])
        with torch.cuda.stream(s_out):
            ctx['host']['data'][lo:hi].copy_(ctx['data'][lo:hi], non_blocking=True)
            ctx['host']['load'][e_lo * d:e_hi * d].copy_(ctx['load'][e_lo * d:e_hi * d], non_blocking=True)

    def _finish_download(self, torch, ctx, tm) -> None:
        if not self.download_result:
            return
        s_out = self._stream(torch, ctx['dev'], 'd2h')
        dv = self._dev
        host = ctx['host']
        if ctx.get('compacted'):
            # rare: exact zeros were pruned, the arrays downloaded so far are the structural ones -- fetch the final ones
            host = {k: dv[k].cpu() for k in ('indptr', 'indices', 'data', 'load')}
        else:
            s_out.synchronize()
            if ctx.get('spill_applied'):
                host['data'][0:1].copy_(dv['data'][0:1])
        self._host = {k: v.numpy() for k, v in host.items()}
        tm['d2h_bytes'] = int(sum(v.nbytes for v in self._host.values()))

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
                                rtol=float(so.get('rtol', 1e-14)), check_every=int(so.get('check_every', 10)),
                                accept_rtol=float(so.get('accept_rtol', 1e-7)))
        prec = so.get('preconditioner', 'auto')
        if prec == 'auto':
            # diffusion: multigrid; advection-reaction only: the flow-ordered block Gauss-Seidel sweep (a direct solve when
            # the upwind graph has no cycle)
            prec = 'mg' if self.diffusion is not None else 'sweep'
        if prec == 'sweep' and (self.diffusion is not None or method != 'bicgstab' or getattr(self, '_upwind_host', None) is None):
            prec = 'bj'
        info = _lib.RbSolveInfo()
        t0 = time.perf_counter()
        s_indptr, s_indices, s_data = dv['structural']
        mg = C.c_void_p(None)
        mg_levels = None
        sweep = None
        if prec == 'sweep':
            from .flow import flow_levels
            fgw = flat_geometry(self._geo)
            order, level_off, n_cyclic = flow_levels(fgw.ielem, self._upwind_host, N // d)
            if n_cyclic and level_off.shape[0] - 1 > 8 * int(np.sqrt(N // d)) + 64:
                prec = 'bj'        # closed streamlines and a nearly sequential order: the sweep would be all launch latency
        if prec == 'sweep':
            sweep = dict(order=torch.from_numpy(order).to(dev), off=np.ascontiguousarray(level_off, dtype=np.int32),
                         levels=int(level_off.shape[0] - 1), cyclic=n_cyclic)
            opts.sweep_elems = _ptr(sweep['order'])
            opts.sweep_level_off = sweep['off'].ctypes.data
            opts.sweep_levels = sweep['levels']
        with torch.cuda.device(dev):
            # a side stream: the library replays the multigrid application as a CUDA graph, which the legacy default
            # stream cannot capture
            caller = torch.cuda.current_stream()
            side = self._stream(torch, dev, 'solve')
            side.wait_stream(caller)
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
                                       elem_bbox=_ptr(dv['bbox']) if agglo else None,
                                       fp32_levels=int(bool(so.get('mg_fp32', os.environ.get('REYNA_B200_MG_FP32', '0') != '0'))))
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
                torch.cuda.set_stream(caller)      # the caller's stream context is left as it was found
                if mg.value:
                    L.reyna_b200_mg_destroy(mg)
        self.solve_info = dict(method=method, iterations=int(info.iterations), rel_residual=float(info.rel_residual),
                               converged=bool(info.converged), restarts=int(info.restarts),
                               preconditioner=prec, mg_levels=mg_levels, setup_seconds=setup_s,
                               stagnated=bool(info.stagnated), singular_blocks=int(info.singular_blocks),
                               sweep_levels=sweep['levels'] if sweep else None,
                               flow_cycles_broken=sweep['cyclic'] if sweep else None,
                               seconds=time.perf_counter() - t0)
        self.timings['solve_s'] = self.solve_info['seconds']
        if not info.converged:
            warnings.warn(f"Krylov solver stopped after {info.iterations} iterations with relative residual "
                          f"{info.rel_residual:.3e} (accepted: {float(so.get('accept_rtol', 1e-7)):.1e}); the solution may be "
                          f"inaccurate")
        if info.singular_blocks:
            warnings.warn(f"{info.singular_blocks} singular diagonal blocks were replaced by identity in the preconditioner")
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
        from .tools import plot_DG
        plot_DG(self.solution, self.geometry, self.polydegree)
