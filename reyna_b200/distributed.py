"""Multi-GPU ``dgfem(solve=True)``: one process per GPU, elements partitioned into contiguous (Morton-ordered) ranges.

The reference is a single-process program; this module is new.  Assembly needs no collective
(``partition.py``).  The Krylov solve runs the same ``reyna_b200_solve`` on every rank with

* the rank's block rows in LOCAL column numbering (owned coefficients followed by ghost coefficients),
* a halo exchange of ghost-element coefficients before every SpMV and an all-reduce of every dot product, both issued
  from inside the library over NCCL on the solver's stream (``csrc/dist.cu``),
* the multigrid hierarchy distributed as well: fine and P1 levels are partitioned (two 3-coefficient halo exchanges per
  application), the piecewise-constant chain is GLOBAL and replicated (its rows are all-gathered once per solve, its
  right-hand side once per application), so the iteration counts do not grow with the number of GPUs.

``torch.distributed`` hands out the NCCL unique id, gathers the solution and -- for the multigrid set-up -- all-gathers CUDA
tensors: the process group must use the NCCL backend (with gloo only ``preconditioner='bj'`` / ``'sweep'`` works).
"""
from __future__ import annotations

import ctypes as C
import os
import time
import typing

import numpy as np

from . import _lib
from .dgfem import DGFEM, _ptr
from .geometry import flat_geometry
from .partition import HaloPlan, PartitionedGeometry, partition_bounds, partition_geometry


def nccl_library_path() -> typing.Optional[str]:
    """The libnccl.so.2 that PyTorch loaded (so that exactly one NCCL lives in the process)."""
    try:
        import nvidia.nccl
        for base in nvidia.nccl.__path__:
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return None


_COMM = {}


def _communicator(torch, dev, rank: int, world: int):
    """The process's NCCL communicator for this world (created once: ncclCommInitRank takes 0.3 ... several seconds)."""
    key = (world, rank, str(dev))
    if key in _COMM:
        return _COMM[key]
    import torch.distributed as dist
    L = _lib.lib()
    path = nccl_library_path()
    cpath = path.encode() if path else None
    idbuf = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(L.reyna_b200_nccl_unique_id(cpath, idbuf), "reyna_b200_nccl_unique_id")
    t = torch.tensor(list(idbuf.raw), dtype=torch.uint8, device=dev if dist.get_backend() == "nccl" else "cpu")
    dist.broadcast(t, src=0)
    comm = C.c_void_p(None)
    _lib.check(L.reyna_b200_comm_create(cpath, bytes(t.cpu().tolist()), rank, world, C.byref(comm)), "reyna_b200_comm_create")
    _COMM[key] = comm
    return comm


class DistributedDGFEM:
    """``DGFEM`` on one rank's partition plus the distributed solve.

    >>> ddg = DistributedDGFEM(geometry, polynomial_degree=2)        # inside a torch.distributed process group
    >>> ddg.add_data(...); ddg.dgfem(solve=True)
    >>> ddg.solution            # this rank's coefficients;  ddg.gather_solution() -> the global vector on every rank
    """

    def __init__(self, geometry, polynomial_degree: int = 1, rank: int = None, world: int = None, device=None):
        import torch.distributed as dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        # contiguous index ranges are compact patches only when the numbering has locality: a geometry in the reference's random
        # seed order is partitioned (and solved) in its Morton renumbering; gather_solution() hands the caller's numbering back
        from .partition import locality_order, RenumberedGeometry
        self._perm = locality_order(geometry)
        if self._perm is not None:
            geometry = RenumberedGeometry(geometry, self._perm)
        self.global_geometry = geometry
        self.bounds = partition_bounds(int(geometry.n_elements), self.world)
        if np.any(np.diff(self.bounds) <= 0):
            # a rank without elements would skip the collectives of the Krylov solve and hang the others
            raise ValueError(f'{int(geometry.n_elements)} elements cannot be split over {self.world} ranks '
                             f'(ranges start at multiples of 4): use fewer ranks')
        self.part = partition_geometry(geometry, int(self.bounds[self.rank]), int(self.bounds[self.rank + 1]))
        self.plan = HaloPlan(geometry, self.bounds, self.rank)
        self.dg = DGFEM(self.part, polynomial_degree)
        if device is not None:
            self.dg.device = device
        self.solution: typing.Optional[np.ndarray] = None
        self.solve_info: typing.Optional[dict] = None
        self._dist = None

    # ---- pass-through of the reference API ------------------------------------------------------------------------------
    def add_data(self, **kwargs):
        self.dg.add_data(**kwargs)

    @property
    def solver_options(self):
        return self.dg.solver_options

    @property
    def dim_elem(self):
        return self.dg.dim_elem

    def dgfem(self, solve: bool = True, verbose: int = 0):
        self.dg.dgfem(solve=False, verbose=verbose)
        if solve:
            self.solution = self.solve()

    # ---- NCCL context ---------------------------------------------------------------------------------------------------
    def _context(self, torch, dev, d: int):
        if self._dist is not None:
            return self._dist
        L = _lib.lib()
        comm = _communicator(torch, dev, self.rank, self.world)
        peers = self.plan.peers()
        send_idx, send_off, recv_off = [], [0], [0]
        for q in peers:
            s = self.plan.send.get(q, np.zeros(0, dtype=np.int64))
            r = self.plan.recv.get(q, np.zeros(0, dtype=np.int64))
            send_idx.append(np.asarray(s, dtype=np.int32))
            send_off.append(send_off[-1] + len(s))
            if len(r):
                # ghosts of one peer are a contiguous block of the (gid-sorted) ghost layer
                assert int(r[0]) - self.plan.n_rows == recv_off[-1] and np.all(np.diff(r) == 1)
            recv_off.append(recv_off[-1] + len(r))
        sidx = np.concatenate(send_idx) if send_idx else np.zeros(0, dtype=np.int32)
        sidx_dev = torch.from_numpy(np.ascontiguousarray(sidx, dtype=np.int32)).to(dev) if sidx.size else \
            torch.zeros(1, dtype=torch.int32, device=dev)
        n = len(peers)
        arr = lambda v: (C.c_int * max(1, len(v)))(*v)       # noqa: E731
        handle = C.c_void_p(None)
        _lib.check(L.reyna_b200_dist_create(comm, self.rank, self.world, d, self.plan.n_rows, n, arr(peers),
                                            arr(send_off), arr(recv_off), sidx_dev.data_ptr(), C.byref(handle)),
                   "reyna_b200_dist_create")
        self._dist = dict(handle=handle, send_idx=sidx_dev)
        return self._dist

    def close(self):
        if self._dist is not None:
            _lib.lib().reyna_b200_dist_destroy(self._dist["handle"])
            self._dist = None

    # ---- solve ----------------------------------------------------------------------------------------------------------
    def solve(self) -> np.ndarray:
        torch = _lib.require_cuda()
        L = _lib.lib()
        dg = self.dg
        dv = dg._dev
        d, N = dv["d"], dv["n"]
        fg = flat_geometry(self.part)
        n_ghost = (fg.n_elem - fg.n_rows) * d
        dev = dv["data"].device
        so = dg.solver_options
        method = so.get("method", "auto")
        if method == "auto":
            method = "cg" if dg.advection is None else "bicgstab"
        prec = so.get("preconditioner", "auto")
        if prec == "auto":
            # advection-reaction without diffusion: the flow-ordered block Gauss-Seidel sweep of the rank's own rows (ghosts from
            # the last halo exchange) preconditions BiCGStab -- exact inside a partition, Jacobi across partitions
            prec = "mg" if dg.diffusion is not None else "sweep"
        if prec == "sweep" and (dg.diffusion is not None or method != "bicgstab" or getattr(dg, "_upwind_host", None) is None):
            prec = "bj"
        opts = _lib.RbSolveOpts(method=0 if method == "cg" else 1, max_iter=int(so.get("max_iter", 200000)),
                                rtol=float(so.get("rtol", 1e-14)), check_every=int(so.get("check_every", 10)),
                                accept_rtol=float(so.get("accept_rtol", 1e-7)))
        info = _lib.RbSolveInfo()
        t0 = time.perf_counter()
        sweep = None
        if prec == "sweep":
            from .flow import flow_levels
            order, level_off, n_cyclic = flow_levels(fg.ielem, dg._upwind_host, N // d)     # edges between owned elements only
            if n_cyclic and level_off.shape[0] - 1 > 8 * int(np.sqrt(N // d)) + 64:
                prec = "bj"
            else:
                sweep = dict(order=torch.from_numpy(order).to(dev), off=np.ascontiguousarray(level_off, dtype=np.int32))
                opts.sweep_elems = _ptr(sweep["order"])
                opts.sweep_level_off = sweep["off"].ctypes.data
                opts.sweep_levels = int(level_off.shape[0] - 1)
        with torch.cuda.device(dev):
            sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            ctx = self._context(torch, dev, d)
            t_ctx = time.perf_counter() - t0
            # the rank's rows with local column numbering
            g, st = dv["_assemble_call"][0], dv["_assemble_call"][1]
            s_indptr, _, s_data = dv["structural"]
            indices_local = torch.empty_like(dv["structural"][1])
            indptr_local = torch.empty_like(s_indptr)
            _lib.check(L.reyna_b200_csr_pattern_local(C.byref(g), C.byref(st), dg.polydegree, _ptr(indptr_local),
                                                      _ptr(indices_local), sp), "reyna_b200_csr_pattern_local")
            mg = C.c_void_p(None)
            keep = []
            try:
                if prec == "mg":
                    agglo = so.get("mg_hierarchy", "agglomerated") == "agglomerated"
                    mo = _lib.RbMgOpts(nu=int(so.get("mg_nu", 2 if agglo else 3)),
                                       omega=float(so.get("mg_omega", 0.8 if agglo else 0.7)),
                                       alpha=float(so.get("mg_alpha", 2.2 if agglo else 1.8)),
                                       fp32_levels=int(bool(so.get("mg_fp32", os.environ.get("REYNA_B200_MG_FP32", "0") != "0"))))
                    bounds32 = (C.c_int32 * (self.world + 1))(*[int(b) for b in self.bounds])
                    dd = _lib.RbMgDist(dist=ctx["handle"], rank=self.rank, world=self.world,
                                       n_ghost_elem=fg.n_elem - fg.n_rows, n_global_elem=int(self.bounds[-1]),
                                       bounds=C.cast(bounds32, C.c_void_p))
                    if agglo:
                        l1_indptr, l1_indices, l1_data, l1_bbox, l0_tr = self._global_level1(torch, dev, dv)
                        keep += [l1_indptr, l1_indices, l1_data, l1_bbox, l0_tr, bounds32]
                        dd.l1_indptr, dd.l1_indices, dd.l1_data = _ptr(l1_indptr), _ptr(l1_indices), _ptr(l1_data)
                        dd.l1_nnz, dd.l1_bbox, dd.l0_tr = int(l1_data.numel()), _ptr(l1_bbox), _ptr(l0_tr)
                    else:
                        p0_indptr, p0_indices, p0_data = self._global_p0(torch, dev, dv)
                        keep += [p0_indptr, p0_indices, p0_data, bounds32]
                        dd.p0_indptr, dd.p0_indices, dd.p0_data = _ptr(p0_indptr), _ptr(p0_indices), _ptr(p0_data)
                        dd.p0_nnz = int(p0_data.numel())
                    _lib.check(L.reyna_b200_mg_create(N, d, dg.polydegree, _ptr(indptr_local), _ptr(indices_local),
                                                      _ptr(s_data), C.byref(mo), C.byref(dd), C.byref(mg), sp),
                               "reyna_b200_mg_create")
                torch.cuda.synchronize()
                t_setup = time.perf_counter() - t0
                x = torch.zeros(N + n_ghost, dtype=torch.float64, device=dev)
                ws_bytes = L.reyna_b200_solve_workspace(N, n_ghost, d, int(s_data.numel()))
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                halo = C.cast(L.reyna_b200_nccl_halo, C.c_void_p)
                allr = C.cast(L.reyna_b200_nccl_allreduce, C.c_void_p)
                _lib.check(L.reyna_b200_solve(N, n_ghost, d, 1, _ptr(indptr_local), _ptr(indices_local), _ptr(s_data),
                                              _ptr(dv["load"]), _ptr(x), C.byref(opts), C.byref(info), halo, allr,
                                              ctx["handle"], mg, _ptr(ws), ws_bytes, sp), "reyna_b200_solve")
                sol = x[:N].cpu().numpy()
            finally:
                if mg.value:
                    L.reyna_b200_mg_destroy(mg)
        ex, ar = C.c_ulonglong(0), C.c_ulonglong(0)
        L.reyna_b200_dist_counters(ctx["handle"], C.byref(ex), C.byref(ar))
        self.solve_info = dict(method=method, preconditioner=prec, iterations=int(info.iterations),
                               rel_residual=float(info.rel_residual), converged=bool(info.converged),
                               restarts=int(info.restarts), stagnated=bool(info.stagnated), world=self.world,
                               halo_exchanges=int(ex.value), allreduces=int(ar.value), n_ghost_elements=fg.n_elem - fg.n_rows,
                               singular_blocks=int(info.singular_blocks), sweep_levels=int(opts.sweep_levels) or None,
                               context_seconds=t_ctx, setup_seconds=t_setup, seconds=time.perf_counter() - t0)
        if not info.converged:
            import warnings
            warnings.warn(f"Krylov solver stopped after {info.iterations} iterations with relative residual "
                          f"{info.rel_residual:.3e}; the solution may be inaccurate")
        return sol

    def _global_p0(self, torch, dev, dv):
        """The GLOBAL piecewise-constant matrix A0 = (mode-(0,0) entry of every d x d block of B), replicated on every
        rank: each rank extracts its rows (global element numbering) from its structural CSR and the rows are
        all-gathered once per solve (8 B + 4 B per block: 0.1 GB at 1M elements)."""
        import torch.distributed as dist
        d = dv["d"]
        s_indptr, s_indices, s_data = dv["structural"]          # global column numbering
        N = dv["n"]
        starts = s_indptr[0:N:d].to(torch.int64)                # first scalar row of every block row
        nb = ((s_indptr[1:N + 1:d].to(torch.int64) - starts) // d)
        rows = torch.repeat_interleave(torch.arange(nb.numel(), device=dev), nb)
        first = torch.cumsum(nb, 0) - nb
        pos = starts[rows] + (torch.arange(rows.numel(), device=dev) - first[rows]) * d
        cols = (s_indices[pos] // d).to(torch.int32)
        vals = s_data[pos].contiguous()
        # all-gather (uneven): sizes first, then padded payloads
        cnt = torch.tensor([cols.numel(), nb.numel()], dtype=torch.int64, device=dev)
        cnts = [torch.zeros_like(cnt) for _ in range(self.world)]
        dist.all_gather(cnts, cnt)
        sizes = [int(c[0]) for c in cnts]
        nrows = [int(c[1]) for c in cnts]
        mx, mr = max(sizes), max(nrows)

        def gather(t, n_each, width, dtype):
            pad = torch.zeros(width, dtype=dtype, device=dev)
            pad[:t.numel()] = t
            outs = [torch.empty(width, dtype=dtype, device=dev) for _ in range(self.world)]
            dist.all_gather(outs, pad)
            return torch.cat([o[:n] for o, n in zip(outs, n_each)])

        g_cols = gather(cols, sizes, mx, torch.int32)
        g_vals = gather(vals, sizes, mx, torch.float64)
        g_nb = gather(nb.to(torch.int32), nrows, mr, torch.int32)
        g_indptr = torch.zeros(g_nb.numel() + 1, dtype=torch.int32, device=dev)
        g_indptr[1:] = torch.cumsum(g_nb.to(torch.int64), 0).to(torch.int32)
        return g_indptr, g_cols.contiguous(), g_vals.contiguous()

    def _all_gather_cat(self, torch, dev, t, dtype):
        """Concatenation over ranks of 1-D tensors of different lengths (sizes first, then padded payloads)."""
        import torch.distributed as dist
        cnt = torch.tensor([t.numel()], dtype=torch.int64, device=dev)
        cnts = [torch.zeros_like(cnt) for _ in range(self.world)]
        dist.all_gather(cnts, cnt)
        sizes = [int(c[0]) for c in cnts]
        width = max(max(sizes), 1)
        pad = torch.zeros(width, dtype=dtype, device=dev)
        pad[:t.numel()] = t
        outs = [torch.empty(width, dtype=dtype, device=dev) for _ in range(self.world)]
        dist.all_gather(outs, pad)
        return torch.cat([o[:n] for o, n in zip(outs, sizes)])

    @staticmethod
    def _box_coefficients(torch, bb):
        hx, hy = 0.5 * (bb[:, 1] - bb[:, 0]), 0.5 * (bb[:, 3] - bb[:, 2])
        mx, my = 0.5 * (bb[:, 1] + bb[:, 0]), 0.5 * (bb[:, 3] + bb[:, 2])
        r = torch.rsqrt(hx * hy)
        c01 = 0.8660254037844386 * r                      # c0 c1 of the orthonormal Legendre basis
        return mx, my, 0.5 * r, c01 / hx, c01 / hy

    def _global_level1(self, torch, dev, dv):
        """First agglomerated level of the multigrid hierarchy (csrc/multigrid.cu), GLOBAL and replicated: every rank forms the
        Galerkin rows of its own aggregates, ``sum_{e in G, f in H} Pe^T A1[e,f] Pf`` with the 3 x 3 mode blocks A1 of its
        structural CSR (global column numbering) and the bounding-box transfer matrices P, and the rows are all-gathered
        once per solve.  Returns (indptr, indices, data) of the scalar CSR matrix with complete 3 x 3 blocks, the aggregate
        bounding boxes and the transfer coefficients of the owned elements."""
        d, N = dv["d"], dv["n"]
        p = self.dg.polydegree
        s_indptr, s_indices, s_data = dv["structural"]
        E_loc = N // d
        lo = int(self.bounds[self.rank])
        E_glob = int(self.bounds[-1])
        n_agg = (E_glob + 3) // 4
        # bounding boxes: all elements (the global geometry is replicated on the host) and all aggregates
        bb = torch.from_numpy(np.ascontiguousarray(flat_geometry(self.global_geometry).bbox, dtype=np.float64)).to(dev).reshape(-1, 4)
        pad = n_agg * 4 - E_glob
        bbp = torch.cat([bb, bb[-1:].expand(pad, 4)]) if pad else bb
        g4 = bbp.reshape(n_agg, 4, 4)
        bb_agg = torch.stack([g4[:, :, 0].min(1).values, g4[:, :, 1].max(1).values, g4[:, :, 2].min(1).values,
                              g4[:, :, 3].max(1).values], dim=1).contiguous()
        agg_of = torch.arange(E_glob, device=dev) // 4
        mxf, myf, k00f, kxf, kyf = self._box_coefficients(torch, bb)
        mxc, myc, k00c, kxc, kyc = [a[agg_of] for a in self._box_coefficients(torch, bb_agg)]
        tr = torch.zeros(E_glob, 8, dtype=torch.float64, device=dev)
        tr[:, 0] = k00c / k00f
        tr[:, 1] = kyc / kyf
        tr[:, 2] = kyc * (myf - myc) / k00f
        tr[:, 3] = kxc / kxf
        tr[:, 4] = kxc * (mxf - mxc) / k00f
        P = torch.zeros(E_glob, 3, 3, dtype=torch.float64, device=dev)     # rows: element modes, columns: aggregate modes
        P[:, 0, 0], P[:, 0, 1], P[:, 0, 2], P[:, 1, 1], P[:, 2, 2] = tr[:, 0], tr[:, 2], tr[:, 4], tr[:, 1], tr[:, 3]
        # 3 x 3 mode blocks of this rank's block rows
        sel = torch.tensor([0, 1, p + 1] if d > 3 else [0, 1, 2], device=dev)
        starts = s_indptr[0:N:d].to(torch.int64)
        rowlen = s_indptr[1:N + 1:d].to(torch.int64) - starts
        nb = rowlen // d
        brow = torch.repeat_interleave(torch.arange(E_loc, device=dev), nb)
        first = torch.cumsum(nb, 0) - nb
        slot = torch.arange(brow.numel(), device=dev) - first[brow]
        base = starts[brow] + slot * d                                         # entry (row mode 0, column mode 0) of the block
        f_glob = (s_indices[base] // d).to(torch.int64)
        pos = base[:, None, None] + sel[:, None] * rowlen[brow][:, None, None] + sel[None, :]
        A1 = s_data[pos]                                                       # [n_blocks, 3, 3]
        e_glob = brow + lo
        At = torch.bmm(torch.bmm(P[e_glob].transpose(1, 2), A1), P[f_glob])
        key = (e_glob // 4) * n_agg + f_glob // 4
        order = torch.argsort(key, stable=True)
        key_s, At_s = key[order], At[order]
        ukey, inv, counts = torch.unique_consecutive(key_s, return_inverse=True, return_counts=True)
        # deterministic segmented sum: the j-th member of every group is added in pass j (no two writes collide)
        gstart = torch.cumsum(counts, 0) - counts
        rank_in = torch.arange(key_s.numel(), device=dev) - gstart[inv]
        blocks = torch.zeros(ukey.numel(), 3, 3, dtype=torch.float64, device=dev)
        for j in range(int(counts.max().item()) if counts.numel() else 0):
            m = rank_in == j
            blocks[inv[m]] += At_s[m]
        g_key = self._all_gather_cat(torch, dev, ukey, torch.int64)
        g_blk = self._all_gather_cat(torch, dev, blocks.reshape(-1), torch.float64).reshape(-1, 3, 3)
        G, H = g_key // n_agg, g_key % n_agg
        nbG = torch.bincount(G, minlength=n_agg)
        boff = torch.cumsum(nbG, 0) - nbG
        slotG = torch.arange(G.numel(), device=dev) - boff[G]
        i3 = torch.arange(3, device=dev)
        posg = (9 * boff[G])[:, None, None] + i3[:, None] * (3 * nbG[G])[:, None, None] + (3 * slotG)[:, None, None] + i3[None, :]
        nnz = int(9 * G.numel())
        indices = torch.empty(nnz, dtype=torch.int32, device=dev)
        data = torch.empty(nnz, dtype=torch.float64, device=dev)
        indices[posg.reshape(-1)] = ((3 * H)[:, None, None] + i3[None, :].expand(3, 3)).reshape(-1).to(torch.int32)
        data[posg.reshape(-1)] = g_blk.reshape(-1)
        indptr = torch.empty(3 * n_agg + 1, dtype=torch.int32, device=dev)
        rows = (9 * boff)[:, None] + i3[None, :] * (3 * nbG)[:, None]
        indptr[:-1] = rows.reshape(-1).to(torch.int32)
        indptr[-1] = nnz
        return indptr, indices, data, bb_agg, tr[lo:lo + E_loc].contiguous()

    def gather_solution(self) -> np.ndarray:
        """The global coefficient vector (element order of the unpartitioned geometry) on every rank."""
        import torch
        import torch.distributed as dist
        d = self.dg.dim_elem
        sizes = [int(self.bounds[r + 1] - self.bounds[r]) * d for r in range(self.world)]
        on_gpu = dist.get_backend() == "nccl"
        dev = self.dg._dev["data"].device if on_gpu else "cpu"
        mine = torch.from_numpy(np.ascontiguousarray(self.solution)).to(dev)
        outs = [torch.empty(s, dtype=torch.float64, device=dev) for s in sizes]
        dist.all_gather(outs, mine) if len(set(sizes)) == 1 else self._all_gather_uneven(dist, outs, mine)
        sol = torch.cat(outs).cpu().numpy()
        if self._perm is not None:
            out = np.empty_like(sol)
            out.reshape(-1, d)[self._perm] = sol.reshape(-1, d)
            sol = out
        return sol

    def _all_gather_uneven(self, dist, outs, mine):
        for r in range(self.world):
            if r == self.rank:
                outs[r].copy_(mine)
            dist.broadcast(outs[r], src=r)
