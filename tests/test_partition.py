"""Multi-GPU host logic on CPU: element partitioning, ghost layers, halo plans, and a world_size-2 gloo exchange."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import reyna_b200 as rb
from reyna_b200.partition import HaloPlan, PartitionedGeometry, partition_bounds, partition_geometry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def geo():
    return rb.DGFEMGeometry(rb.tiled_voronoi_mesh(1024, tiles=2, seed=3), subtriangulation="fan")


@pytest.mark.parametrize("world", [2, 3, 8])
def test_partition_covers_mesh_once(geo, world):
    fg = rb.flat_geometry(geo)
    bounds = partition_bounds(geo.n_elements, world)
    assert bounds[0] == 0 and bounds[-1] == geo.n_elements and np.all(np.diff(bounds) > 0)
    n_tri = n_bf = 0
    face_hits = np.zeros(fg.n_iface, dtype=int)
    for r in range(world):
        pg = partition_geometry(geo, bounds[r], bounds[r + 1])
        f = pg._rb_flat
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        assert f.n_rows == hi - lo and f.n_elem == f.n_rows + pg.ghost_gid.shape[0]
        assert np.array_equal(f.elem_gid[:f.n_rows], np.arange(lo, hi))
        assert np.all(np.diff(f.elem_gid[f.n_rows:]) > 0)                       # ghosts ascending, none owned
        assert not np.any((pg.ghost_gid >= lo) & (pg.ghost_gid < hi))
        # every local face has at least one owned side and maps back to the global face
        assert np.all((f.ielem < f.n_rows).any(axis=1))
        assert np.array_equal(f.elem_gid[f.ielem], fg.ielem[pg.face_gid])
        assert np.array_equal(f.inorm, fg.inorm[pg.face_gid]) and np.array_equal(f.iedge, fg.iedge[pg.face_gid])
        face_hits[pg.face_gid] += (f.ielem < f.n_rows).sum(axis=1)
        # triangles and boundary faces of owned elements only, geometry of ghosts replicated
        assert np.array_equal(f.tri_elem + lo, fg.tri_elem[fg.tri_off[lo]:fg.tri_off[hi]])
        assert np.array_equal(f.bbox, fg.bbox[f.elem_gid]) and np.array_equal(f.area, fg.area[f.elem_gid])
        for k in (0, f.n_elem - 1):
            g = f.elem_gid[k]
            assert np.array_equal(f.poly_vtx[f.poly_off[k]:f.poly_off[k + 1]], fg.poly_vtx[fg.poly_off[g]:fg.poly_off[g + 1]])
        assert np.array_equal(f.belem + lo, fg.belem[pg.bface_gid])
        n_tri += f.n_tri
        n_bf += f.n_bface
    assert n_tri == fg.n_tri and n_bf == fg.n_bface
    assert np.all(face_hits == 2)        # every (face, side) item is assembled by exactly one rank


@pytest.mark.parametrize("world", [2, 4])
def test_halo_plans_agree_between_ranks(geo, world):
    bounds = partition_bounds(geo.n_elements, world)
    plans = [HaloPlan(geo, bounds, r) for r in range(world)]
    for r, p in enumerate(plans):
        pg = partition_geometry(geo, bounds[r], bounds[r + 1])
        assert p.n_ghost == pg.ghost_gid.shape[0] and np.array_equal(p.ghost_gid, pg.ghost_gid)
        got = np.sort(np.concatenate([v for v in p.recv.values()])) if p.recv else np.zeros(0, int)
        assert np.array_equal(got, p.n_rows + np.arange(p.n_ghost))             # every ghost filled exactly once
        for q, idx in p.send.items():
            sent_gid = idx + bounds[r]
            recv_gid = plans[q].ghost_gid[plans[q].recv[r] - plans[q].n_rows]
            assert np.array_equal(sent_gid, recv_gid)                           # same elements, same order


def _gloo_worker(rank, world, port, n_elements, d):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(n_elements, tiles=2, seed=3), subtriangulation="fan")
        bounds = partition_bounds(geo.n_elements, world)
        plan = HaloPlan(geo, bounds, rank)
        x = torch.full(((plan.n_rows + plan.n_ghost) * d,), -1.0, dtype=torch.float64)
        own_gid = torch.arange(int(bounds[rank]), int(bounds[rank + 1]), dtype=torch.float64)
        x.view(-1, d)[:plan.n_rows] = own_gid[:, None] * d + torch.arange(d, dtype=torch.float64)[None]
        plan.exchange_torch(x, d)
        want = torch.as_tensor(plan.ghost_gid, dtype=torch.float64)[:, None] * d + torch.arange(d, dtype=torch.float64)[None]
        assert torch.equal(x.view(-1, d)[plan.n_rows:], want)
        # and the all-reduce the Krylov dots need
        s = torch.tensor([float(plan.n_rows)], dtype=torch.float64)
        dist.all_reduce(s)
        assert int(s.item()) == geo.n_elements
    finally:
        dist.destroy_process_group()


def test_gloo_halo_exchange_world2():
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, 1024, 6), nprocs=2, join=True)


def _level1_worker(rank, world, port, n_elements, p):
    """The replicated first agglomerated multigrid level that DistributedDGFEM gathers, on CPU tensors over gloo, against the
    scipy restatement of the hierarchy (oracle/mg_reference.py)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types
        import scipy.sparse as sp
        from oracle import dg_oracle
        from oracle.mg_reference import AgglomeratedP1
        from reyna_b200.distributed import DistributedDGFEM
        from reyna_b200.problems import PROBLEMS
        geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(n_elements, tiles=2, seed=3), subtriangulation="fan")
        d = (p + 1) * (p + 2) // 2
        B, _, _ = dg_oracle.assemble(geo, p, **PROBLEMS["diffusion"]["data"])
        bsr = sp.csr_matrix(B).tobsr(blocksize=(d, d))
        bsr.sort_indices()
        bounds = partition_bounds(geo.n_elements, world)
        assert np.all(bounds[:-1] % 4 == 0)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        # this rank's block rows as the structural scalar CSR the library assembles (complete blocks, global columns)
        indptr, indices, data = [0], [], []
        for e in range(lo, hi):
            k0, k1 = bsr.indptr[e], bsr.indptr[e + 1]
            cols = bsr.indices[k0:k1]
            for i in range(d):
                indices.append((cols[:, None] * d + np.arange(d)[None]).ravel())
                data.append(bsr.data[k0:k1, i, :].ravel())
                indptr.append(indptr[-1] + (k1 - k0) * d)
        dv = dict(d=d, n=(hi - lo) * d,
                  structural=(torch.tensor(indptr, dtype=torch.int32), torch.from_numpy(np.concatenate(indices).astype(np.int32)),
                              torch.from_numpy(np.concatenate(data))))
        ddg = object.__new__(DistributedDGFEM)
        ddg.rank, ddg.world, ddg.bounds, ddg.global_geometry = rank, world, bounds, geo
        ddg.dg = types.SimpleNamespace(polydegree=p)
        l1_indptr, l1_indices, l1_data, bb_agg, tr = ddg._global_level1(torch, torch.device("cpu"), dv)
        ref = AgglomeratedP1(B, p, rb.flat_geometry(geo).bbox)
        A1 = ref.levels[1]["A"]
        got = sp.csr_matrix((l1_data.numpy(), l1_indices.numpy(), l1_indptr.numpy()), shape=A1.shape)
        diff = abs(got - A1)
        assert diff.max() <= 1e-12 * abs(A1).max(), diff.max()
        assert got.nnz == A1.nnz                                # complete 3 x 3 blocks on both sides
        P0, bb_c = AgglomeratedP1._prolongation(np.asarray(rb.flat_geometry(geo).bbox, dtype=float).reshape(-1, 4))
        assert np.allclose(bb_agg.numpy(), bb_c, rtol=0, atol=0)
        Pd = P0.toarray()
        for k, e in enumerate(range(lo, hi)):
            G = e // 4
            blk = Pd[3 * e:3 * e + 3, 3 * G:3 * G + 3]
            want = np.array([blk[0, 0], blk[1, 1], blk[0, 1], blk[2, 2], blk[0, 2]])
            assert np.allclose(tr[k, :5].numpy(), want, rtol=1e-14, atol=0)
    finally:
        dist.destroy_process_group()


def test_gloo_global_multigrid_level_world2():
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_level1_worker, args=(2, port, 256, 2), nprocs=2, join=True)


def test_partition_bounds_keep_aggregates_on_one_rank():
    for n, world in ((1000, 3), (1048576, 8), (1023, 2), (17, 4)):
        b = partition_bounds(n, world)
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0) and np.all(b[:-1] % 4 == 0)
