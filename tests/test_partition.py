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
