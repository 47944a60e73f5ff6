"""Host-side logic on CPU: geometry (against the reference's arrays in the golden fixtures), quadrature, boundary
classification, batched sampling, the reference's API validation behaviour, mesh generators and the geometry cache."""
import warnings

import numpy as np
import pytest

import reyna_b200 as rb
from reyna_b200 import sampling
from reyna_b200.problems import PROBLEMS
from reyna_b200.quadrature import basis_index_2d, reference_quadrature
from tests import util


@pytest.mark.parametrize("mesh", util.golden_names())
def test_geometry_matches_reference_arrays(mesh):
    z, ref_geo = util.load_golden(mesh)
    off, idx = z["g_region_offsets"], z["g_region_indices"]
    pm = rb.PolyMesh(z["g_vertices"].copy(), [idx[off[i]:off[i + 1]] for i in range(off.shape[0] - 1)],
                     z["g_filtered_points"].copy(), None)
    g = rb.DGFEMGeometry(pm)
    assert g.n_elements == ref_geo.n_elements and g.n_triangles == ref_geo.n_triangles
    assert np.array_equal(g.nodes, z["g_nodes"])
    assert np.array_equal(np.asarray(g.elem_bounding_boxes), z["g_elem_bounding_boxes"])
    assert np.array_equal(g.subtriangulation, z["g_subtriangulation"])
    assert np.array_equal(np.asarray(g.triangle_to_polygon), z["g_triangle_to_polygon"])
    assert np.array_equal(g.interior_edges, z["g_interior_edges"])
    assert np.array_equal(g.boundary_edges, z["g_boundary_edges"])
    assert np.array_equal(g.interior_edges_to_element, z["g_interior_edges_to_element"])
    assert np.array_equal(g.boundary_edges_to_element, z["g_boundary_edges_to_element"])
    assert np.allclose(g.interior_normals, z["g_interior_normals"], rtol=0, atol=1e-15)
    assert np.allclose(g.boundary_normals, z["g_boundary_normals"], rtol=0, atol=1e-15)
    assert np.allclose(g.areas, z["g_areas"], rtol=1e-13)
    assert np.allclose(g.h_s, z["g_h_s"], rtol=1e-10)
    assert abs(g.h - float(z["g_h"])) <= 1e-10 * g.h


def test_fan_subtriangulation_covers_elements():
    g = rb.DGFEMGeometry(rb.voronoi_mesh(300, seed=2), subtriangulation="fan")
    V = g.nodes[g.subtriangulation]
    a = 0.5 * np.abs((V[:, 1, 0] - V[:, 0, 0]) * (V[:, 2, 1] - V[:, 0, 1]) - (V[:, 1, 1] - V[:, 0, 1]) * (V[:, 2, 0] - V[:, 0, 0]))
    per_elem = np.bincount(g.triangle_to_polygon, weights=a, minlength=g.n_elements)
    assert np.allclose(per_elem, g.areas, rtol=1e-12)
    assert abs(g.areas.sum() - 1.0) < 1e-12
    assert np.all(np.diff(g.triangle_to_polygon) >= 0)


def test_geometry_invariants_like_reference_tests():
    # tests/test_geometry.py:11-138 of the reference, on our generator
    g = rb.DGFEMGeometry(rb.voronoi_mesh(200, seed=3))
    assert g.n_elements == 200 and len(g.mesh.filtered_regions) == 200
    bb = g.elem_bounding_boxes
    for e in range(g.n_elements):
        v = g.nodes[g.mesh.filtered_regions[e]]
        assert bb[e][0] == v[:, 0].min() and bb[e][1] == v[:, 0].max() and bb[e][2] == v[:, 1].min() and bb[e][3] == v[:, 1].max()
    mid = 0.5 * (g.nodes[g.boundary_edges[:, 0]] + g.nodes[g.boundary_edges[:, 1]])
    assert np.all(np.minimum(np.minimum(mid[:, 0], 1 - mid[:, 0]), np.minimum(mid[:, 1], 1 - mid[:, 1])) < 1e-12)
    for f in range(g.interior_edges.shape[0]):
        e0, e1 = g.interior_edges_to_element[f]
        assert e0 < e1
        for v in g.interior_edges[f]:
            assert v in g.mesh.filtered_regions[e0] and v in g.mesh.filtered_regions[e1]
        # unit normal from e0 to e1, orthogonal to the edge
        n = g.interior_normals[f]
        t = g.nodes[g.interior_edges[f, 1]] - g.nodes[g.interior_edges[f, 0]]
        assert abs(np.dot(n, t)) < 1e-12 and abs(np.linalg.norm(n) - 1) < 1e-12
        assert np.dot(n, g.mesh.filtered_points[e1] - g.mesh.filtered_points[e0]) > 0
    for f in range(g.boundary_edges.shape[0]):
        e = g.boundary_edges_to_element[f]
        assert np.dot(g.boundary_normals[f], mid[f] - g.mesh.filtered_points[e]) > 0


def test_tiled_mesh_is_conforming():
    m = rb.tiled_voronoi_mesh(4096, tiles=4, seed=3)
    g = rb.DGFEMGeometry(m, subtriangulation="fan")
    n_edges = g.interior_edges.shape[0] + g.boundary_edges.shape[0]
    assert m.vertices.shape[0] - n_edges + g.n_elements == 1          # Euler characteristic of a disc
    assert abs(g.areas.sum() - 1.0) < 1e-12
    with pytest.raises(ValueError):
        rb.tiled_voronoi_mesh(1000, tiles=4)


def test_geometry_cache_roundtrip(tmp_path):
    g = rb.DGFEMGeometry(rb.voronoi_mesh(128, seed=4))
    path = str(tmp_path / "geo.npz")
    rb.save_geometry_cache(g, path)
    g2 = rb.load_geometry_cache(path)
    f, f2 = rb.flat_geometry(g), rb.flat_geometry(g2)
    for k, v in vars(f).items():
        assert np.array_equal(v, getattr(f2, k)), k
    assert g2.h == g.h


def test_save_geometry_pickle_flag_is_inverted_like_reference(tmp_path):
    import pickle
    g = rb.DGFEMGeometry(rb.voronoi_mesh(16, seed=4))
    g.save_geometry(str(tmp_path / "a.pkl"), save_mesh=True)
    assert "mesh" not in pickle.load(open(tmp_path / "a.pkl", "rb"))      # DGFEM.py:267
    g.save_geometry(str(tmp_path / "b.pkl"), save_mesh=False)
    assert "mesh" in pickle.load(open(tmp_path / "b.pkl", "rb"))


def test_quadrature_and_orders_shapes():
    for p in (1, 2, 3):
        (w, pts), (w1, x1) = reference_quadrature(p)
        assert w.shape == ((p + 1) ** 2, 1) and pts.shape == ((p + 1) ** 2, 2)
        assert w1.shape == (p + 1, 1) and x1.shape == (p + 1, 1)
        assert abs(w.sum() - 4.0) < 1e-14 and abs(w1.sum() - 2.0) < 1e-14
        assert np.all(pts >= 0) and np.all(pts.sum(axis=1) <= 1)
        o = basis_index_2d(p)
        assert o.shape == ((p + 1) * (p + 2) // 2, 2) and o.dtype == np.int64
    assert basis_index_2d(2).tolist() == [[0, 0], [0, 1], [0, 2], [1, 0], [1, 1], [2, 0]]
    with pytest.raises(ValueError):
        basis_index_2d(-1)


def test_quadrature_identical_to_oracle():
    from oracle import dg_oracle
    for p in (1, 2, 3):
        (w, pts), (w1, x1) = reference_quadrature(p)
        ow, oref, ow1, ox1 = dg_oracle.quadrature(p)
        assert np.array_equal(w[:, 0], ow) and np.array_equal(pts, oref)
        assert np.array_equal(w1[:, 0], ow1) and np.array_equal(x1[:, 0], ox1)


def test_add_data_validation_matches_reference():
    g = rb.DGFEMGeometry(rb.voronoi_mesh(16, seed=1))
    ones = lambda x: np.ones(x.shape[0], dtype=np.float64)   # noqa: E731
    with pytest.raises(ValueError, match="No advection or diffusion"):            # tests/test_DGFEM.py:19-29
        rb.DGFEM(g).add_data(reaction=ones, dirichlet_bcs=ones)
    with pytest.raises(ValueError, match="Dirichlet"):                            # tests/test_DGFEM.py:31-43
        rb.DGFEM(g).add_data(advection=lambda x: np.ones(x.shape, dtype=np.float64))
    dg = rb.DGFEM(g, polynomial_degree=0)                                         # tests/test_DGFEM.py:45-62
    with pytest.warns(UserWarning, match="polynomial degree given is too low"):
        dg.add_data(diffusion=PROBLEMS["adr"]["data"]["diffusion"], dirichlet_bcs=ones)
    assert dg.polydegree == 1
    with pytest.raises(ValueError):
        rb.DGFEM(g).errors(exact_solution=ones)
    with pytest.raises(ValueError):
        rb.DGFEM(g).plot_DG()


@pytest.mark.parametrize("mesh,key", util.golden_cases()[:6])
def test_boundary_classification_matches_reference(mesh, key):
    z, geo = util.load_golden(mesh)
    name, p = util.split_case(key)
    dg = rb.DGFEM(geo, polynomial_degree=p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dg.add_data(**PROBLEMS[name]["data"])
    bi = dg.boundary_information
    data = PROBLEMS[name]["data"]
    if data.get("diffusion") is not None:
        assert np.array_equal(bi.elliptical_indecies, z[key + "_elliptic"])
    else:
        assert bi.elliptical_indecies is None
    if data.get("advection") is not None:
        assert np.array_equal(bi.inflow_indecies, z[key + "_inflow"])
    else:
        assert bi.inflow_indecies is None
    assert bi.tolerence == 1e-12


def test_batched_sampling_is_chunk_invariant(monkeypatch):
    from numba import njit, f8
    f = njit(f8[:](f8[:, :]), nogil=True)(PROBLEMS["adr"]["data"]["forcing"])
    a = njit(f8[:, :, :](f8[:, :]), nogil=True)(PROBLEMS["variable"]["data"]["diffusion"])
    X = np.random.default_rng(0).random((50000, 2))
    monkeypatch.setattr(sampling, "_CHUNK", 1 << 30)
    whole_f, whole_a = sampling.evaluate_batched(f, X, ()), sampling.evaluate_batched(a, X, (2, 2))
    monkeypatch.setattr(sampling, "_CHUNK", 777)
    monkeypatch.setenv("REYNA_B200_HOST_THREADS", "4")
    assert np.array_equal(sampling.evaluate_batched(f, X, ()), whole_f)          # bitwise
    assert np.array_equal(sampling.evaluate_batched(a, X, (2, 2)), whole_a)
    # and identical to the reference's per-item calls on (p+1)^2 points
    per_item = np.concatenate([f(np.ascontiguousarray(X[i:i + 9])) for i in range(0, 900, 9)])
    assert np.array_equal(per_item, whole_f[:900])


def test_submitted_jobs_equal_batched_evaluation(monkeypatch):
    """submit_terms (jobs queued up front on the thread pool, consumed in order: the dgfem() pipeline) writes bit for bit
    what evaluate_batched writes, for ranges, several callables per task and a callable that raises."""
    from numba import njit, f8
    f = njit(f8[:](f8[:, :]), nogil=True)(PROBLEMS["adr"]["data"]["forcing"])
    a = njit(f8[:, :, :](f8[:, :]), nogil=True)(PROBLEMS["variable"]["data"]["diffusion"])
    X = np.random.default_rng(5).random((9001, 2))
    monkeypatch.setattr(sampling, "_CHUNK", 1 << 30)
    whole_f, whole_a = sampling.evaluate_batched(f, X, ()), sampling.evaluate_batched(a, X, (2, 2))
    monkeypatch.setattr(sampling, "_CHUNK", 640)
    monkeypatch.setenv("REYNA_B200_HOST_THREADS", "3")
    out_f, out_a = np.full(9001, np.nan), np.full((9001, 2, 2), np.nan)
    jobs = [sampling.submit_terms([(f, (), out_f), (a, (2, 2), out_a)], X, lo, hi) for lo, hi in ((0, 4000), (4000, 4100), (4100, 9001))]
    for j in jobs:
        j.wait()
    assert np.array_equal(out_f, whole_f) and np.array_equal(out_a, whole_a)
    assert sampling.submit_terms([(f, (), out_f)], X, 10, 10)._futures == ()      # empty range

    def bad(x):
        raise RuntimeError("callable failed")
    with pytest.raises(RuntimeError, match="callable failed"):
        sampling.submit_terms([(bad, (), out_f)], X).wait()
    # after a failure the other jobs of the call are abandoned: cancelled or finished, never left running, no exception
    rest = [sampling.submit_terms([(bad, (), out_f)], X), sampling.submit_terms([(f, (), out_f)], X)]
    for j in rest:
        j.abandon()
    assert all(j._futures == () for j in rest)


def test_host_threads_are_shared_between_the_ranks_of_a_node(monkeypatch):
    monkeypatch.delenv("REYNA_B200_HOST_THREADS", raising=False)
    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    n1 = sampling.host_threads()
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "2")
    assert sampling.host_threads() == max(1, min(32, len(__import__("os").sched_getaffinity(0)) // 2)) <= n1


def test_sampling_points_match_oracle_layout():
    from oracle import dg_oracle
    g = rb.DGFEMGeometry(rb.voronoi_mesh(40, seed=6))
    fg = rb.flat_geometry(g)
    og = dg_oracle.flatten_geometry(g)
    (w, ref2), (w1, x1) = reference_quadrature(2)
    Xv = sampling.volume_points(fg, ref2).reshape(9, fg.n_tri, 2)
    Xo, _ = dg_oracle.triangle_points(og, ref2, slice(0, fg.n_tri))
    assert np.allclose(Xv.transpose(1, 0, 2), Xo, rtol=0, atol=1e-15)
    mid, Xe = sampling.edge_points(fg.nodes, fg.iedge, x1)
    omid, _, Xeo, _ = dg_oracle.edge_points(og.nodes, og.iedge, x1[:, 0])
    assert np.array_equal(mid, omid)
    assert np.allclose(Xe.reshape(3, fg.n_iface, 2).transpose(1, 0, 2), Xeo, rtol=0, atol=1e-15)


def test_morton_order_is_a_permutation_and_local():
    pts = np.random.default_rng(1).random((4096, 2))
    perm = rb.morton_order(pts)
    assert np.array_equal(np.sort(perm), np.arange(4096))
    q = pts[perm]
    # consecutive points along the curve are close on average (locality)
    assert np.mean(np.linalg.norm(np.diff(q, axis=0), axis=1)) < 0.05
