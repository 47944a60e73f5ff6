"""Parity of the CUDA path (through DGFEM -> ctypes -> libreyna_b200.so) with the reference.

* golden fixtures (outputs of the unmodified reference, tests/golden/) -- every case;
* the CPU oracle on larger synthetic meshes (same geometry object handed to both);
* size-independent properties at benchmark scale (polynomial exactness from the reference's own test-suite,
  symmetry of the pure-diffusion matrix, B @ 1 identities, determinism, CSR canonical form).
Tolerances are the north star's: pattern identical (up to rounding-noise entries, see tests/util.compare_csr), entries
1e-12 block-relative, solution 1e-9 relative L2, errors to 3 significant digits.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from reyna_b200.problems import PROBLEMS
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import reyna_b200
    reyna_b200.lib()   # raises if the extension is missing: no silent fallback
    return reyna_b200


def _run(rb, geo, name, p, solve=True, **opts):
    dg = rb.DGFEM(geo, polynomial_degree=p)
    dg.add_data(**PROBLEMS[name]["data"])
    dg.solver_options.update(opts)
    dg.dgfem(solve=solve)
    return dg


@pytest.mark.parametrize("mesh,key", util.golden_cases())
def test_golden_parity(rb, mesh, key):
    z, geo = util.load_golden(mesh)
    name, p = util.split_case(key)
    Bref, Lref, xref = util.golden_system(z, key)
    dg = _run(rb, geo, name, p)
    B = dg.B
    assert B.indices.dtype == np.int32 and B.indptr.dtype == np.int32
    assert B.has_canonical_format
    max_rel, bad, noise = util.compare_csr(B, Bref, dg.dim_elem)
    util.note(f"[parity] golden {mesh} {key}: max block-relative error {max_rel:.2e}, noise-level pattern entries {noise}")
    assert bad == 0, f"{bad} pattern mismatches that are not rounding noise"
    assert max_rel <= util.TOL_ENTRY
    if noise == 0:
        assert np.array_equal(B.indptr, Bref.indptr) and np.array_equal(B.indices, Bref.indices)
    L = np.asarray(dg.L.todense()).reshape(-1)
    assert util.rel_inf(L, Lref) <= util.TOL_ENTRY
    assert dg.L.shape == (dg.dim_system, 1)
    assert util.rel_l2(dg.solution, xref) <= util.TOL_SOLUTION, dg.solve_info
    if key + "_errors" in z.files:
        pr = PROBLEMS[name]
        has_diff = pr["data"].get("diffusion") is not None
        e = dg.errors(exact_solution=pr["exact"], grad_exact_solution=pr["grad_exact"] if has_diff else None)
        eref = z[key + "_errors"]
        assert abs(e[0] - eref[0]) <= util.TOL_ERRORS * abs(eref[0])
        assert abs(e[1] - eref[1]) <= util.TOL_ERRORS * abs(eref[1])
        if has_diff:
            assert abs(e[2] - eref[2]) <= util.TOL_ERRORS * abs(eref[2])
        else:
            assert e[2] is None


@pytest.mark.parametrize("E,name,p", [(3000, "adr", 2), (3000, "variable", 3), (5000, "variable_hyperbolic", 2),
                                      (5000, "diffusion", 1), (2000, "hyperbolic", 3), (2000, "adr_tensor", 1)])
def test_oracle_parity_synthetic(rb, E, name, p):
    from oracle import dg_oracle
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(E, seed=E + p), subtriangulation="fan")
    dg = _run(rb, geo, name, p)
    B, L, info = dg_oracle.assemble(geo, p, **PROBLEMS[name]["data"])
    max_rel, bad, noise = util.compare_csr(dg.B, B, info.dim_elem)
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_inf(np.asarray(dg.L.todense()), np.asarray(L.todense())) <= util.TOL_ENTRY
    x = dg_oracle.solve(B, L)
    assert util.rel_l2(dg.solution, x) <= util.TOL_SOLUTION, dg.solve_info
    pr = PROBLEMS[name]
    if pr["exact"] is not None:
        has_diff = pr["data"].get("diffusion") is not None
        e = dg.errors(exact_solution=pr["exact"], grad_exact_solution=pr["grad_exact"] if has_diff else None)
        eo = dg_oracle.errors(geo, p, x, pr["exact"], pr["grad_exact"] if has_diff else None,
                              advection=pr["data"].get("advection"), diffusion=pr["data"].get("diffusion"),
                              reaction=pr["data"].get("reaction"))
        for a, b in zip(e, eo):
            if b is not None:
                assert abs(a - b) <= util.TOL_ERRORS * abs(b)


def test_tiled_and_delaunay_subtriangulation(rb):
    """Mirrored-tile mesh (merged side vertices) with the reference's per-polygon Delaunay sub-triangulation."""
    from oracle import dg_oracle
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(1024, tiles=2, seed=5))
    dg = _run(rb, geo, "variable", 2)
    B, L, info = dg_oracle.assemble(geo, 2, **PROBLEMS["variable"]["data"])
    max_rel, bad, noise = util.compare_csr(dg.B, B, info.dim_elem)
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_l2(dg.solution, dg_oracle.solve(B, L)) <= util.TOL_SOLUTION


# ---- the reference's polynomial-exactness known-answer tests (tests/test_DGFEM.py:85-479), at benchmark-like size --------
def _identity(x):
    out = np.zeros((x.shape[0], 2, 2), dtype=np.float64)
    for i in range(x.shape[0]):
        out[i, 0, 0] = 1.0
        out[i, 1, 1] = 1.0
    return out


def _nondiag(x):
    out = np.zeros((x.shape[0], 2, 2), dtype=np.float64)
    for i in range(x.shape[0]):
        out[i, 0, 0] = 2.0
        out[i, 0, 1] = 1.0
        out[i, 1, 0] = 1.0
        out[i, 1, 1] = 3.0
    return out


def test_exactness_advection_linear_p1(rb):
    # tests/test_DGFEM.py:109-133: b = (1,1), u = x + y, f = 2
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(20000, seed=1), subtriangulation="fan")
    dg = rb.DGFEM(geo, polynomial_degree=1)
    dg.add_data(advection=lambda x: np.ones(x.shape, dtype=np.float64),
                forcing=lambda x: 2.0 * np.ones(x.shape[0], dtype=np.float64),
                dirichlet_bcs=lambda x: x[:, 0] + x[:, 1])
    dg.dgfem(solve=True)
    l2, dgn, h1 = dg.errors(exact_solution=lambda x: x[:, 0] + x[:, 1])
    assert h1 is None and l2 < 1e-10 and dgn < 1e-10


def test_exactness_advection_reaction_quadratic_p2(rb):
    # tests/test_DGFEM.py:215-243: b = (1,1), c = 1, u = x^2 + y^2
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(20000, seed=2), subtriangulation="fan")
    dg = rb.DGFEM(geo, polynomial_degree=2)
    dg.add_data(advection=lambda x: np.ones(x.shape, dtype=np.float64),
                reaction=lambda x: np.ones(x.shape[0], dtype=np.float64),
                forcing=lambda x: 2.0 * x[:, 0] + 2.0 * x[:, 1] + x[:, 0] ** 2 + x[:, 1] ** 2,
                dirichlet_bcs=lambda x: x[:, 0] ** 2 + x[:, 1] ** 2)
    dg.dgfem(solve=True)
    l2, dgn, _ = dg.errors(exact_solution=lambda x: x[:, 0] ** 2 + x[:, 1] ** 2)
    assert l2 < 1e-10 and dgn < 1e-10


def test_exactness_diffusion_quadratic_p2(rb):
    # tests/test_DGFEM.py:335-367: a = I, u = x^2 + y^2, f = -4
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(4000, seed=3), subtriangulation="fan")
    dg = rb.DGFEM(geo, polynomial_degree=2)
    dg.add_data(diffusion=_identity, forcing=lambda x: -4.0 * np.ones(x.shape[0], dtype=np.float64),
                dirichlet_bcs=lambda x: x[:, 0] ** 2 + x[:, 1] ** 2)
    dg.dgfem(solve=True)
    l2, dgn, h1 = dg.errors(exact_solution=lambda x: x[:, 0] ** 2 + x[:, 1] ** 2,
                            grad_exact_solution=lambda x: 2.0 * x)
    assert l2 < 1e-10 and dgn < 1e-8 and h1 < 1e-8, (l2, dgn, h1, dg.solve_info)


def test_exactness_nondiagonal_tensor_p2(rb):
    # tests/test_DGFEM.py:405-479 (full non-diagonal tensor): a = [[2,1],[1,3]], u = x^2 + x y, -div(a grad u) = -(4 + 2) = -6
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(4000, seed=4), subtriangulation="fan")
    dg = rb.DGFEM(geo, polynomial_degree=2)
    u = lambda x: x[:, 0] ** 2 + x[:, 0] * x[:, 1]                                          # noqa: E731
    gu = lambda x: np.vstack((2.0 * x[:, 0] + x[:, 1], x[:, 0])).T                           # noqa: E731
    dg.add_data(diffusion=_nondiag, forcing=lambda x: -6.0 * np.ones(x.shape[0], dtype=np.float64), dirichlet_bcs=u)
    dg.dgfem(solve=True)
    l2, dgn, h1 = dg.errors(exact_solution=u, grad_exact_solution=gu)
    assert l2 < 1e-10 and dgn < 1e-8 and h1 < 1e-8, (l2, dgn, h1, dg.solve_info)


# ---- structural properties at scale -----------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def big(rb):
    return rb.DGFEMGeometry(rb.tiled_voronoi_mesh(65536, tiles=4, seed=1337), subtriangulation="fan")


def test_scale_diffusion_symmetric_and_full_pattern(rb, big):
    dg = _run(rb, big, "diffusion", 2, solve=False)
    B = dg.B
    d, E, F = dg.dim_elem, big.n_elements, big.interior_edges.shape[0]
    assert B.nnz == d * d * (E + 2 * F)            # full pattern whenever diffusion is present
    assert B.has_canonical_format and B.indices.dtype == np.int32
    A = (B - B.T).tocoo()
    assert np.abs(A.data).max() <= 1e-12 * np.abs(B.data).max()       # SIPG with symmetric a is symmetric
    # constants are in the kernel of the volume + interior-face diffusion form: (B @ 1)_(e,j) only sees Dirichlet faces
    one = np.zeros(dg.dim_system)
    hx = 0.5 * (big.elem_bounding_boxes[:, 1] - big.elem_bounding_boxes[:, 0])
    hy = 0.5 * (big.elem_bounding_boxes[:, 3] - big.elem_bounding_boxes[:, 2])
    one[0::d] = 2.0 * np.sqrt(hx * hy)             # coefficient of phi_0 = 1/(2 sqrt(hx hy)) representing u = 1
    r = (B @ one).reshape(E, d)
    interior = np.ones(E, dtype=bool)
    interior[big.boundary_edges_to_element] = False
    assert np.abs(r[interior]).max() <= 1e-9 * np.abs(B.data).max()


def test_scale_hyperbolic_one_sided_pattern(rb, big):
    dg = _run(rb, big, "hyperbolic", 2, solve=False)
    d, E, F = dg.dim_elem, big.n_elements, big.interior_edges.shape[0]
    # advection-only: one off-diagonal block per face, minus rounding-noise zeros
    assert abs(dg.B.nnz - d * d * (E + F)) <= 1e-3 * d * d * (E + F)
    assert dg.B.has_canonical_format
    assert dg.timings["exact_zeros"] >= 0


def test_scale_deterministic(rb, big):
    a = _run(rb, big, "adr", 2, solve=False)
    b = _run(rb, big, "adr", 2, solve=False)
    assert np.array_equal(a.B.indptr, b.B.indptr) and np.array_equal(a.B.indices, b.B.indices)
    assert np.array_equal(a.B.data, b.B.data)                          # bitwise: no floating-point atomics anywhere
    assert np.array_equal(np.asarray(a.L.todense()), np.asarray(b.L.todense()))


def test_scale_oracle_parity_65k_p2(rb, big):
    """BASELINE config 3a (65k elements, p=2 ADR) entry by entry against the oracle."""
    from oracle import dg_oracle
    dg = _run(rb, big, "adr", 2, solve=False)
    B, L, info = dg_oracle.assemble(big, 2, **PROBLEMS["adr"]["data"])
    max_rel, bad, noise = util.compare_csr(dg.B, B, info.dim_elem)
    util.note(f"[parity] 65536 elements adr p=2: max block-relative error {max_rel:.2e}, noise-level pattern entries {noise}")
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_inf(np.asarray(dg.L.todense()), np.asarray(L.todense())) <= util.TOL_ENTRY


def test_spmv_matches_scipy(rb, big):
    import ctypes as C
    import torch
    dg = _run(rb, big, "adr", 1, solve=False)
    dv = dg._dev
    x = torch.linspace(-1.0, 1.0, dv["n"], dtype=torch.float64, device=dv["data"].device)
    y = torch.empty_like(x)
    rc = rb.lib().reyna_b200_spmv(dv["n"], dv["d"], int(dv["blocked"]), dv["indptr"].data_ptr(), dv["indices"].data_ptr(),
                                  dv["data"].data_ptr(), x.data_ptr(), y.data_ptr(),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    yref = dg.B @ x.cpu().numpy()
    assert util.rel_inf(y.cpu().numpy(), yref) <= 1e-13


def test_geometry_cache_roundtrip_on_device(rb, tmp_path):
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(500, seed=9), subtriangulation="fan")
    path = str(tmp_path / "g.npz")
    rb.save_geometry_cache(geo, path)
    geo2 = rb.load_geometry_cache(path)
    a = _run(rb, geo, "adr", 2, solve=False)
    b = _run(rb, geo2, "adr", 2, solve=False)
    assert np.array_equal(a.B.data, b.B.data) and np.array_equal(a.B.indices, b.B.indices)


def test_reference_api_errors_on_gpu(rb):
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(64, seed=1))
    dg = rb.DGFEM(geo, polynomial_degree=1)
    dg.add_data(advection=lambda x: np.ones(x.shape, dtype=np.float64),
                dirichlet_bcs=lambda x: np.ones(x.shape[0], dtype=np.float64))
    with pytest.raises(AssertionError):
        dg.dgfem(solve=True, verbose=2)                 # tests/test_DGFEM.py:64-83
    with pytest.raises(ValueError):
        dg.errors(exact_solution=lambda x: np.ones(x.shape[0]))      # no solution yet
    dg.dgfem(solve=True)
    with pytest.raises(ValueError):                     # tests/test_DGFEM.py:245-268 (mirror: grad without diffusion)
        dg.errors(exact_solution=lambda x: np.ones(x.shape[0]), grad_exact_solution=lambda x: np.zeros(x.shape))
    l2, dgn, h1 = dg.errors(exact_solution=lambda x: np.ones(x.shape[0]))
    assert l2 < 1e-10 and dgn < 1e-10 and h1 is None    # tests/test_DGFEM.py:85-107 constant solution


@pytest.mark.parametrize("world,name,p", [(3, "adr", 2), (4, "variable_hyperbolic", 1), (2, "variable", 3)])
def test_partitioned_assembly_equals_global_rows(rb, world, name, p):
    """Multi-GPU assembly path on one device: every rank's block rows, stacked, are the global matrix BITWISE (cut faces
    are computed from both sides, each rank keeping its own rows; columns carry global numbering)."""
    import scipy.sparse as sps
    from reyna_b200.partition import partition_bounds, partition_geometry
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(4096, tiles=4, seed=11), subtriangulation="fan")
    ref = _run(rb, geo, name, p, solve=False)
    bounds = partition_bounds(geo.n_elements, world)
    blocks, loads = [], []
    for r in range(world):
        part = partition_geometry(geo, bounds[r], bounds[r + 1])
        dg = _run(rb, part, name, p, solve=False)
        assert dg.B.shape == ((bounds[r + 1] - bounds[r]) * dg.dim_elem, ref.dim_system)
        blocks.append(dg.B)
        loads.append(np.asarray(dg.L.todense()).reshape(-1))
        with pytest.raises(rb.ReynaB200Error):
            dg._solve_device()
    B = sps.vstack(blocks).tocsr()
    assert np.array_equal(B.indptr, ref.B.indptr) and np.array_equal(B.indices, ref.B.indices)
    assert np.array_equal(B.data, ref.B.data)
    assert np.array_equal(np.concatenate(loads), np.asarray(ref.L.todense()).reshape(-1))


def test_evaluate_matches_oracle_basis(rb):
    """tools.evaluate (tools.py:117-153): element-wise signature and the batched form, against the oracle's basis."""
    from oracle import dg_oracle
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(300, seed=12), subtriangulation="fan")
    for p in (1, 2, 3):
        dg = _run(rb, geo, "adr", p)
        orders = dg_oracle.basis_orders(p)
        d = orders.shape[0]
        rng = np.random.default_rng(p)
        el = rng.integers(0, geo.n_elements, 500)
        bb = np.asarray(geo.elem_bounding_boxes)[el]
        # points inside the bounding boxes of their elements (the basis is defined there)
        pts = np.column_stack((bb[:, 0] + rng.random(500) * (bb[:, 1] - bb[:, 0]), bb[:, 2] + rng.random(500) * (bb[:, 3] - bb[:, 2])))
        phi, _ = dg_oracle.basis_and_grad(pts, bb, orders, need_grad=False)
        want = np.einsum("kn,nk->n", phi, dg.solution.reshape(-1, d)[el])
        got = dg.evaluate(pts, el)
        assert util.rel_inf(got, want) <= 1e-13
        e0 = int(el[0])
        mask = el == e0
        one = rb.evaluate(pts[mask], dg.solution[e0 * d:(e0 + 1) * d], np.asarray(geo.elem_bounding_boxes)[e0], orders=orders)
        assert util.rel_inf(one, want[mask]) <= 1e-13
        assert np.allclose(rb.evaluate(pts[mask], dg.solution[e0 * d:(e0 + 1) * d], np.asarray(geo.elem_bounding_boxes)[e0]), one)


def _kinked_mesh(rb):
    """2 x 2 polygons whose common borders are POLYLINES (two collinear-free edges each): neighbouring elements share two
    interior faces, so the (element, face) adjacency contains repeated neighbours that must merge into one CSR block."""
    v = np.array([[0.0, 0.0], [0.5, 0.0], [1.0, 0.0], [0.0, 0.5], [0.45, 0.25], [0.55, 0.75], [1.0, 0.5], [0.0, 1.0],
                  [0.5, 1.0], [1.0, 1.0], [0.5, 0.5], [0.25, 0.55], [0.75, 0.45]])
    regions = [np.array([0, 1, 4, 10, 11, 3]), np.array([1, 2, 6, 12, 10, 4]), np.array([3, 11, 10, 5, 8, 7]),
               np.array([10, 12, 6, 9, 8, 5])]
    pts = np.array([[0.22, 0.25], [0.78, 0.22], [0.22, 0.78], [0.78, 0.75]])
    return rb.PolyMesh(v, regions, pts, None)


@pytest.mark.parametrize("name,p", [("variable", 1), ("variable", 2), ("variable_hyperbolic", 2), ("adr", 3)])
def test_repeated_neighbours_merge_into_one_block(rb, name, p):
    from oracle import dg_oracle
    geo = rb.DGFEMGeometry(_kinked_mesh(rb))
    assert geo.interior_edges.shape[0] == 8                  # every pair of neighbours shares two faces
    dg = _run(rb, geo, name, p)
    assert dg._dev["structure"]["n_dup_items"] > 0
    B, L, info = dg_oracle.assemble(geo, p, **PROBLEMS[name]["data"])
    max_rel, bad, noise = util.compare_csr(dg.B, B, info.dim_elem)
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_inf(np.asarray(dg.L.todense()), np.asarray(L.todense())) <= util.TOL_ENTRY
    assert util.rel_l2(dg.solution, dg_oracle.solve(B, L)) <= util.TOL_SOLUTION


@pytest.mark.parametrize("name,p,chunk", [("adr", 2, 700), ("variable", 3, 1500), ("variable_hyperbolic", 1, 333),
                                          ("variable", 2, 10 ** 9)])
def test_pipelined_download_equals_lazy_materialisation(rb, name, p, chunk):
    """dgfem() with ``download_result`` (volume samples evaluated / uploaded / assembled / downloaded chunk by chunk on three
    streams) leaves, in host memory, bitwise the system a plain dgfem() + lazy ``B`` / ``L`` gives."""
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(4096, tiles=4, seed=5), subtriangulation="fan")
    ref = _run(rb, geo, name, p, solve=False)
    dg = rb.DGFEM(geo, polynomial_degree=p)
    dg.add_data(**PROBLEMS[name]["data"])
    dg.download_result = True
    dg.pipeline_chunk_elements = chunk
    for _ in range(2):                       # the second call reuses the pinned staging buffers
        dg.dgfem(solve=False)
        assert dg._host is not None
        assert np.array_equal(dg.B.indptr, ref.B.indptr) and np.array_equal(dg.B.indices, ref.B.indices)
        assert np.array_equal(dg.B.data, ref.B.data)
        assert np.array_equal(np.asarray(dg.L.todense()), np.asarray(ref.L.todense()))


def _fan_mesh(rb, n_spokes, double_edges=False):
    """One polygon with ``n_spokes`` + 3 vertices (a disc sector) surrounded by ``n_spokes`` thin neighbours: the centre element
    has many interior faces and many sub-triangles (volume segments, rounds of its own).  ``double_edges``: every border between
    the centre and a neighbour is a two-edge polyline, so every neighbour is repeated (merged blocks)."""
    th = np.linspace(0.0, 0.5 * np.pi, n_spokes + 1)
    inner = np.column_stack((0.5 * np.cos(th), 0.5 * np.sin(th)))
    outer = np.column_stack((np.cos(th), np.sin(th)))
    v = np.vstack(([[0.0, 0.0]], inner, outer))
    ni = n_spokes + 1
    pts = [[0.2, 0.2]]
    if not double_edges:
        regions = [np.concatenate(([0], 1 + np.arange(ni)))]
        for k in range(n_spokes):
            regions.append(np.array([1 + k, 1 + ni + k, 1 + ni + k + 1, 1 + k + 1]))
            pts.append(list(0.25 * (inner[k] + inner[k + 1] + outer[k] + outer[k + 1])))
        return rb.PolyMesh(v, regions, np.array(pts), None)
    thm = 0.5 * (th[:-1] + th[1:])
    mid = np.column_stack((0.53 * np.cos(thm), 0.53 * np.sin(thm)))          # kink of every border (not collinear)
    v = np.vstack((v, mid))
    m0 = 1 + 2 * ni
    centre = [0]
    for k in range(n_spokes):
        centre += [1 + k, m0 + k]
    centre.append(1 + n_spokes)
    regions = [np.array(centre)]
    for k in range(n_spokes):
        regions.append(np.array([1 + k, 1 + ni + k, 1 + ni + k + 1, 1 + k + 1, m0 + k]))
        pts.append(list(0.25 * (inner[k] + inner[k + 1] + outer[k] + outer[k + 1])))
    return rb.PolyMesh(v, regions, np.array(pts), None)


@pytest.mark.parametrize("n_spokes,name,p", [(20, "variable", 2), (31, "adr", 1), (31, "variable", 3), (29, "variable_hyperbolic", 2)])
def test_element_with_many_faces_and_triangles(rb, n_spokes, name, p):
    """A polygon with up to 31 neighbours / 32 sub-triangles (the limits of one warp round) next to small elements."""
    from oracle import dg_oracle
    geo = rb.DGFEMGeometry(_fan_mesh(rb, n_spokes))
    dg = _run(rb, geo, name, p, solve=False)
    B, L, info = dg_oracle.assemble(geo, p, **PROBLEMS[name]["data"])
    max_rel, bad, noise = util.compare_csr(dg.B, B, info.dim_elem)
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_inf(np.asarray(dg.L.todense()), np.asarray(L.todense())) <= util.TOL_ENTRY


def test_element_with_more_than_32_triangles_uses_segments(rb):
    """More than 32 sub-triangles in one element: its volume rounds are summed through segments."""
    from oracle import dg_oracle
    from reyna_b200.sampling import volume_stream_of
    geo = rb.DGFEMGeometry(_fan_mesh(rb, 30), subtriangulation="fan")
    # refine the sub-triangulation of element 0 artificially: split every triangle of it in two (same quadrature exactness)
    fg = rb.flat_geometry(geo)
    t0, t1 = int(fg.tri_off[0]), int(fg.tri_off[1])
    assert t1 - t0 >= 29
    nodes = np.asarray(geo.nodes, dtype=float)
    tri = np.asarray(geo.subtriangulation).reshape(-1, 3)
    new_nodes, new_tri = [nodes], []
    nn = nodes.shape[0]
    for t in range(t0, t1):
        a, b, c = tri[t]
        new_nodes.append(0.5 * (nodes[b] + nodes[c])[None, :])
        new_tri += [[a, b, nn], [a, nn, c]]
        nn += 1
    geo.nodes = np.vstack(new_nodes)
    geo.mesh.vertices = geo.nodes
    geo.subtriangulation = np.vstack((np.array(new_tri), tri[t1:]))
    geo.triangle_to_polygon = np.concatenate((np.zeros(len(new_tri), dtype=int), np.asarray(geo.triangle_to_polygon)[t1:]))
    geo.n_triangles = geo.subtriangulation.shape[0]
    geo.n_nodes = geo.nodes.shape[0]
    for attr in ("_rb_flat", "_rb_dev", "_rb_points"):
        if hasattr(geo, attr):
            delattr(geo, attr)
    vs = volume_stream_of(rb.flat_geometry(geo))
    assert vs.tri_seg is not None and vs.n_segs > geo.n_elements
    dg = _run(rb, geo, "variable", 2, solve=False)
    B, L, info = dg_oracle.assemble(geo, 2, **PROBLEMS["variable"]["data"])
    max_rel, bad, noise = util.compare_csr(dg.B, B, info.dim_elem)
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_inf(np.asarray(dg.L.todense()), np.asarray(L.todense())) <= util.TOL_ENTRY


@pytest.mark.parametrize("n_spokes,name,p", [(40, "variable", 2), (60, "adr", 1), (45, "variable", 3), (50, "variable_hyperbolic", 2)])
def test_element_with_33_to_64_faces_takes_the_giant_path(rb, n_spokes, name, p):
    """More than 32 (element, face) items in one element (many-sided / agglomerated polygons): dg_face_giant_kernel assembles
    its block row (and its > 32 sub-triangles go through volume segments); the neighbours stay on the fast path."""
    from oracle import dg_oracle
    geo = rb.DGFEMGeometry(_fan_mesh(rb, n_spokes))
    dg = _run(rb, geo, name, p, solve=False)
    B, L, info = dg_oracle.assemble(geo, p, **PROBLEMS[name]["data"])
    max_rel, bad, noise = util.compare_csr(dg.B, B, info.dim_elem)
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_inf(np.asarray(dg.L.todense()), np.asarray(L.todense())) <= util.TOL_ENTRY


@pytest.mark.parametrize("n_spokes,name,p", [(20, "variable", 2), (12, "adr", 3), (25, "variable_hyperbolic", 1)])
def test_giant_element_with_repeated_neighbours(rb, n_spokes, name, p):
    """2 * n_spokes items in the centre element, every neighbour twice: merged blocks inside and across the 32-item chunks."""
    from oracle import dg_oracle
    geo = rb.DGFEMGeometry(_fan_mesh(rb, n_spokes, double_edges=True))
    dg = _run(rb, geo, name, p, solve=False)
    assert dg._dev["structure"]["n_dup_items"] > 0
    B, L, info = dg_oracle.assemble(geo, p, **PROBLEMS[name]["data"])
    max_rel, bad, noise = util.compare_csr(dg.B, B, info.dim_elem)
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_inf(np.asarray(dg.L.todense()), np.asarray(L.todense())) <= util.TOL_ENTRY


def test_element_with_more_than_64_faces_is_rejected(rb):
    geo = rb.DGFEMGeometry(_fan_mesh(rb, 70))
    dg = rb.DGFEM(geo, polynomial_degree=1)
    dg.add_data(**PROBLEMS["adr"]["data"])
    with pytest.raises(rb.ReynaB200Error):
        dg.dgfem(solve=False)


@pytest.mark.parametrize("n_cells,name,p", [(1, "adr", 2), (1, "hyperbolic", 1), (3, "variable", 3), (4, "adr", 1)])
def test_tiny_meshes(rb, n_cells, name, p):
    """One to four elements (no or very few interior faces; a single interior edge is rejected like upstream leaves it undefined): windows, rounds and bulk stores at their smallest."""
    from oracle import dg_oracle
    xs = np.linspace(0.0, 1.0, n_cells + 1)
    v = np.array([[x, y] for y in (0.0, 1.0) for x in xs])
    regions = [np.array([i, i + 1, n_cells + 1 + i + 1, n_cells + 1 + i]) for i in range(n_cells)]
    pts = np.array([[0.5 * (xs[i] + xs[i + 1]), 0.5] for i in range(n_cells)])
    geo = rb.DGFEMGeometry(rb.PolyMesh(v, regions, pts, None))
    dg = _run(rb, geo, name, p, solve=False)
    B, L, info = dg_oracle.assemble(geo, p, **PROBLEMS[name]["data"])
    max_rel, bad, noise = util.compare_csr(dg.B, B, info.dim_elem)
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_inf(np.asarray(dg.L.todense()), np.asarray(L.todense())) <= util.TOL_ENTRY
