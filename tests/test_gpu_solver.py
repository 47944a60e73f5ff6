"""The Krylov solver that replaces spsolve (main.py:231): multigrid hierarchy and cycle against a scipy restatement of the
same algorithm, iteration counts, and the solution against SuperLU."""
import ctypes as C

import numpy as np
import pytest

from reyna_b200.problems import PROBLEMS
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import torch
    assert torch.cuda.is_available()
    import reyna_b200
    reyna_b200.lib()
    return reyna_b200


@pytest.mark.parametrize("E,name,p", [(6000, "diffusion", 2), (6000, "adr", 1), (5000, "variable", 3)])
def test_multigrid_operator_matches_scipy_restatement(rb, E, name, p):
    import torch
    from oracle.mg_reference import CoarseCorrection
    from reyna_b200 import _lib
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(E, seed=7), subtriangulation="fan")
    dg = rb.DGFEM(geo, polynomial_degree=p)
    dg.add_data(**PROBLEMS[name]["data"])
    dg.dgfem(solve=False)
    dv = dg._dev
    assert dv["blocked"]
    ref = CoarseCorrection(dg.B, p)
    L = rb.lib()
    ip, ix, dat = dv["structural"]
    mg = C.c_void_p(None)
    sp_ = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mo = _lib.RbMgOpts(nu=3, omega=0.7, alpha=1.8)
    _lib.check(L.reyna_b200_mg_create(dv["n"], dv["d"], p, ip.data_ptr(), ix.data_ptr(), dat.data_ptr(), C.byref(mo),
                                      None, C.byref(mg), sp_), "mg_create")
    try:
        nl = C.c_int32(0)
        rows = (C.c_int32 * 16)()
        nnz = (C.c_longlong * 16)()
        L.reyna_b200_mg_info(mg, C.byref(nl), rows, nnz)
        assert [rows[i] for i in range(nl.value)] == ref.level_sizes()
        assert [nnz[i] for i in range(nl.value)] == [lv["A"].nnz for lv in ref.levels]
        rng = np.random.default_rng(0)
        for _ in range(2):
            r = rng.standard_normal(dv["n"])
            rt = torch.from_numpy(r).cuda()
            zt = torch.empty_like(rt)
            _lib.check(L.reyna_b200_mg_apply(mg, rt.data_ptr(), zt.data_ptr(), sp_), "mg_apply")
            z = zt.cpu().numpy()
            zr = ref.apply(r)
            assert util.rel_l2(z, zr) <= 1e-10
    finally:
        L.reyna_b200_mg_destroy(mg)


@pytest.mark.parametrize("E,name,p", [(6000, "diffusion", 2), (6000, "adr", 1), (5000, "variable", 3), (60, "diffusion", 1)])
def test_agglomerated_multigrid_matches_scipy_restatement(rb, E, name, p):
    """The agglomerated-P1 chain (rb_mg_opts.elem_bbox): level sizes, operator sizes and the cycle against scipy."""
    import torch
    from oracle.mg_reference import AgglomeratedP1
    from reyna_b200 import _lib
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(E, seed=11), subtriangulation="fan")
    dg = rb.DGFEM(geo, polynomial_degree=p)
    dg.add_data(**PROBLEMS[name]["data"])
    dg.dgfem(solve=False)
    dv = dg._dev
    assert dv["blocked"]
    bbox = dv["bbox"].cpu().numpy()
    ref = AgglomeratedP1(dg.B, p, bbox, nu=2, omega=0.8, alpha=2.2)
    L = rb.lib()
    ip, ix, dat = dv["structural"]
    mg = C.c_void_p(None)
    sp_ = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mo = _lib.RbMgOpts(nu=2, omega=0.8, alpha=2.2, elem_bbox=dv["bbox"].data_ptr())
    _lib.check(L.reyna_b200_mg_create(dv["n"], dv["d"], p, ip.data_ptr(), ix.data_ptr(), dat.data_ptr(), C.byref(mo),
                                      None, C.byref(mg), sp_), "mg_create")
    try:
        nl = C.c_int32(0)
        rows = (C.c_int32 * 16)()
        nnz = (C.c_longlong * 16)()
        L.reyna_b200_mg_info(mg, C.byref(nl), rows, nnz)
        assert [rows[i] for i in range(nl.value)] == ref.level_sizes()
        # the CUDA levels store complete 3 x 3 blocks (structural zeros included); scipy prunes nothing either
        assert [nnz[i] for i in range(1, nl.value)] == [lv["A"].nnz for lv in ref.levels[1:]]
        rng = np.random.default_rng(0)
        for _ in range(2):
            r = rng.standard_normal(dv["n"])
            rt = torch.from_numpy(r).cuda()
            zt = torch.empty_like(rt)
            _lib.check(L.reyna_b200_mg_apply(mg, rt.data_ptr(), zt.data_ptr(), sp_), "mg_apply")
            z = zt.cpu().numpy()
            zr = ref.apply(r)
            assert util.rel_l2(z, zr) <= 1e-10
    finally:
        L.reyna_b200_mg_destroy(mg)


def test_agglomerated_hierarchy_needs_fewer_iterations(rb):
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(16384, tiles=4, seed=3), subtriangulation="fan")
    its = {}
    for hierarchy in ("aggregation", "agglomerated"):
        dg = rb.DGFEM(geo, polynomial_degree=2)
        dg.add_data(**PROBLEMS["diffusion"]["data"])
        dg.solver_options.update(rtol=1e-8, mg_hierarchy=hierarchy)
        dg.dgfem(solve=True)
        assert dg.solve_info["converged"], dg.solve_info
        its[hierarchy] = dg.solve_info["iterations"]
    assert its["agglomerated"] <= 0.7 * its["aggregation"], its


@pytest.mark.parametrize("name,p,limit", [("diffusion", 2, 450), ("adr", 2, 300), ("diffusion", 1, 450), ("adr_tensor", 3, 450)])
def test_multigrid_iterations_are_mesh_independent(rb, name, p, limit):
    its = []
    for E in (4096, 16384):
        geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(E, tiles=4, seed=3), subtriangulation="fan")
        dg = rb.DGFEM(geo, polynomial_degree=p)
        dg.add_data(**PROBLEMS[name]["data"])
        dg.solver_options.update(rtol=1e-8)       # well above the kappa * eps floor of the true residual
        dg.dgfem(solve=True)
        assert dg.solve_info["preconditioner"] == "mg"
        assert dg.solve_info["converged"], dg.solve_info
        its.append(dg.solve_info["iterations"])
    assert max(its) <= limit, its
    assert its[1] <= 1.5 * its[0] + 20, its           # block Jacobi alone doubles here


def test_multigrid_solution_matches_spsolve(rb):
    from oracle import dg_oracle
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(16384, tiles=4, seed=5), subtriangulation="fan")
    for name, p in (("adr", 2), ("diffusion", 2)):
        dg = rb.DGFEM(geo, polynomial_degree=p)
        dg.add_data(**PROBLEMS[name]["data"])
        dg.dgfem(solve=True)
        B, L, info = dg_oracle.assemble(geo, p, **PROBLEMS[name]["data"])
        x = dg_oracle.solve(B, L)
        assert util.rel_l2(dg.solution, x) <= util.TOL_SOLUTION, dg.solve_info


@pytest.mark.parametrize("name,p", [("adr", 2), ("diffusion", 1)])
def test_single_precision_level_matrices_only_change_the_preconditioner(rb, name, p):
    """rb_mg_opts.fp32_levels: the smoothing sweeps read float copies of the level matrices; the Krylov iteration, its
    residuals and the stopping rule stay in double precision, so the solution agrees to the solve tolerance and the iteration
    count moves by a few at most."""
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(16384, tiles=4, seed=7), subtriangulation="fan")
    out = {}
    for fp32 in (False, True):
        dg = rb.DGFEM(geo, polynomial_degree=p)
        dg.add_data(**PROBLEMS[name]["data"])
        dg.solver_options.update(rtol=1e-9, mg_fp32=fp32)
        dg.dgfem(solve=True)
        assert dg.solve_info["converged"], dg.solve_info
        out[fp32] = (dg.solution.copy(), dg.solve_info["iterations"])
    assert util.rel_l2(out[True][0], out[False][0]) <= 1e-7
    assert out[True][1] <= out[False][1] * 1.15 + 4, (out[True][1], out[False][1])


def test_block_jacobi_still_available(rb):
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(2000, seed=2), subtriangulation="fan")
    dg = rb.DGFEM(geo, polynomial_degree=2)
    dg.add_data(**PROBLEMS["diffusion"]["data"])
    dg.solver_options.update(preconditioner="bj", rtol=1e-8)
    dg.dgfem(solve=True)
    assert dg.solve_info["preconditioner"] == "bj" and dg.solve_info["converged"]
    bj_its = dg.solve_info["iterations"]
    dg.solver_options.update(preconditioner="mg")
    dg.dgfem(solve=True)
    assert dg.solve_info["iterations"] < bj_its


def test_solution_matches_spsolve_at_config3a(rb):
    """BASELINE config 3a (65 536 elements, p=2 ADR): the multigrid-Krylov solution against SuperLU on the oracle's matrix."""
    import time
    from oracle import dg_oracle
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(65536, tiles=4, seed=1337), subtriangulation="fan")
    dg = rb.DGFEM(geo, polynomial_degree=2)
    dg.add_data(**PROBLEMS["adr"]["data"])
    dg.dgfem(solve=True)
    B, L, info = dg_oracle.assemble(geo, 2, **PROBLEMS["adr"]["data"])
    t0 = time.time()
    x = dg_oracle.solve(B, L)
    print(f"spsolve: {time.time() - t0:.1f} s; ours: {dg.solve_info}")
    assert util.rel_l2(dg.solution, x) <= util.TOL_SOLUTION, dg.solve_info
    pr = PROBLEMS["adr"]
    e = dg.errors(pr["exact"], pr["grad_exact"])
    eo = dg_oracle.errors(geo, 2, x, pr["exact"], pr["grad_exact"], advection=pr["data"]["advection"],
                          diffusion=pr["data"]["diffusion"], reaction=pr["data"]["reaction"])
    for a, b in zip(e, eo):
        assert abs(a - b) <= util.TOL_ERRORS * abs(b)


def test_default_solve_converges_at_the_rounding_floor_and_warns_otherwise(rb):
    """Default options run to the rounding floor of the true residual and report it as converged; a solve that is cut short
    above the acceptance tolerance warns instead of returning silently (the reference's spsolve is exact)."""
    import warnings
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(4096, tiles=4, seed=9), subtriangulation="fan")
    dg = rb.DGFEM(geo, polynomial_degree=2)
    dg.add_data(**PROBLEMS["adr"]["data"])
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        dg.dgfem(solve=True)
    si = dg.solve_info
    assert si["converged"] and si["rel_residual"] <= 1e-7 and si["singular_blocks"] == 0, si
    dg.solver_options.update(max_iter=5)
    with pytest.warns(UserWarning, match="may be inaccurate"):
        dg.dgfem(solve=True)
    assert not dg.solve_info["converged"]


def test_solver_is_independent_of_the_element_numbering(rb):
    """The reference's mesher numbers elements in random seed order (polymesher/two_dimensional/main.py:111-165).  A shuffled
    65k-element mesh is solved on its Morton renumbering: about the same iteration count as the ordered mesh, the solution
    (in the caller's numbering) equal to the ordered one, and B / L in the caller's numbering too."""
    from reyna_b200.partition import RenumberedGeometry, locality_order
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(65536, tiles=4, seed=3), subtriangulation="fan")
    assert locality_order(geo) is None
    ref = rb.DGFEM(geo, polynomial_degree=2)
    ref.add_data(**PROBLEMS["adr"]["data"])
    ref.dgfem(solve=True)
    shuffle = np.random.default_rng(7).permutation(geo.n_elements)
    sgeo = RenumberedGeometry(geo, shuffle)                       # new element i = old element shuffle[i]
    assert locality_order(sgeo) is not None
    dg = rb.DGFEM(sgeo, polynomial_degree=2)
    dg.add_data(**PROBLEMS["adr"]["data"])
    dg.dgfem(solve=True)
    assert dg._perm is not None
    assert dg.solve_info["converged"] and dg.solve_info["iterations"] <= 1.3 * ref.solve_info["iterations"] + 10, \
        (dg.solve_info, ref.solve_info)
    d = dg.dim_elem
    mine = dg.solution.reshape(-1, d)                             # shuffled numbering
    theirs = ref.solution.reshape(-1, d)[shuffle]
    assert util.rel_l2(mine.reshape(-1), theirs.reshape(-1)) <= util.TOL_SOLUTION
    # the matrix comes back in the caller's (shuffled) numbering: compare with a plain assembly in that numbering
    asm = rb.DGFEM(sgeo, polynomial_degree=2)
    asm.add_data(**PROBLEMS["adr"]["data"])
    asm.dgfem(solve=False)
    assert asm._perm is None
    max_rel, bad, noise = util.compare_csr(dg.B, asm.B, d)
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_inf(np.asarray(dg.L.todense()), np.asarray(asm.L.todense())) <= util.TOL_ENTRY
    # without the renumbering the hierarchy aggregates unrelated elements
    dg.renumber = False
    dg.solver_options.update(max_iter=3 * ref.solve_info["iterations"])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dg.dgfem(solve=True)
    assert not dg.solve_info["converged"] or dg.solve_info["iterations"] > 1.5 * ref.solve_info["iterations"]


def _rotating(x):
    out = np.empty(x.shape, dtype=np.float64)
    out[:, 0] = -(x[:, 1] - 0.5)
    out[:, 1] = x[:, 0] - 0.5
    return out


def test_hyperbolic_flow_sweep_is_a_direct_solve(rb):
    """Advection-reaction without diffusion: in the topological order of the upwind graph B is block lower triangular, the
    flow-ordered block Gauss-Seidel sweep solves it; block Jacobi needed hundreds of iterations."""
    from oracle import dg_oracle
    name = "hyperbolic"
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(16384, tiles=4, seed=3), subtriangulation="fan")
    dg = rb.DGFEM(geo, polynomial_degree=2)
    dg.add_data(**PROBLEMS[name]["data"])
    dg.dgfem(solve=True)
    si = dg.solve_info
    assert si["preconditioner"] == "sweep" and si["converged"] and si["flow_cycles_broken"] == 0, si
    assert si["iterations"] <= 5, si
    B, L, info = dg_oracle.assemble(geo, 2, **PROBLEMS[name]["data"])
    xo = dg_oracle.solve(B, L)
    assert util.rel_l2(dg.solution, xo) <= util.TOL_SOLUTION, si
    # the same system with plain block Jacobi: same solution, many more iterations
    dg.solver_options.update(preconditioner="bj")
    dg.dgfem(solve=True)
    assert dg.solve_info["preconditioner"] == "bj"
    assert util.rel_l2(dg.solution, xo) <= util.TOL_SOLUTION
    assert dg.solve_info["iterations"] > 100


def _rotating(x):
    out = np.empty(x.shape, dtype=np.float64)
    out[:, 0] = -(x[:, 1] - 0.5)
    out[:, 1] = x[:, 0] - 0.5
    return out


@pytest.mark.parametrize("E", [400, 4000])
def test_flow_sweep_with_closed_streamlines_still_converges(rb, E):
    """A rotating advection field: the upwind graph is all cycles.  They are broken at the elements with the fewest unresolved
    upwind neighbours; the sweep is a preconditioner then (small mesh) or is dropped for block Jacobi when the resulting order
    is nearly sequential (large mesh) -- either way the solve converges to the oracle's solution."""
    from oracle import dg_oracle
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(E, seed=4), subtriangulation="fan")
    data = dict(advection=_rotating, reaction=lambda x: np.full(x.shape[0], 2.0), forcing=lambda x: np.ones(x.shape[0]),
                dirichlet_bcs=lambda x: np.zeros(x.shape[0]))
    dg = rb.DGFEM(geo, polynomial_degree=1)
    dg.add_data(**data)
    dg.dgfem(solve=True)
    assert dg.solve_info["converged"], dg.solve_info
    assert dg.solve_info["preconditioner"] in ("sweep", "bj")
    B, L, info = dg_oracle.assemble(geo, 1, **data)
    assert util.rel_l2(dg.solution, dg_oracle.solve(B, L)) <= util.TOL_SOLUTION, dg.solve_info
