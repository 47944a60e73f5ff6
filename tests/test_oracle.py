"""The CPU oracle (oracle/dg_oracle.py) against (i) the reference's published constants, (ii) the golden fixtures written
by running the unmodified reference (oracle/make_golden.py) and (iii) the live reference when /root/reference exists."""
import numpy as np
import pytest

from oracle import dg_oracle, ref_harness
from reyna_b200.problems import PROBLEMS
from tests import util


def test_quadrature_known_answers():
    # SURVEY.md Appendix C (probed from the reference, p = 2)
    w, ref, w1, x1 = dg_oracle.quadrature(2)
    assert np.allclose(np.sort(x1), [-0.7745966692414834, 0.0, 0.7745966692414834], atol=1e-15)
    assert np.allclose(np.sort(w1), [0.5555555555555557, 0.5555555555555557, 0.8888888888888888], atol=1e-15)
    assert abs(w.sum() - 4.0) < 1e-14
    assert np.allclose(ref[:3], [[0.10271765480962625, 0.08858795951270393],
                                 [0.06655406783916451, 0.40946686444073477],
                                 [0.0239311322870806, 0.787659461760847]], atol=1e-15)
    assert np.allclose(w[:3], [0.4465153638643548, 0.5094246807990805, 0.15517106644767598], atol=1e-15)


def test_jacobi_known_answers():
    y = np.array([0.7])
    P = dg_oracle.jacobi_orthonormal(y, 0, 3)[:, 0]
    assert np.allclose(P, [0.7071067811865475, 0.8573214099741122, 0.3715676250697846, -0.36013452347699165], atol=1e-15)
    Q = dg_oracle.jacobi_orthonormal(y, 1, 2)[:, 0]
    assert np.allclose(Q, [0.8660254037844387, 1.355544171172596, 1.1746342515864248], atol=1e-15)


def test_basis_known_answers():
    orders = dg_oracle.basis_orders(2)
    assert orders.tolist() == [[0, 0], [0, 1], [0, 2], [1, 0], [1, 1], [2, 0]]
    m, h = (0.32, 0.41), (0.05, 0.06)
    bbox = np.array([[m[0] - h[0], m[0] + h[0], m[1] - h[1], m[1] + h[1]]])
    phi, grad = dg_oracle.basis_and_grad(np.array([[0.3, 0.4]]), bbox, orders)
    assert np.allclose(phi[:, 0], [9.128709291752767, -2.635231383473637, -9.355689989796867, -6.324555320336763,
                                   1.8257418583505467, -5.307227776030211], rtol=1e-12)
    assert np.allclose(grad[4, 0], [-91.28709291752726, -182.57418583505552], rtol=1e-11)
    assert np.allclose(grad[5, 0], [-489.89794855663604, 0.0], rtol=1e-11)


def test_basis_orders_negative():
    with pytest.raises(ValueError):
        dg_oracle.basis_orders(-1)


@pytest.mark.parametrize("mesh,key", util.golden_cases())
def test_oracle_matches_golden(mesh, key):
    z, geo = util.load_golden(mesh)
    name, p = util.split_case(key)
    pr = PROBLEMS[name]
    Bref, Lref, xref = util.golden_system(z, key)
    B, L, info = dg_oracle.assemble(geo, p, **pr["data"])
    d = info.dim_elem
    max_rel, bad, noise = util.compare_csr(B, Bref, d)
    assert bad == 0, f"{bad} non-noise pattern mismatches"
    assert max_rel <= util.TOL_ENTRY
    if noise == 0:
        assert np.array_equal(B.indptr, Bref.indptr) and np.array_equal(B.indices, Bref.indices)
        assert B.indices.dtype == np.int32
    assert util.rel_inf(np.asarray(L.todense()).reshape(-1), Lref) <= util.TOL_ENTRY
    x = dg_oracle.solve(B, L)
    assert util.rel_l2(x, xref) <= util.TOL_SOLUTION
    ell = z[key + "_elliptic"]
    inf = z[key + "_inflow"]
    if pr["data"].get("diffusion") is not None:
        assert np.array_equal(info.elliptical_indecies, ell)
    if pr["data"].get("advection") is not None:
        assert np.array_equal(info.inflow_indecies, inf)
    if key + "_errors" in z.files:
        eref = z[key + "_errors"]
        has_diff = pr["data"].get("diffusion") is not None
        e = dg_oracle.errors(geo, p, xref, pr["exact"], pr["grad_exact"] if has_diff else None,
                             advection=pr["data"].get("advection"), diffusion=pr["data"].get("diffusion"),
                             reaction=pr["data"].get("reaction"))
        assert abs(e[0] - eref[0]) <= 1e-9 * abs(eref[0])
        assert abs(e[1] - eref[1]) <= 1e-9 * abs(eref[1])
        if has_diff:
            assert abs(e[2] - eref[2]) <= 1e-9 * abs(eref[2])
        else:
            assert e[2] is None


@pytest.mark.reference
@pytest.mark.parametrize("name,p", [("adr", 2), ("variable", 3), ("variable_hyperbolic", 1)])
def test_oracle_matches_live_reference(name, p):
    ref = ref_harness.load()
    np.random.seed(99)
    dom = ref.domains.RectangleDomain(np.array([[0.0, 1.0], [0.0, 1.0]]))
    mesh = ref.poly_mesher(dom, max_iterations=4, n_points=150, cleaned=True, seed=5)
    geo = ref.DGFEMGeometry(mesh)
    pr = PROBLEMS[name]
    dg = ref.DGFEM(geo, polynomial_degree=p)
    dg.add_data(**pr["data"])
    dg.dgfem(solve=True)
    B, L, info = dg_oracle.assemble(geo, p, **pr["data"])
    max_rel, bad, noise = util.compare_csr(B, dg.B, info.dim_elem)
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_inf(np.asarray(L.todense()), np.asarray(dg.L.todense())) <= util.TOL_ENTRY
    assert util.rel_l2(dg_oracle.solve(B, L), dg.solution) <= util.TOL_SOLUTION
