"""The reference's convergence-rate benchmarks (tests/test_DGFEM.py:481-669) against the CUDA path: same PDE data, same mesh
sizes (32 ... 16384 elements), same thresholds on the last three observed rates.  Meshes come from the synthetic Voronoi
generator (the reference's poly_mesher is absent on the GPU box)."""
import numpy as np
import pytest

from reyna_b200.problems import PROBLEMS

pytestmark = pytest.mark.gpu

N_ELEMENTS = [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]


@pytest.fixture(scope="module")
def rb():
    import torch
    assert torch.cuda.is_available()
    import reyna_b200
    reyna_b200.lib()
    return reyna_b200


def _rates(rb, data, p, exact, grad_exact):
    h, l2s, dgs, h1s = [], [], [], []
    for i, n in enumerate(N_ELEMENTS):
        geo = rb.DGFEMGeometry(rb.voronoi_mesh(n, seed=1142 + i, lloyd_iterations=6), subtriangulation="fan")
        dg = rb.DGFEM(geo, polynomial_degree=p)
        dg.add_data(**data)
        dg.dgfem(solve=True)
        l2, dgn, h1 = dg.errors(exact_solution=exact, grad_exact_solution=grad_exact)
        h.append(geo.h)
        l2s.append(l2)
        dgs.append(dgn)
        h1s.append(h1)
    lh = np.diff(np.log(h))

    def rate(v):
        pw = np.diff(np.log(v)) / lh
        return float(np.exp(np.mean(np.log(pw[-3:]))))         # test_DGFEM.py:543: geometric mean of the last three rates

    return rate(l2s), rate(dgs), (rate(h1s) if h1s[0] is not None else None)


def test_benchmark_diffusion_p1(rb):
    # tests/test_DGFEM.py:481-545: a = I, c = pi^2, f = 3 pi^2 u, u = sin(pi x) sin(pi y); L2 > 1.6, dG > 0.7, H1 > 0.7
    pr = PROBLEMS["adr"]
    data = dict(diffusion=pr["data"]["diffusion"], reaction=pr["data"]["reaction"], dirichlet_bcs=pr["data"]["dirichlet_bcs"],
                forcing=lambda x: 3.0 * np.pi ** 2 * np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]))
    l2, dg, h1 = _rates(rb, data, 1, pr["exact"], pr["grad_exact"])
    assert l2 > 1.6 and dg > 0.7 and h1 > 0.7, (l2, dg, h1)


def test_benchmark_advection_reaction_p1(rb):
    # tests/test_DGFEM.py:547-596: b = (1,1), c = pi; L2 > 1.6, dG > 1.2
    pr = PROBLEMS["hyperbolic"]
    l2, dg, h1 = _rates(rb, pr["data"], 1, pr["exact"], None)
    assert h1 is None and l2 > 1.6 and dg > 1.2, (l2, dg)


def test_benchmark_advection_diffusion_reaction_p3(rb):
    # tests/test_DGFEM.py:598-669: a = [[2,1],[1,2]], b = (2,1), c = 2, u = exp(1-x) exp(y-1) / 2; L2 > 3.6, dG > 2.7, H1 > 2.7
    pr = PROBLEMS["adr_tensor"]
    l2, dg, h1 = _rates(rb, pr["data"], 3, pr["exact"], pr["grad_exact"])
    assert l2 > 3.6 and dg > 2.7 and h1 > 2.7, (l2, dg, h1)
