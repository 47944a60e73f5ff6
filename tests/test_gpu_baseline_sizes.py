"""Entry-by-entry parity at the sizes BASELINE.json names (configs 3b, 4, 5).  Configs 4 and 5 are too large for one oracle
call: the oracle assembles ghost-inclusive sub-meshes cut with ``partition_geometry`` (tests/util.oracle_rows) -- eight
slices spread over the mesh including the last one, where 32-bit offsets would break first."""
import numpy as np
import pytest

from reyna_b200.problems import PROBLEMS
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import reyna_b200
    reyna_b200.lib()
    return reyna_b200


def _assemble(rb, geo, name, p):
    dg = rb.DGFEM(geo, polynomial_degree=p)
    dg.add_data(**PROBLEMS[name]["data"])
    dg.dgfem(solve=False)
    return dg


def _check_slices(rb, E, name, p, width, n_slices=8):
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(E, tiles=4, seed=1337), subtriangulation="fan")
    dg = _assemble(rb, geo, name, p)
    B, Lg = dg.B, np.asarray(dg.L.todense()).reshape(-1)
    d = dg.dim_elem
    assert B.indices.dtype == np.int32 and B.has_canonical_format
    starts = [int(round(k * (E - width) / (n_slices - 1))) // 4 * 4 for k in range(n_slices)]
    starts[-1] = E - width                                    # the last block rows of the matrix
    worst, total_noise = 0.0, 0
    for lo in starts:
        Bo, Lo, d_o = util.oracle_rows(geo, lo, lo + width, name, p, dg.dim_system)
        assert d_o == d
        max_rel, bad, noise = util.compare_csr(B[lo * d:(lo + width) * d], Bo, d)
        assert bad == 0 and max_rel <= util.TOL_ENTRY, (lo, max_rel, bad, noise)
        assert util.rel_inf(Lg[lo * d:(lo + width) * d], Lo) <= util.TOL_ENTRY
        worst, total_noise = max(worst, max_rel), total_noise + noise
    util.note(f"[parity] {E} elements {name} p={p}: {n_slices} slices of {width} elements, max block-relative error {worst:.2e}, "
          f"noise-level pattern entries {total_noise}")


def test_config_3b_65k_p3_adr_entry_by_entry(rb):
    """BASELINE config 3 (second half): 65 536 elements, p = 3 ADR, the whole matrix against one oracle call."""
    from oracle import dg_oracle
    geo = rb.DGFEMGeometry(rb.tiled_voronoi_mesh(65536, tiles=4, seed=1337), subtriangulation="fan")
    dg = _assemble(rb, geo, "adr", 3)
    B, L, info = dg_oracle.assemble(geo, 3, **PROBLEMS["adr"]["data"])
    max_rel, bad, noise = util.compare_csr(dg.B, B, info.dim_elem)
    util.note(f"[parity] 65536 elements adr p=3: max block-relative error {max_rel:.2e}, noise-level pattern entries {noise}")
    assert bad == 0 and max_rel <= util.TOL_ENTRY
    assert util.rel_inf(np.asarray(dg.L.todense()), np.asarray(L.todense())) <= util.TOL_ENTRY


def test_config_4_262k_p2_hyperbolic_slices(rb):
    """BASELINE config 4: 262 144 elements, p = 2, advection-reaction (one-sided upwind pattern, inflow faces)."""
    _check_slices(rb, 262144, "hyperbolic", 2, width=8192)


def test_config_5_1m_p2_adr_slices(rb):
    """BASELINE config 5: 1 048 576 elements, p = 2 ADR -- 263.5 M non-zeros, the size where 32-bit offsets overflow first."""
    _check_slices(rb, 1048576, "adr", 2, width=4096)
