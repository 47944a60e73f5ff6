"""Verbose parity report of the CUDA path against the golden fixtures and the oracle (development aid, GPU box).

    python tools/gpu_debug.py [> gpurun_out/debug.log]
Prints, per case, the block-relative error of B split into diagonal / off-diagonal blocks, the worst entry, pattern
mismatches, load and solution errors.  Never stops at the first failure.
"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import reyna_b200 as rb                      # noqa: E402
from reyna_b200.problems import PROBLEMS     # noqa: E402
from oracle import dg_oracle                 # noqa: E402
from tests import util                       # noqa: E402


def report(tag, dg, Bref, Lref, xref, d):
    B = dg.B
    L = np.asarray(dg.L.todense()).reshape(-1)
    max_rel, bad, noise = util.compare_csr(B, Bref, d)
    D = (B - Bref).tocoo()
    line = f"{tag}: nnz {B.nnz} vs {Bref.nnz}  B rel {max_rel:.2e}  pattern bad {bad} noise {noise}"
    if D.nnz:
        k = np.argmax(np.abs(D.data))
        r, c = int(D.row[k]), int(D.col[k])
        line += f"  worst abs {abs(D.data[k]):.3e} at row ({r // d},{r % d}) col ({c // d},{c % d}) ours {B[r, c]:.6e} ref {Bref[r, c]:.6e}"
        diag = (D.row // d) == (D.col // d)
        if diag.any():
            line += f"  |dDiag| {np.abs(D.data[diag]).max():.2e}"
        if (~diag).any():
            line += f"  |dOff| {np.abs(D.data[~diag]).max():.2e}"
    line += f"  L rel {util.rel_inf(L, Lref):.2e}"
    if dg.solution is not None and xref is not None:
        line += f"  x rel {util.rel_l2(dg.solution, xref):.2e}  solve {dg.solve_info}"
    print(line, flush=True)


def main():
    import torch
    print("torch", torch.__version__, "cuda", torch.cuda.is_available(), torch.cuda.get_device_name(0))
    print("lib", rb.lib().reyna_b200_version().decode())
    for mesh, key in util.golden_cases():
        try:
            z, geo = util.load_golden(mesh)
            name, p = util.split_case(key)
            Bref, Lref, xref = util.golden_system(z, key)
            dg = rb.DGFEM(geo, polynomial_degree=p)
            dg.add_data(**PROBLEMS[name]["data"])
            dg.dgfem(solve=True)
            report(f"golden {mesh}/{key}", dg, Bref, Lref, xref, dg.dim_elem)
        except Exception:
            print(f"golden {mesh}/{key}: EXCEPTION\n{traceback.format_exc()}", flush=True)
    # larger synthetic meshes against the oracle
    for E, name, p in ((2048, "adr", 2), (2048, "variable", 3), (4096, "variable_hyperbolic", 2), (4096, "diffusion", 1)):
        try:
            t0 = time.time()
            mesh = rb.voronoi_mesh(E, seed=3)
            geo = rb.DGFEMGeometry(mesh, subtriangulation="fan")
            t1 = time.time()
            dg = rb.DGFEM(geo, polynomial_degree=p)
            dg.add_data(**PROBLEMS[name]["data"])
            dg.dgfem(solve=True)
            t2 = time.time()
            B, L, info = dg_oracle.assemble(geo, p, **PROBLEMS[name]["data"])
            x = dg_oracle.solve(B, L)
            print(f"mesh+geom {t1 - t0:.2f}s dgfem {t2 - t1:.2f}s oracle {time.time() - t2:.2f}s timings {dg.timings}")
            report(f"oracle E={E} {name} p={p}", dg, B, np.asarray(L.todense()).reshape(-1), x, dg.dim_elem)
        except Exception:
            print(f"oracle E={E} {name} p={p}: EXCEPTION\n{traceback.format_exc()}", flush=True)


if __name__ == "__main__":
    main()
