"""Krylov solver scaling on synthetic Voronoi meshes (development aid, GPU box): iterations, time, residual."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import reyna_b200 as rb
from reyna_b200.problems import PROBLEMS
from bench import workload_geometry

def main():
    sizes = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else [16384, 65536, 262144]
    probs = sys.argv[2].split(',') if len(sys.argv) > 2 else ['adr', 'diffusion', 'hyperbolic']
    p = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    opts = json.loads(sys.argv[4]) if len(sys.argv) > 4 else {}
    for E in sizes:
        geo, _, _ = workload_geometry(E)
        for name in probs:
            dg = rb.DGFEM(geo, p)
            dg.add_data(**PROBLEMS[name]['data'])
            dg.solver_options.update(opts)
            dg.dgfem(solve=False)
            t0 = time.perf_counter()
            dg.solution = dg._solve_device()
            dt = time.perf_counter() - t0
            pr = PROBLEMS[name]
            has_diff = pr['data'].get('diffusion') is not None
            err = dg.errors(pr['exact'], pr['grad_exact'] if has_diff else None) if pr['exact'] is not None else None
            print(json.dumps(dict(E=E, problem=name, p=p, dof=dg.dim_system, solve_s=round(dt, 3), info=dg.solve_info, errors=err)), flush=True)

if __name__ == '__main__':
    main()
