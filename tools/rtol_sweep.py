"""Iterations / true residual / L2 error of dgfem(solve=True) as a function of the requested rtol (one process):
python tools/rtol_sweep.py E problem p"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import reyna_b200 as rb
from reyna_b200.problems import PROBLEMS
from bench import workload_geometry
E, name, p = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
geo, _, _ = workload_geometry(E)
pr = PROBLEMS[name]
dg = rb.DGFEM(geo, p)
dg.add_data(**pr['data'])
dg.dgfem(solve=False)
for rtol in [1e-6, 1e-8, 1e-9, 1e-10, 1e-11, 1e-12, 1e-14]:
    dg.solver_options['rtol'] = rtol
    t = time.perf_counter()
    dg.dgfem(solve=True)
    wall = time.perf_counter() - t
    si = dg.solve_info
    has_diff = pr['data'].get('diffusion') is not None
    l2, dgn, h1 = dg.errors(pr['exact'], pr['grad_exact'] if has_diff else None)
    print(json.dumps({'rtol': rtol, 'iterations': si['iterations'], 'rel_residual': si['rel_residual'], 'converged': si['converged'],
                      'stagnated': si['stagnated'], 'solve_s': si['seconds'], 'wall_s': wall, 'l2': l2, 'dg': dgn}), flush=True)
