"""One dgfem(solve=True) on a cached workload (profiling aid): python tools/solve_once.py E problem p"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import reyna_b200 as rb
from reyna_b200.problems import PROBLEMS
from bench import workload_geometry
E, name, p = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
geo, _, _ = workload_geometry(E)
dg = rb.DGFEM(geo, p)
dg.add_data(**PROBLEMS[name]['data'])
if len(sys.argv) > 4:
    dg.solver_options.update(json.loads(sys.argv[4]))
dg.dgfem(solve=True)
print(json.dumps(dg.solve_info))
