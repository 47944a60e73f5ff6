"""Repeated solves of one assembled system (A/B timing aid; the first solve carries allocator / graph warm-up):
python tools/solve_repeat.py E problem p [repeats] [solver-options-json]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import reyna_b200 as rb
from reyna_b200.problems import PROBLEMS
from bench import workload_geometry
E, name, p = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
geo, _, _ = workload_geometry(E)
dg = rb.DGFEM(geo, p)
dg.add_data(**PROBLEMS[name]['data'])
if len(sys.argv) > 5:
    dg.solver_options.update(json.loads(sys.argv[5]))
dg.dgfem(solve=False)
out = []
for _ in range(reps):
    t = time.perf_counter()
    dg._solve_device()
    out.append(dict(wall_s=round(time.perf_counter() - t, 4), solve_s=round(dg.solve_info['seconds'], 4),
                    setup_s=round(dg.solve_info['setup_seconds'], 4), iterations=dg.solve_info['iterations'],
                    rel_residual=dg.solve_info['rel_residual']))
print(json.dumps(dict(lib=os.environ.get('REYNA_B200_LIB', 'default'), elements=E, problem=name, p=p, runs=out)))
