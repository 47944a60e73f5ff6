"""Time-to-solution of dgfem(solve=True) over multigrid parameters (one process): python tools/mg_param_sweep.py E problem p"""
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
combos = [('agglomerated', 2, 0.8, 2.2), ('agglomerated', 3, 0.8, 2.2), ('agglomerated', 1, 0.8, 2.2), ('agglomerated', 2, 0.9, 2.4),
          ('agglomerated', 2, 0.7, 2.0), ('agglomerated', 2, 0.8, 1.8), ('agglomerated', 2, 0.8, 2.6), ('aggregation', 3, 0.7, 1.8)]
for hier, nu, om, al in combos:
    dg.solver_options.update(mg_hierarchy=hier, mg_nu=nu, mg_omega=om, mg_alpha=al)
    best = None
    for rep in range(2):
        t = time.perf_counter()
        dg.dgfem(solve=True)
        wall = time.perf_counter() - t
        si = dg.solve_info
        if best is None or si['seconds'] < best['solve_s']:
            best = {'hierarchy': hier, 'nu': nu, 'omega': om, 'alpha': al, 'iterations': si['iterations'],
                    'rel_residual': si['rel_residual'], 'solve_s': si['seconds'], 'setup_s': si['setup_seconds'], 'wall_s': wall,
                    'levels': si['mg_levels']}
    print(json.dumps(best), flush=True)
