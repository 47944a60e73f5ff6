"""Where the host side of dgfem(solve=False) spends its time (no GPU work): python tools/host_eval_breakdown.py [E]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import reyna_b200 as rb
from reyna_b200 import sampling
from reyna_b200.problems import PROBLEMS
from bench import workload_geometry
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
geo, _, _ = workload_geometry(E)
dg = rb.DGFEM(geo, 2); dg.add_data(**PROBLEMS['adr']['data'])
fg = rb.flat_geometry(geo)
pts = sampling.cached_points(geo, fg, dg.element_reference_quadrature, dg.edge_reference_quadrature)
out = {'threads': sampling.host_threads()}
for key, fn, X, tail in (('vol_a', dg.diffusion, pts['Xv'], (2, 2)), ('vol_b', dg.advection, pts['Xv'], (2,)), ('vol_c', dg.reaction, pts['Xv'], ()),
                         ('vol_f', dg.forcing, pts['Xv'], ()), ('if_a', dg.diffusion, pts['Xi'], (2, 2)), ('if_b', dg.advection, pts['Xi'], (2,)),
                         ('if_a_mid', dg.diffusion, pts['mid_i'], (2, 2))):
    buf = np.empty((X.shape[0],) + tail)
    sampling.evaluate_batched(fn, X, tail, out=buf)
    t = time.perf_counter()
    for _ in range(2): sampling.evaluate_batched(fn, X, tail, out=buf)
    dt = (time.perf_counter() - t) / 2
    out[key] = {'ms': round(dt * 1e3, 1), 'points_M': round(X.shape[0] / 1e6, 2), 'GB/s_out': round(buf.nbytes / dt / 1e9, 1)}
a = np.random.rand(64_000_000); b = np.empty_like(a); b[...] = a
t = time.perf_counter(); b[...] = a; out['memcpy_1thread_GB/s'] = round(2 * a.nbytes / (time.perf_counter() - t) / 1e9, 1)
print(json.dumps(out))
