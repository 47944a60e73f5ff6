"""Host evaluation of the benchmark callables: chunk size and malloc behaviour (numba returns a fresh array per call; above
glibc's mmap threshold every call page-faults its result).  python tools/host_eval_tuning.py"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reyna_b200 as rb
from reyna_b200 import sampling
from reyna_b200.problems import PROBLEMS

geo = rb.DGFEMGeometry(rb.voronoi_mesh(64, seed=1))
dg = rb.DGFEM(geo, 2)
dg.add_data(**PROBLEMS['adr']['data'])
M = 16_000_000
X = np.random.rand(M, 2)
terms = [('a', dg.diffusion, (2, 2)), ('b', dg.advection, (2,)), ('c', dg.reaction, ()), ('f', dg.forcing, ())]
bufs = {k: np.empty((M,) + t) for k, _, t in terms}


def run():
    t = time.perf_counter()
    sampling.evaluate_many([(fn, tail, bufs[k]) for k, fn, tail in terms], X)
    return time.perf_counter() - t


def per_term():
    out = {}
    for k, fn, tail in terms:
        t = time.perf_counter()
        sampling.evaluate_batched(fn, X, tail, out=bufs[k])
        out[k] = round((time.perf_counter() - t) * 1e3, 1)
    return out


print('threads', sampling.host_threads())
for chunk in (1 << 17, 1 << 15, 1 << 13):
    sampling._CHUNK = chunk
    run()
    print('chunk', chunk, 'best %.3f s' % min(run() for _ in range(3)), per_term())
libc = ctypes.CDLL('libc.so.6')
print('mallopt', libc.mallopt(-3, 64 * 1024 * 1024), libc.mallopt(-1, 1 << 30), libc.mallopt(-2, 0))
for chunk in (1 << 17, 1 << 15):
    sampling._CHUNK = chunk
    run()
    print('mallopt chunk', chunk, 'best %.3f s' % min(run() for _ in range(3)), per_term())
