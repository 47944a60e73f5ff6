"""Per-call GPU time of one assembly step (CUDA events, synchronised between calls; development aid)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import reyna_b200 as rb
from reyna_b200 import _lib
from reyna_b200.problems import PROBLEMS
from bench import workload_geometry

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
geo, _, _ = workload_geometry(E)
dg = rb.DGFEM(geo, p)
dg.add_data(**PROBLEMS['adr']['data'])
dg.keep_device_inputs = True
dg.dgfem(solve=False)
L = rb.lib()
orig = {}
times = {}

def wrap(name):
    f = getattr(L, name)
    def g(*a):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream()); 
        r = f(*a)
        torch.cuda.synchronize()
        e1.record(torch.cuda.current_stream()); torch.cuda.synchronize()
        times.setdefault(name, []).append(e0.elapsed_time(e1))
        return r
    g.restype = f.restype
    return g

class Proxy:
    def __getattr__(self, n):
        if n in ('reyna_b200_structure', 'reyna_b200_csr_pattern', 'reyna_b200_assemble', 'reyna_b200_boundary'):
            return wrap(n)
        return getattr(L, n)

fg, gd, cd, bd = dg._inputs
dev = torch.device('cuda', 0)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(6):
    dg._assemble_on_device(torch, Proxy(), dev, sp, fg, gd, cd, bd, {})
for k, v in times.items():
    print(k, 'ms (wall incl. launch, serialised):', [round(x, 3) for x in v[2:]])
