"""All five BASELINE.json configurations on one B200: assembly (device-only and end-to-end) and dgfem(solve=True)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ctypes as C
import reyna_b200 as rb
from reyna_b200.problems import PROBLEMS
from bench import workload_geometry

CONFIGS = [("1 README", 1024, 1, "adr"), ("2 pure diffusion", 16384, 1, "diffusion"), ("3a ADR", 65536, 2, "adr"),
           ("3b ADR", 65536, 3, "adr"), ("4 hyperbolic", 262144, 2, "hyperbolic"), ("5 ADR", 1048576, 2, "adr")]
dev = torch.device("cuda", 0)
L = rb.lib()
for name, E, p, prob in CONFIGS:
    geo, _, _ = workload_geometry(E)
    dg = rb.DGFEM(geo, p)
    dg.add_data(**PROBLEMS[prob]["data"])
    dg.keep_device_inputs = True
    for _ in range(2):
        dg.dgfem(solve=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        dg.dgfem(solve=False)
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / 3
    fg, gd, cd, bd = dg._inputs
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        dv = dg._assemble_on_device(torch, L, dev, sp, fg, gd, cd, bd, {})
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10):
        dv = dg._assemble_on_device(torch, L, dev, sp, fg, gd, cd, bd, {})
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    nnz = int(dv["data"].numel())
    dg._inputs = None; dg.keep_device_inputs = False; del dv
    dg.dgfem(solve=True)                       # warm-up: solver workspaces, CUDA graph of the preconditioner, allocator pools
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dg.dgfem(solve=True)
    wall = time.perf_counter() - t0
    pr = PROBLEMS[prob]
    has_diff = pr["data"].get("diffusion") is not None
    err = dg.errors(pr["exact"], pr["grad_exact"] if has_diff else None)
    si = dg.solve_info
    print(json.dumps(dict(config=name, elements=E, p=p, problem=prob, dof=dg.dim_system, nnz=nnz, assemble_ms=round(ms, 3),
                          gdof_per_s=round(dg.dim_system / ms / 1e6, 3), e2e_assemble_s=round(e2e, 4),
                          dgfem_solve_true_s=round(wall, 3), krylov=si["method"], precond=si["preconditioner"],
                          iterations=si["iterations"], true_rel_residual=si["rel_residual"], errors=err)), flush=True)
    del dg
    torch.cuda.empty_cache()
