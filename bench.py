#!/usr/bin/env python
"""bench.py -- assembled DoF/s of the DGFEM hot path (BASELINE.json: p=2 ADR on a 1M-element Voronoi mesh).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--elements E] [--degree p]

A "step" is one complete assembly of the stiffness matrix B (CSR) and load vector L of the workload:
element->face adjacency + segmented sort, CSR pattern, fused volume + interior-face kernel, boundary-face kernel,
exact-zero pruning when needed.

  value   assembled DoF/s with every input (geometry, batched coefficient samples) already resident in HBM
  e2e     the same metric through the drop-in call ``DGFEM.dgfem(solve=False)`` from HOST data: batched evaluation of the
          user callables on the host, pinned upload, kernels, and the device->host read of B (indptr/indices/data) and L
  roofline  the fused assembly kernel alone: algorithmic bytes / CUDA-event time against the measured HBM peak
  cpu_baseline / --impl reference   the oracle port (numpy restatement of the reference) on the box's host cores

N > 1 (torchrun): strong scaling -- the SAME mesh, elements partitioned into N contiguous (Morton-ordered) chunks, every
rank assembling its own block rows; no collective on the assembly path.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CACHE_DIR = os.environ.get("REYNA_B200_CACHE", os.path.join(tempfile.gettempdir(), "reyna_b200_cache"))


# ----------------------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------------------
def workload_geometry(n_elements: int, seed: int = 1337, verbose: bool = False):
    """Synthetic unit-square Voronoi mesh + DGFEMGeometry, cached as one flat .npz ("generated once on host")."""
    import reyna_b200 as rb
    os.makedirs(CACHE_DIR, exist_ok=True)
    path = os.path.join(CACHE_DIR, f"voronoi_unit_square_E{n_elements}_seed{seed}.npz")
    if os.path.exists(path):
        return rb.load_geometry_cache(path), path, True
    t0 = time.time()
    tiles = 4 if n_elements >= 16384 and n_elements % 16 == 0 else 1
    if tiles > 1:
        mesh = rb.tiled_voronoi_mesh(n_elements, tiles=tiles, seed=seed, lloyd_iterations=3)
    else:
        mesh = rb.voronoi_mesh(n_elements, seed=seed, lloyd_iterations=3)
    geo = rb.DGFEMGeometry(mesh, subtriangulation="fan")
    rb.save_geometry_cache(geo, path)
    if verbose:
        print(f"[bench] generated {n_elements}-element mesh + geometry in {time.time() - t0:.1f}s -> {path}", file=sys.stderr)
    return geo, path, False


def workload_name(n_elements: int, p: int, problem: str) -> str:
    return (f"unit-square Voronoi mesh, {n_elements} elements (4x4 mirrored tiles of a Lloyd-relaxed bounded Voronoi mesh, "
            f"seed 1337, fan sub-triangulation), p={p} {problem} (a=I, b=(1,1), c=pi^2, README forcing, Dirichlet g=u)")


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def _oracle_worker(args):
    g, p, problem = args
    from oracle import dg_oracle
    from reyna_b200.problems import PROBLEMS
    co = dg_oracle.wrap_callables(**PROBLEMS[problem]["data"])
    # warm the numba-compiled callables (compile time is not part of the reference's assembly time either)
    x = np.zeros((4, 2))
    for f in (co.advection, co.diffusion, co.reaction, co.forcing, co.dirichlet_bcs):
        if f is not None:
            f(x)
    t0 = time.perf_counter()
    B, L, info = dg_oracle.assemble(g, p, precompiled=co, mirror_index_quirk=False)
    d = info.dim_elem
    rows = g.n_rows * d
    return time.perf_counter() - t0, int(rows), int(B[:rows].nnz)


def oracle_throughput(n_elements: int, sample_elements: int, p: int, problem: str, procs: int, repeats: int = 1):
    """Assembled DoF/s of the oracle port on REAL SLICES of the workload: `procs` concurrent processes, process k assembling
    the block rows of `sample_elements` consecutive elements of the `n_elements` mesh (slices spread over the mesh; the
    ghost-inclusive sub-mesh is cut with reyna_b200.partition.partition_geometry).  The reference is a single-threaded Python
    program; independent row slices are the only way it scales.  Returns (DoF per step, step times, nnz per step)."""
    import multiprocessing as mp
    from oracle import dg_oracle
    from reyna_b200.partition import partition_geometry
    geo, _, _ = workload_geometry(n_elements)
    sample_elements = min(sample_elements, n_elements)
    starts = [((k * (n_elements - sample_elements)) // max(1, procs - 1)) // 4 * 4 if procs > 1 else 0 for k in range(procs)]
    subs = [dg_oracle.flatten_partition(partition_geometry(geo, lo, lo + sample_elements)) for lo in starts]
    ctx = mp.get_context("spawn")
    times = []
    ndof = nnz = 0
    with ctx.Pool(procs) as pool:
        for _ in range(repeats):
            res = pool.map(_oracle_worker, [(g, p, problem) for g in subs])
            ndof = sum(r[1] for r in res)
            nnz = sum(r[2] for r in res)
            # per-step time = slowest worker (excludes process start-up / numba compilation, which dgfem() timing excludes too)
            times.append(max(r[0] for r in res))
    return ndof, times, nnz


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank: int) -> None:
    if rank != 0:
        return
    cores = max(1, min(host_cores(), 32))
    sample = args.reference_sample
    ndof, times, nnz = oracle_throughput(args.elements, sample, args.degree, args.problem, cores,
                                         repeats=args.warmup + args.steps)
    timed = times[args.warmup:]
    t = float(np.mean(timed))
    value = ndof / t
    d = (args.degree + 1) * (args.degree + 2) // 2
    sample_txt = (f"{cores} concurrent oracle processes (oracle/dg_oracle.py, numpy port of reyna's dgfem(solve=False)), each "
                  f"assembling the block rows of {sample} consecutive elements of the {args.elements}-element workload per step "
                  f"({ndof} DoF, {nnz} non-zeros per step; slices spread over the mesh, ghost layer included)")
    line = {
        "impl": "reference", "metric": "assembled DoF/s", "value": value, "unit": "DoF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.elements, args.degree, args.problem), "elements": args.elements,
                   "degree": args.degree, "problem": args.problem, "dof": args.elements * d,
                   "sampled_dof_per_step": ndof, "sampled_nnz_per_step": nnz},
        "cpu_baseline": {"value": value, "unit": "DoF/s", "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reyna is pure Python (no C sources to compile into oracle/_ref); its own loops run at ~5e3 DoF/s on one "
                "core (SURVEY.md section 6), the vectorised oracle port timed here is ~8x faster per core",
    }
    emit_json_line(line)


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes_assemble(fg, cd, dv, d: int) -> int:
    """Bytes the fused assembly kernel must read once and write once (DESIGN.md 'roofline'): geometry + adjacency +
    coefficient samples in, CSR values + load out."""
    b = 0
    for name in ("nodes", "tri", "tri_elem", "tri_off", "bbox", "area", "poly_off", "poly_vtx", "iedge", "ielem", "inorm"):
        b += getattr(fg, name).nbytes
    for k in ("vol_a", "vol_b", "vol_c", "vol_f", "if_a", "if_b", "if_a_mid", "if_upwind"):
        if k in cd:
            b += cd[k].numel() * cd[k].element_size()
    st = dv["structure"]
    b += st["adj_off"].numel() * 4 + st["n_items"] * 4 + st["grp_off"].numel() * 4
    nnz_struct = d * d * (fg.n_elem + st["n_groups"])
    b += 8 * nnz_struct + 8 * fg.n_elem * d
    return int(b)


def algorithmic_flops_assemble(fg, d: int, p: int, n_items: int, has_diff: bool, has_adv: bool) -> float:
    """FMA = 2 flops; volume: (2 [diffusion] + 1 [advection/reaction]) FMA per (i,j,q); faces: per (element, face) item one
    diagonal and one off-diagonal d x d block, 2 FMA per entry per point with diffusion, 1 without."""
    qv, qe = (p + 1) ** 2, p + 1
    vol = fg.n_tri * qv * (2.0 * d * d * ((2 if has_diff else 0) + 1) + 2.0 * d * 8)
    face = n_items * qe * (2.0 * 2 * d * d * (2 if has_diff else 1) + 2.0 * d * 12)
    return vol + face


def run_ours(args, rank: int, world: int, local_rank: int) -> None:
    import torch
    import torch.distributed as dist
    import reyna_b200 as rb
    from reyna_b200 import _lib
    from reyna_b200.problems import PROBLEMS

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload (rank 0 generates + caches, the others load) ------------------------------------------------------
    if rank == 0:
        geo, path, cached = workload_geometry(args.elements, verbose=True)
    barrier()
    if rank != 0:
        geo, path, cached = workload_geometry(args.elements)
    p = args.degree
    L = rb.lib()
    ddg = None
    if world > 1:
        from reyna_b200.distributed import DistributedDGFEM
        ddg = DistributedDGFEM(geo, polynomial_degree=p, device=dev)      # this rank's contiguous Morton range + halo plan
        dg = ddg.dg
    else:
        dg = rb.DGFEM(geo, polynomial_degree=p)
    dg.add_data(**PROBLEMS[args.problem]["data"])
    dg.keep_device_inputs = True
    dg.device = dev
    n_dof_total = geo.n_elements * ((p + 1) * (p + 2) // 2)
    d = (p + 1) * (p + 2) // 2

    # ---- e2e: DGFEM.dgfem(solve=False) from host data; the result (indptr, indices, data, load) lands in pinned host memory
    # inside the call (download_result: evaluate | upload | assemble | download are pipelined chunk by chunk) -----------------
    dg.download_result = True
    first_call_s = None
    for _ in range(args.warmup):
        tf = time.perf_counter()
        dg.dgfem(solve=False)
        if first_call_s is None:          # includes quadrature-point generation, pinned allocation, geometry upload
            first_call_s = time.perf_counter() - tf
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()                      # sampled over every timed region of this arm: e2e steps, device-only steps, roofline loop
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        dg.dgfem(solve=False)
        d2h = int(dg.timings["d2h_bytes"])
        assert dg._host is not None and dg._host["data"].shape[0] == dg._dev["data"].shape[0]
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    h2d = int(dg.timings.get("h2d_bytes", 0))
    e2e_breakdown = {k: v for k, v in dg.timings.items() if k.endswith("_s")}
    dg.download_result = False
    dg._host = None
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())

    # ---- value: inputs resident in HBM ----------------------------------------------------------------------------------
    fg, gd, cd, bd = dg._inputs
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    tm = {}
    # A step runs every kernel of the path (structure, pattern, prepass, volume, faces, boundary) and ends WITHOUT a host
    # round trip: the structure counters of this geometry are known from the first assembly (reyna_b200_structure_known) and
    # the exact-zero counter is read once after the timed loop (it decides whether scipy's zero pruning has to run; if it is
    # not zero the loop is repeated with the per-step check + compaction inside).
    def timed_steps(resolve):
        for _ in range(max(args.warmup, 20)):
            dv_ = dg._assemble_on_device(torch, L, dev, sp, fg, gd, cd, bd, tm, resolve=resolve)
        barrier()
        l0 = L.reyna_b200_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            dv_ = dg._assemble_on_device(torch, L, dev, sp, fg, gd, cd, bd, tm, resolve=resolve)
        e1.record()
        barrier()
        return dv_, int(L.reyna_b200_launch_count() - l0), e0, e1

    def timed_graph_steps():
        """The same step (all of its launches, no host round trip) captured ONCE into a CUDA graph and replayed: the ~14 short
        launches of a step no longer leave the GPU idle between kernels (matters for small partitions, N = 4..8)."""
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(cap):
            spc = C.c_void_p(cap.cuda_stream)
            for _ in range(3):
                dg._assemble_on_device(torch, L, dev, spc, fg, gd, cd, bd, tm, resolve=False)
            cap.synchronize()
            l0 = L.reyna_b200_launch_count()
            with torch.cuda.graph(graph, stream=cap):
                dv_ = dg._assemble_on_device(torch, L, dev, C.c_void_p(torch.cuda.current_stream().cuda_stream), fg, gd, cd, bd, tm,
                                             resolve=False)
            per_step = int(L.reyna_b200_launch_count() - l0)
            for _ in range(max(args.warmup, 20)):
                graph.replay()
            cap.synchronize()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cap)
            for _ in range(args.steps):
                graph.replay()
            e1.record(cap)
            cap.synchronize()
            barrier()
        return dv_, per_step * args.steps, e0, e1, graph

    dv, launches, ev0, ev1 = timed_steps(False)
    exact_zeros = int(dv["zero_count"].item())
    step_mode = "stream launches"
    graph_keep = None
    if exact_zeros != 0:
        dv, launches, ev0, ev1 = timed_steps(True)
        step_mode = "stream launches + per-step zero check / compaction"
    elif not args.no_graph:
        plain_ms = ev0.elapsed_time(ev1) / args.steps
        try:
            dvg, lg, g0, g1, graph_keep = timed_graph_steps()
            if int(dvg["zero_count"].item()) == exact_zeros and g0.elapsed_time(g1) > 0:
                dv, launches, ev0, ev1 = dvg, lg, g0, g1
                step_mode = f"one CUDA graph per step ({lg // args.steps} kernel nodes; {plain_ms:.4f} ms with plain stream launches)"
        except Exception as exc:      # capture unavailable: keep the plain measurement
            step_mode = f"stream launches (graph capture failed: {type(exc).__name__})"
            torch.cuda.synchronize()
    tm["exact_zeros"], tm["nnz"] = exact_zeros, int(dv["data"].numel())
    ms_step = ev0.elapsed_time(ev1) / args.steps
    if world > 1:
        tt = torch.tensor([ms_step], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_step = float(tt.item())

    # ---- roofline: the fused assembly kernel alone ------------------------------------------------------------------------
    n_roof = max(args.steps, 5)
    for _ in range(3):
        dg._relaunch_assemble(dv, sp)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_roof)]
    torch.cuda.synchronize()
    for a, b in evs:
        a.record()
        dg._relaunch_assemble(dv, sp)
        b.record()
    torch.cuda.synchronize()
    clock_info = clocks.stop()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    alg_bytes = algorithmic_bytes_assemble(fg, cd, dv, d)
    alg_flops = algorithmic_flops_assemble(fg, d, p, dv["structure"]["n_items"], "vol_a" in cd, "vol_b" in cd)
    peaks = {}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as fh:
            peaks = json.load(fh)
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        key = f"E{args.elements}_p{p}_{args.problem}_n{world}"
        traffic = tj.get(key)

    # FP64 FMA ceiling (measured with the library's own probe) for context
    scratch = torch.empty(148 * 8 * 256, dtype=torch.float64, device=dev)
    fl = C.c_double(0.0)
    for _ in range(2):
        L.reyna_b200_fma_probe(20000, scratch.data_ptr(), C.byref(fl), sp)
    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pa.record()
    L.reyna_b200_fma_probe(20000, scratch.data_ptr(), C.byref(fl), sp)
    pb.record()
    torch.cuda.synchronize()
    fp64_peak_tflops = fl.value / (pa.elapsed_time(pb) * 1e-3) / 1e12
    # FP64 tensor-core (DMMA m8n8k4) rate beside it: the A/B that decides whether the 6x6 contractions belong on tensor cores
    L.reyna_b200_dmma_probe(4000, scratch.data_ptr(), C.byref(fl), sp)
    pa.record()
    L.reyna_b200_dmma_probe(4000, scratch.data_ptr(), C.byref(fl), sp)
    pb.record()
    torch.cuda.synchronize()
    dmma_peak_tflops = fl.value / (pa.elapsed_time(pb) * 1e-3) / 1e12

    if not args.no_solve:
        # the second half of BASELINE.json's metric: dgfem(solve=True) wall time on the same workload (host callables,
        # upload, assembly, multigrid-preconditioned Krylov solve [NCCL halo exchange / all-reduce for N > 1], solution back
        # on the host)
        dg.keep_device_inputs = False
        dg._inputs = None
        del dv
        torch.cuda.empty_cache()
        if ddg is not None:
            ddg.dgfem(solve=True)          # first distributed solve creates the NCCL communicator (not timed)
        elif PROBLEMS[args.problem]["data"].get("diffusion") is None:
            dg.dgfem(solve=True)           # first flow-ordered sweep compiles the host-side level builder with numba (not timed)
        barrier()
        ts = time.perf_counter()
        if ddg is not None:
            ddg.dgfem(solve=True)
            si = dict(ddg.solve_info)
            sol_norm = float(np.linalg.norm(ddg.solution) ** 2)
        else:
            dg.dgfem(solve=True)
            si = dict(dg.solve_info)
        barrier()
        wall = time.perf_counter() - ts
        pr = PROBLEMS[args.problem]
        l2 = None
        if pr.get("exact") is not None:
            has_diff = pr["data"].get("diffusion") is not None
            if ddg is None:
                l2, dgn, h1 = dg.errors(pr["exact"], pr["grad_exact"] if has_diff else None)
            else:
                # the gathered solution of the distributed solve, measured on rank 0 with the single-GPU error kernels
                xg = ddg.gather_solution()
                if rank == 0:
                    chk = rb.DGFEM(geo, polynomial_degree=p)
                    chk.device = dev
                    chk.add_data(**pr["data"])
                    chk.orders = None
                    chk.dim_elem = d
                    chk.dim_system = d * geo.n_elements
                    chk.solution = xg
                    l2, dgn, h1 = chk.errors(pr["exact"], pr["grad_exact"] if has_diff else None)
                    del chk
        solve_line = {"dgfem_solve_true_wall_s": wall, "solve_s": si["seconds"], "setup_s": si["setup_seconds"],
                      "method": si["method"], "preconditioner": si["preconditioner"], "iterations": si["iterations"],
                      "true_rel_residual": si["rel_residual"], "stopped_at_rounding_floor": si["stagnated"],
                      "l2_error_vs_exact": l2, "n_gpus": world}
        if ddg is not None:
            solve_line["halo_exchanges"] = si["halo_exchanges"]
            solve_line["allreduces"] = si["allreduces"]
            ddg.close()
    else:
        solve_line = None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": "assembled DoF/s", "value": n_dof_total / (ms_step * 1e-3), "unit": "DoF/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.elements, p, args.problem), "elements": args.elements, "degree": p,
                   "problem": args.problem, "dof": n_dof_total, "nnz": int(tm.get("nnz", 0)),
                   "partition": f"{world} contiguous Morton-ordered element ranges" if world > 1 else "single GPU",
                   "l2": "inputs + outputs per step (GBs) far exceed the 126 MB L2; no explicit flush",
                   "step": step_mode},
        "e2e": {"value": n_dof_total / e2e_s, "unit": "DoF/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h),
                "s_per_step": e2e_s, "first_call_s": first_call_s, "host_threads": rb.sampling.host_threads(),
                "breakdown_last_step": e2e_breakdown,
                "note": "steady state: quadrature points, pinned staging buffers and device geometry are cached on the "
                        "geometry/solver objects after the first call (first_call_s)"},
        "gpu_launches": launches,
        "clocks": clock_info,
        "roofline": {"kernel": "dg_volume_kernel + dg_face_kernel (timed: the whole reyna_b200_assemble call = dg_face_sigma_kernel "
                               "prepass + volume kernel + face kernel, CUDA events on the launching stream)", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes": alg_bytes, "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_step,
                     "fp64": {"algorithmic_tflop": alg_flops / 1e12, "achieved_tflops": alg_flops / (k_ms * 1e-3) / 1e12,
                              "fma_probe_peak_tflops": fp64_peak_tflops,
                                       "dmma_probe_peak_tflops": dmma_peak_tflops}},
    }
    if solve_line is not None:
        line["solve"] = solve_line
    if world == 1 and not args.no_cpu_baseline:
        cores = 1
        ndof, times, _ = oracle_throughput(args.elements, args.baseline_sample, p, args.problem, cores, repeats=1)
        line["cpu_baseline"] = {
            "value": ndof / times[0], "unit": "DoF/s", "cores": cores, "kind": "port",
            "sample": f"oracle/dg_oracle.py (numpy port of reyna's dgfem(solve=False)) assembling the block rows of the first "
                      f"{min(args.baseline_sample, args.elements)} elements of the workload, {times[0]:.1f} s, 1 process"}
    emit_json_line(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def claim_stdout() -> None:
    """Keep stdout for the ONE JSON line: everything else written to fd 1 from here on (NCCL's version banner, library
    prints of child code) is routed to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json_line(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main() -> None:
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elements", type=int, default=1048576)
    ap.add_argument("--degree", type=int, default=2)
    ap.add_argument("--problem", default="adr")
    ap.add_argument("--config", type=int, default=None, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configuration (sets --elements/--degree/--problem): 1 README 1024 p1 adr, 2 16k p1 "
                         "diffusion, 3 65k p2 adr (use --degree 3 for 3b), 4 262k p2 hyperbolic, 5 1M p2 adr (default)")
    ap.add_argument("--baseline-sample", type=int, default=65536, help="elements of the cpu_baseline sample")
    ap.add_argument("--reference-sample", type=int, default=32768, help="elements per process and step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the device-resident step with plain stream launches only")
    ap.add_argument("--no-solve", action="store_true", help="skip the dgfem(solve=True) wall-time measurement")
    args = ap.parse_args()
    if args.config is not None:
        preset = {1: (1024, 1, "adr"), 2: (16384, 1, "diffusion"), 3: (65536, 2, "adr"), 4: (262144, 2, "hyperbolic"),
                  5: (1048576, 2, "adr")}[args.config]
        explicit = {a.split("=")[0] for a in sys.argv[1:] if a.startswith("--")}
        if "--elements" not in explicit:
            args.elements = preset[0]
        if "--degree" not in explicit:
            args.degree = preset[1]
        if "--problem" not in explicit:
            args.problem = preset[2]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
