"""Two-process NCCL solve (one process per GPU) against the single-GPU solve of the same system.  Needs two GPUs."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(extra):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dist_check.py")] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:]
    return lines


def test_two_gpu_nccl_solve_matches_single_gpu():
    """Partitioned assembly + NCCL halo exchange / all-reduce Krylov solve with the replicated coarse level: same iteration
    counts as one GPU (within a check interval) and the same solution."""
    for r in _run(["65536", "adr,diffusion", "2"]):
        assert r["info"]["converged"], r
        assert r["rel_l2_vs_single_gpu"] <= 1e-11, r
        assert abs(r["info"]["iterations"] - r["single_gpu"]["iterations"]) <= 20, r


def test_two_gpu_solve_of_a_shuffled_numbering():
    """Random element numbering (the reference's mesher): partitions and aggregates are built on the Morton renumbering."""
    for r in _run(["16384", "adr", "2", "shuffle"]):
        assert r["info"]["converged"], r
        assert r["rel_l2_vs_single_gpu"] <= 1e-9, r
        assert r["info"]["n_ghost_elements"] < 2000, r            # a compact patch, not a random half of the mesh


def test_two_gpu_hyperbolic_uses_the_local_flow_sweep():
    """Advection-reaction without diffusion on two GPUs: the flow-ordered sweep of each rank's own rows preconditions BiCGStab
    (exact inside a partition): a handful of iterations instead of the ~1000 of block Jacobi."""
    for r in _run(["65536", "hyperbolic", "2"]):
        assert r["info"]["converged"] and r["info"]["preconditioner"] == "sweep", r
        assert r["info"]["iterations"] <= 60, r
        assert r["rel_l2_vs_single_gpu"] <= 1e-9, r
