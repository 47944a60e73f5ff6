"""Distributed solve check (torchrun, N GPUs): the partitioned NCCL solve against the single-GPU solve of the same system.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_check.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import reyna_b200 as rb
from reyna_b200.distributed import DistributedDGFEM
from reyna_b200.problems import PROBLEMS
from bench import workload_geometry


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [65536]
    probs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["adr", "diffusion"]
    p = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    shuffle = len(sys.argv) > 4 and sys.argv[4] == "shuffle"      # element numbering without locality (the reference's meshes)
    for E in sizes:
        if rank == 0:
            workload_geometry(E)
        dist.barrier()
        geo, _, _ = workload_geometry(E)
        if shuffle:
            from reyna_b200.partition import RenumberedGeometry
            geo = RenumberedGeometry(geo, np.random.default_rng(5).permutation(geo.n_elements))
        for name in probs:
            ddg = DistributedDGFEM(geo, p, device=dev)
            ddg.add_data(**PROBLEMS[name]["data"])
            ddg.dgfem(solve=False)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            ddg.solution = ddg.solve()
            torch.cuda.synchronize()
            dist.barrier()
            dt = time.perf_counter() - t0
            x = ddg.gather_solution()
            out = dict(E=E, problem=name, p=p, world=world, shuffled=bool(shuffle), solve_s=round(dt, 3), info=ddg.solve_info)
            if rank == 0 and E <= 262144:
                dg = rb.DGFEM(geo, p)
                dg.device = dev
                dg.add_data(**PROBLEMS[name]["data"])
                dg.dgfem(solve=True)
                out["rel_l2_vs_single_gpu"] = float(np.linalg.norm(x - dg.solution) / np.linalg.norm(dg.solution))
                out["single_gpu"] = dict(seconds=dg.solve_info["seconds"], iterations=dg.solve_info["iterations"])
                del dg
            if rank == 0:
                print(json.dumps(out), flush=True)
            ddg.close()
            del ddg
            torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
