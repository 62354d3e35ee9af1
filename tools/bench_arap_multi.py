#!/usr/bin/env python
"""arap_mesh_deformation ~500k vertices, strong scaling over N GPUs (vertex-slab decomposition).  Run under torchrun:
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_arap_multi.py
Prints one JSON line on rank 0 (PCG steps/s of the whole job, CUDA events, max over ranks)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from opt_b200 import OptSolver, spec_path, workloads
    from opt_b200.distributed import VertexSlab, connect_plan
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 708
    liters, steps, warmup = 100, 6, 3
    host = workloads.arap_mesh(nx, nx, n_handles=32)
    N = nx * nx
    dev = workloads.to_device(host)
    s = OptSolver([N], spec_path("arap_mesh_deformation"), "gaussNewtonGPU")
    if world > 1:
        slab = VertexSlab(N, world, rank, host[7], host[8])
        connect_plan(s, slab)
    s.set_solver_parameters({"nIterations": steps + warmup + 1, "lIterations": liters})
    s.init(dev)
    for _ in range(warmup):
        s.step(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop = []
    for _ in range(steps):
        s.step(dev); loop.append(s.pcg_loop_ms_last_step())
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        print(json.dumps({"config": f"arap_mesh_deformation {N} vertices / {host[6]} directed edges fp32 GN", "n_gpus": world, "pcg_steps_per_s": steps * liters / (ms / 1e3),
                          "ms_per_gn_iter": ms / steps, "pcg_loop_us_per_step": 1e3 * float(np.mean(loop)) / liters, "cost": s.current_cost(),
                          "halo_vertices": (slab.halo if world > 1 else 0)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
