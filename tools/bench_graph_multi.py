#!/usr/bin/env python
"""Graph-domain configs of BASELINE.json, strong scaling over N GPUs (vertex-slab decomposition, opt_b200.distributed.VertexSlab):
    --workload arap500k   arap_mesh_deformation, 708x708 grid mesh (501 264 vertices), fp32 GN, lIterations 100   (config 4: 1 -> 4 GPUs)
    --workload rna1m      robust_nonrigid_alignment, 1000x1000 grid mesh, fp64 LM, lIterations 250                (config 5: 8 GPUs)
Run under torchrun (or plain python for 1 GPU):
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_graph_multi.py --workload rna1m
Prints one JSON line on rank 0 (PCG steps/s of the whole job, CUDA events, max over ranks)."""
import argparse
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="arap500k", choices=["arap500k", "rna1m"])
    ap.add_argument("--nx", type=int, default=0)
    a = ap.parse_args()
    if a.workload == "arap500k":
        nx = a.nx or 708
        liters, steps, warmup = 100, 6, 3
        host = workloads.arap_mesh(nx, nx, n_handles=32)
        spec, kind, dp, v0, v1, ne = "arap_mesh_deformation", "gaussNewtonGPU", False, host[7], host[8], host[6]
        bytes_per_step = 396 * nx * nx + 8 * ne
    else:
        nx = a.nx or 1000
        liters, steps, warmup = 250, 3, 1
        host = workloads.robust_nonrigid(nx, nx, dtype=np.float64, perturb=0.05)
        spec, kind, dp, v0, v1, ne = "robust_nonrigid_alignment", "LMGPU", True, host[9], host[10], host[8]
        bytes_per_step = 1136 * nx * nx + 8 * ne
    N = nx * nx
    dev = workloads.to_device(host)
    s = OptSolver([N], spec_path(spec), kind, double_precision=dp)
    if world > 1:
        slab = VertexSlab(N, world, rank, v0, v1)
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
    done = 0
    for _ in range(steps):
        if not s.step(dev):
            break
        done += 1; loop.append(s.pcg_loop_ms_last_step())
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        us = 1e3 * float(np.mean(loop)) / liters if loop else None
        print(json.dumps({"config": f"{spec} {N} vertices / {ne} directed edges {'fp64' if dp else 'fp32'} {'LM' if 'LM' in kind else 'GN'}, lIterations={liters}", "n_gpus": world,
                          "nonlinear_iterations_timed": done, "pcg_steps_per_s": (done * liters / (ms / 1e3)) if done else None,
                          "ms_per_nonlinear_iter": ms / max(done, 1), "pcg_loop_us_per_step": us, "cost": s.current_cost(),
                          "algorithmic_bytes_per_pcg_step": bytes_per_step,
                          "frac_of_hbm_peak_per_gpu": (bytes_per_step / world / (us * 1e-6) / 1e9 / peak) if us else None,
                          "halo_vertices": (slab.halo if world > 1 else 0)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
