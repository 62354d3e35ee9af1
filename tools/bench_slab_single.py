#!/usr/bin/env python
"""One GPU, image_warping W x H fp32 GN with lIterations=400: PCG-loop microseconds per step of the kernel-per-step loop and (when it
fits) the on-chip loop, for the per-rank slab sizes of the multi-GPU runs -- the no-communication floor those runs are compared with.
    python tools/bench_slab_single.py 2048 512"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(W, H, env):
    import numpy as np
    import torch
    from opt_b200 import OptSolver, spec_path, workloads
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        dev = workloads.to_device(workloads.image_warping(W, H, seed=0, mask="none"))
        s = OptSolver([W, H], spec_path("image_warping"), "gaussNewtonGPU")
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    s.set_solver_parameters({"nIterations": 12, "lIterations": 400})
    s.init(dev)
    for _ in range(3):
        s.step(dev)
    t = []
    for _ in range(6):
        s.step(dev); t.append(s.pcg_loop_ms_last_step())
    torch.cuda.synchronize()
    out = {"us_per_pcg_step": 1e3 * float(np.mean(t)) / 400, "launches": s.launches_last_step()}
    s.close()
    return out


if __name__ == "__main__":
    W, H = int(sys.argv[1]), int(sys.argv[2])
    rec = {"size": [W, H], "per_step_loop": run(W, H, {"OPTB_PERSISTENT": "0"}), "default": run(W, H, {})}
    print(json.dumps(rec), flush=True)
