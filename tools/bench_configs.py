#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (single GPU): PCG steps/s and ms per nonlinear
iteration through Opt_ProblemStep, CUDA events, inputs resident.  Not the driver's bench line (that is bench.py,
image_warping 2048^2); results are recorded in profiles/ and DESIGN.md."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(name, dims, spec, kind, host, dp, liters, steps=6, warmup=3, bytes_per_step=None):
    import torch
    from opt_b200 import OptSolver, spec_path, workloads
    dev = workloads.to_device(host)
    s = OptSolver(dims, spec_path(spec), kind, double_precision=dp)
    s.set_solver_parameters({"nIterations": steps + warmup + 1, "lIterations": liters})
    s.init(dev)
    c0 = s.current_cost()
    for _ in range(warmup):
        s.step(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loop = []
    e0.record()
    done = 0
    for _ in range(steps):
        if not s.step(dev):
            break
        done += 1; loop.append(s.pcg_loop_ms_last_step())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out = {"config": name, "kind": kind, "dtype": "f64" if dp else "f32", "lIterations": liters, "gn_iterations_timed": done,
           "ms_per_nonlinear_iter": ms / max(done, 1), "pcg_steps_per_s": (done * liters) / (ms / 1e3) if done else None,
           "pcg_loop_us_per_step": 1e3 * float(np.mean(loop)) / liters if loop else None, "launches_per_nonlinear_iter": s.launches_last_step(),
           "cost0": c0, "cost": s.current_cost()}
    if bytes_per_step and out["pcg_loop_us_per_step"]:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        out["algorithmic_bytes_per_pcg_step"] = bytes_per_step
        out["achieved_gbs"] = bytes_per_step / (out["pcg_loop_us_per_step"] * 1e-6) / 1e9
        out["frac_of_hbm_peak"] = out["achieved_gbs"] / peak
    print(json.dumps(out), flush=True)
    s.close()


def run_solve(name, dims, spec, kind, host, liters, iterations, repeats=5):
    """whole Opt_ProblemSolve calls (the reference apps' entry point), wall clock around the blocking call: once with the accept / reject
    decision on the host (default) and once with OPTB_DEVICE_LM=1 (all iterations after the first enqueued back to back)"""
    import time
    import torch
    from opt_b200 import OptSolver, spec_path, workloads
    for mode in ("host decision", "device decision"):
        os.environ.pop("OPTB_DEVICE_LM", None)
        if mode == "device decision":
            os.environ["OPTB_DEVICE_LM"] = "1"
        s = OptSolver(dims, spec_path(spec), kind)
        sp = {"nIterations": iterations, "lIterations": liters}
        best, cost, iters, dev_iters = None, None, None, None
        for _ in range(repeats + 1):                      # the first solve builds the graphs
            dev = workloads.to_device(workloads.clone_params(host))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            cost = s.solve(sp, dev)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            iters, dev_iters = int(s.read_scalar("nIter")), int(s.read_scalar("deviceIterations"))
            if _ > 0:
                best = dt if best is None else min(best, dt)
        print(json.dumps({"config": name, "mode": mode, "nIterations": iterations, "lIterations": liters, "iterations_done": iters, "iterations_on_device": dev_iters,
                          "ms_per_solve": 1e3 * best, "ms_per_nonlinear_iter": 1e3 * best / max(iters, 1), "pcg_steps_per_s_upper": iters * liters / best, "cost": cost}), flush=True)
        s.close()
    os.environ.pop("OPTB_DEVICE_LM", None)


ENGINE = " [generic engine]" if os.environ.get("OPTB_GENERIC") else ""


def main():
    from opt_b200 import workloads
    which = sys.argv[1:] or ["iw512", "sfs", "arap500k", "arap500k_generic", "rna"]   # OPTB_GENERIC=1: everything on the cached-Jacobian engine
    if "iw512" in which:
        run("image_warping 512x512 fp32 GN", [512, 512], "image_warping", "gaussNewtonGPU", workloads.image_warping(512, 512), False, 400, bytes_per_step=188 * 512 * 512)
    if "sfs" in which:
        run("shape_from_shading 640x480 fp32 LM (synthetic depth)" + ENGINE, [640, 480], "shape_from_shading", "LMGPU", workloads.shape_from_shading(640, 480), False, 10, steps=10,
            bytes_per_step=90 * 640 * 480)
    if "sfs_solve" in which:      # opt-in: the device-side LM decision against the default, whole solves as the app runs them (60 x 10)
        run_solve("shape_from_shading 640x480 fp32 LM (synthetic depth), Opt_ProblemSolve", [640, 480], "shape_from_shading", "LMGPU", workloads.shape_from_shading(640, 480), 10, 60)
    nx = 708
    nv = nx * nx
    if "arap500k" in which or "arap500k_generic" in which:
        host = workloads.arap_mesh(nx, nx, n_handles=32)
        ne = host[6]
        for tag in ("arap500k", "arap500k_generic"):
            if tag not in which:
                continue
            if tag.endswith("generic"):
                os.environ["OPTB_GENERIC"] = "1"
            run(f"arap_mesh_deformation {nv} vertices / {ne} directed edges fp32 GN" + (" [generic engine]" if tag.endswith("generic") else ""), [nv], "arap_mesh_deformation",
                "gaussNewtonGPU", workloads.clone_params(host), False, 100, bytes_per_step=396 * nv + 8 * ne)
            os.environ.pop("OPTB_GENERIC", None)
    if "rna" in which:
        n2 = 316
        host = workloads.robust_nonrigid(n2, n2, dtype=np.float64, perturb=0.05)   # perturbed: the unperturbed start converges inside the warm-up
        run(f"robust_nonrigid_alignment {n2 * n2} vertices fp64 LM" + ENGINE, [n2 * n2], "robust_nonrigid_alignment", "LMGPU",
            host, True, 250, steps=4, warmup=1, bytes_per_step=1136 * n2 * n2 + 8 * host[8])


if __name__ == "__main__":
    main()
