#!/usr/bin/env python
"""bench.py -- headline benchmark: PCG steps/s and ms per Gauss-Newton iteration, image_warping 2048^2.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is ONE NONLINEAR (Gauss-Newton) ITERATION of the hot path through Opt_ProblemStep:
PCGInit1, `lIterations` PCG iterations (PCGStep1/2/3), PCGLinearUpdate and computeCost
(reference API/src/solverGPUGaussNewton.t:1016-1177) on the configuration BASELINE.json's metric is
quoted on (image_warping 2048x2048, fp32, GN; lIterations = 400 as examples/image_warping/src/main.cpp:114).
value = PCG steps per second of the whole job (K * lIterations / time, all ranks), ms_per_step = ms per
GN iteration.  Timing: CUDA events, barrier + synchronize on both sides, max over ranks.  The working
set (~0.6 GB of live vectors) is far larger than the 126 MB L2, so no explicit L2 flush is needed.

Multi-GPU (N > 1): STRONG scaling of the same 2048^2 problem.  The image is cut into N row slabs
(SURVEY.md 8e), one rank per GPU; each rank plans slab + 1 halo row per neighbour, the plans are wired
together with OptB200_PlanConnect (CUDA IPC over NVLink): halo rows of z/p move by direct peer stores and
the two PCG scalars are summed through peer-memory mailboxes from inside the solver kernels -- no NCCL
call in the timed region (torch.distributed/NCCL is used only for setup and for the timing max-reduce).

--impl reference times the CPU oracle (oracle/, a C++/OpenMP port of the reference solver: the
reference's own Terra/CUDA path cannot be built without Terra) on the box's host cores on a
bounded slab of the same workload and prints the same JSON shape.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

W_FULL = H_FULL = 2048
L_ITERS = 400
U, D_, E_ = 3, 6, 1                       # unknown scalars, problem scalars, exclude scalars per pixel
BYTES_STEP1 = (U + D_ + U) * 4            # PCGStep1: U + D reads, U writes
BYTES_STEP2 = (5 * U + E_ + 3 * U) * 4    # PCGStep2
BYTES_STEP3 = (2 * U + E_ + U) * 4        # PCGStep3
BYTES_PCG_STEP = BYTES_STEP1 + BYTES_STEP2 + BYTES_STEP3   # 188 B / pixel / PCG step


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill(); out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample_reference(rows=256, pcg_iters=4, target_seconds=0.0):
    """kind "reference": matrix-free PCG steps built from the reference's OWN per-pixel functions (examples/image_warping/src/
    WarpingSolverEquations.h: evalMinusJTFDevice, applyJTJDevice) compiled for the host into oracle/_ref/libref_iw.so
    (oracle/ref_build.sh, oracle/ref_iw_shim.cpp ref_iw_pcg), OpenMP over the pixels, on a W_FULL x rows slab.
    Returns None when the library is not there (it is built where /root/reference is mounted and travels with the snapshot)."""
    import ctypes
    import numpy as np
    from opt_b200 import workloads
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "libref_iw.so")
    if not os.path.exists(so):
        return None
    L = ctypes.CDLL(so)
    if not hasattr(L, "ref_iw_pcg"):
        return None
    L.ref_iw_pcg.restype = ctypes.c_double
    host = workloads.image_warping(W_FULL, rows, seed=0, mask="none", n_constraints=3)
    off, ang, ur, cons, m, wf, wr = host[0], host[1], host[2], host[3], host[4], host[5], host[6]
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    nt = ctypes.c_int(0); rz = ctypes.c_double(0)
    t0 = time.time()
    # the hand-written solver takes the UNSQUARED weights (its energy is sum w r^2, twice Opt's convention: oracle/ref_iw_shim.cpp)
    call = lambda n: L.ref_iw_pcg(ctypes.c_int(W_FULL), ctypes.c_int(rows), fp(off), fp(ang), fp(ur), fp(m), fp(cons), ctypes.c_float(wf * wf), ctypes.c_float(wr * wr),
                                  ctypes.c_int(n), ctypes.byref(nt), ctypes.byref(rz))
    if target_seconds > 0:      # size the sample: a short calibration run, then enough iterations for about target_seconds of CPU work
        cal = call(4)
        pcg_iters = int(min(2000, max(pcg_iters, target_seconds / max(cal / 4, 1e-6))))
    sec = call(pcg_iters)
    wall = time.time() - t0
    if not (sec > 0) or not np.isfinite(rz.value):
        return None
    full_rate = (pcg_iters / sec) * rows / H_FULL           # cost is linear in pixels
    return {"value": full_rate, "unit": "PCG steps/s", "cores": int(nt.value), "kind": "reference",
            "sample": f"the reference's own evalMinusJTFDevice/applyJTJDevice (WarpingSolverEquations.h, compiled for the host: oracle/_ref) in a matrix-free "
                      f"PCG with OpenMP over the pixels, {W_FULL}x{rows} slab, {pcg_iters} PCG steps in {sec:.2f}s (wall incl. set-up {wall:.1f}s), "
                      f"scaled by {rows}/{H_FULL} to the full image"}, wall


def cpu_sample(steps=1, rows=256, pcg_iters=4, target_seconds=10.0):
    """CPU baseline on a W_FULL x rows slab of the workload, PCG steps/s scaled to the full image: the reference's own equations on the
    host when oracle/_ref is there (kind "reference"), else the CPU oracle (kind "port")"""
    try:
        ref = cpu_sample_reference(rows=rows, pcg_iters=max(pcg_iters, 8), target_seconds=target_seconds)
    except Exception as e:      # a broken oracle/_ref must not take the bench line down: fall back to the port and say so
        print(f"bench: reference CPU baseline unavailable ({type(e).__name__}: {e}); timing the oracle port instead", file=sys.stderr)
        ref = None
    if ref is not None:
        return ref
    from opt_b200 import workloads
    from oracle.oracle import Oracle, num_threads
    host = workloads.image_warping(W_FULL, rows, seed=0, mask="none", n_constraints=3)
    o = Oracle("image_warping", "gaussNewtonGPU", [W_FULL, rows])
    o.set_solver_parameter("nIterations", steps); o.set_solver_parameter("lIterations", pcg_iters)
    t0 = time.time()
    o.init(host)
    n = 0
    while o.step(host):
        n += 1
    wall = time.time() - t0
    lin_s = o.get_scalar("secondsLinear"); lin_n = o.get_scalar("linearSteps")
    slab_rate = lin_n / lin_s                       # PCG steps/s on the slab
    full_rate = slab_rate * rows / H_FULL           # cost is linear in pixels
    return {"value": full_rate, "unit": "PCG steps/s", "cores": num_threads(), "kind": "port",
            "sample": f"oracle (C++/OpenMP, explicit-J CGNR-style PCG) on a {W_FULL}x{rows} slab, {int(lin_n)} PCG steps in {lin_s:.1f}s "
                      f"(wall incl. linearisation {wall:.1f}s), scaled by {rows}/{H_FULL} to the full image"}, wall


def run_reference(args, rank, world):
    if rank != 0:
        return
    samples, t_all = [], time.time()
    for _ in range(args.warmup + args.steps):
        cb, wall = cpu_sample(steps=1, rows=512, pcg_iters=20)
        samples.append(cb)
        if time.time() - t_all > 240:
            break
    timed = samples[min(args.warmup, len(samples) - 1):]
    rate = statistics.median([s["value"] for s in timed])
    cb = dict(timed[-1]); cb["value"] = rate
    line = {"impl": "reference", "metric": "pcg_steps_per_s", "value": rate, "unit": "PCG steps/s", "n_gpus": args.gpus,
            "steps": len(timed), "warmup": args.warmup, "ms_per_step": 1000.0 * L_ITERS / rate, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"image_warping {W_FULL}x{H_FULL} fp32 GN, lIterations={L_ITERS}", "note": "reference arm on host cores (cpu_baseline.kind says which code: the reference's own equations from oracle/_ref, else the oracle port); bounded slab sample scaled to the full image"},
            "cpu_baseline": cb, "e2e": {"value": rate, "unit": "PCG steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="opt_b200")
    ap.add_argument("--liters", type=int, default=L_ITERS)
    ap.add_argument("--size", type=int, default=W_FULL)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from opt_b200 import OptSolver, spec_path, workloads
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from opt_b200.distributed import Slab, scatter_image_params, connect_plan
    W = H = args.size
    full = workloads.image_warping(W, H, seed=0, mask="none")          # every rank builds the same global problem ...
    slab = Slab(H, world, rank)
    host = scatter_image_params(full, (0, 1, 2, 3, 4), W, slab) if world > 1 else full   # ... and keeps its slab (+halo rows)
    del full
    Hr = slab.local_rows
    dev = workloads.to_device(host)
    pinned = [torch.from_numpy(p).pin_memory() if isinstance(p, np.ndarray) else p for p in host]
    solver = OptSolver([W, Hr], spec_path("image_warping"), "gaussNewtonGPU")
    if world > 1:
        connect_plan(solver, slab)
    total = args.warmup + args.steps
    solver.set_solver_parameters({"nIterations": total + 1, "lIterations": args.liters})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident arm
    solver.init(dev)
    for _ in range(args.warmup):
        solver.step(dev)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, loop_ms = 0, []
    ev0.record()
    for _ in range(args.steps):
        solver.step(dev)
        launches += solver.launches_last_step(); loop_ms.append(solver.pcg_loop_ms_last_step())
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    final_cost = solver.current_cost()

    # ------------------------------------------------------------------ end-to-end arm: host buffers in, unknowns out
    h2d = sum(p.numel() * p.element_size() for p in pinned if hasattr(p, "numel"))
    out_host = [torch.empty_like(pinned[0]).pin_memory(), torch.empty_like(pinned[1]).pin_memory()]
    d2h = sum(t.numel() * t.element_size() for t in out_host)
    solver.set_solver_parameters({"nIterations": 1, "lIterations": args.liters})

    def e2e_step():
        for d, p in zip(dev, pinned):
            if hasattr(d, "copy_"):
                d.copy_(p, non_blocking=True)
        solver.init(dev)
        solver.step(dev)
        out_host[0].copy_(dev[0], non_blocking=True); out_host[1].copy_(dev[1], non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ksteps = max(2, min(args.steps, 4))
    e0.record()
    for _ in range(ksteps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    # ------------------------------------------------------------------ per-kernel replay (roofline)
    prof = solver.profile_kernels(dev, reps=20) if world == 1 else {}

    t = torch.tensor([ms, e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = measured_peaks()
        pixels = W * H                      # whole job (all ranks)
        value = args.steps * args.liters / (ms / 1e3)          # PCG steps/s of the one shared problem
        e2e_value = ksteps * args.liters / (e2e_ms / 1e3)
        loop_step_us = 1e3 * statistics.mean(loop_ms) / args.liters
        # dominant kernel of the inner loop = the larger of the two per-iteration kernels in the CUDA-event replay
        # (same ordering as the ncu launch list, profiles/r1_ncu_iw2048.md); algorithmic bytes per SURVEY 8(d)
        loop_kernels = {"PCGStep2": BYTES_STEP2, "PCGStep3+PCGStep1(fused)": BYTES_STEP1 + BYTES_STEP3, "PCGStep1": BYTES_STEP1}
        cands = [k for k in ("PCGStep2", "PCGStep3+PCGStep1(fused)") if k in prof] or [k for k in ("PCGStep1",) if k in prof]
        traffic = None
        if cands:
            dom_name = max(cands, key=lambda k: prof[k])
            dom_bytes, dom_ms = loop_kernels[dom_name] * pixels, prof[dom_name]
            tp = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
            if os.path.exists(tp) and W == W_FULL and H == H_FULL:
                traffic = json.load(open(tp)).get(dom_name)
        else:                               # multi-GPU: no per-kernel replay; report the PCG step as the unit
            dom_name, dom_bytes, dom_ms = "PCG step (all kernels, rank 0 slab)", BYTES_PCG_STEP * W * (slab.g1 - slab.g0), loop_step_us * 1e-3
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        line = {
            "metric": "pcg_steps_per_s", "value": value, "unit": "PCG steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"image_warping {W}x{H} fp32 GN (BASELINE.json headline), lIterations={args.liters}, step = 1 GN iteration",
                       "parallelism": "single GPU" if world == 1 else f"{world} row slabs of {W}x{slab.g1 - slab.g0} (+1 halo row per neighbour), NVLink peer-memory halo exchange + scalar mailboxes",
                       "l2": "inputs larger than L2 (live PCG vectors ~%.0f MB per rank)" % (W * Hr * 4 * 3 * 7 / 1e6) if W * Hr * 84 > 126e6 else "per-rank working set %.0f MB fits the 126 MB L2 (no flush between steps: the solver itself streams >= 7 vectors per PCG step)" % (W * Hr * 4 * 3 * 7 / 1e6),
                       "final_cost": final_cost},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "PCG steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / ksteps, "what": "pinned host inputs -> H2D, Opt_ProblemInit + one Opt_ProblemStep via the C ABI, unknowns -> D2H"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic,
                         "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms,
                         "pcg_step": {"algorithmic_bytes_per_gpu": BYTES_PCG_STEP * W * (slab.g1 - slab.g0), "us_in_timed_region": loop_step_us,
                                      "achieved": BYTES_PCG_STEP * W * (slab.g1 - slab.g0) / (loop_step_us * 1e-6) / 1e9,
                                      "frac": BYTES_PCG_STEP * W * (slab.g1 - slab.g0) / (loop_step_us * 1e-6) / 1e9 / peak},
                         "kernels_ms": prof},
        }
        if not args.no_cpu_baseline and world == 1:
            cb, _ = cpu_sample(steps=1, rows=1024, pcg_iters=40)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
