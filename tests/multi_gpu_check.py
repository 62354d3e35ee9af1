"""Run under torchrun (one rank per GPU): slab-decomposed image_warping solve vs the single-GPU solve of the
same problem (and vs the CPU oracle at this size).  Used by tests/test_gpu_multi.py and by hand:
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _setup(torch, dist):
    """one rank per GPU over NCCL; with fewer GPUs than ranks (OPTB_SHARE_GPU=1, e.g. the driver's 1-GPU test box) the ranks share
    GPUs -- CUDA IPC, the peer-memory mailboxes and the halo stores work the same between two processes on one device (the
    processes time-slice, so it is slow) -- and torch.distributed's plumbing runs over gloo (NCCL refuses two ranks on one GPU)"""
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    ngpu = torch.cuda.device_count()
    if os.environ.get("OPTB_SHARE_GPU") and ngpu < world:
        torch.cuda.set_device(local % ngpu)
        dist.init_process_group("gloo")
        return rank, world, local, "cpu"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, "cuda"


def main_arap():
    """vertex-slab decomposition of arap_mesh_deformation vs the single-GPU solve"""
    import torch
    import torch.distributed as dist
    from opt_b200 import OptSolver, spec_path, workloads
    from opt_b200.distributed import VertexSlab, connect_plan
    rank, world, local, flag_dev = _setup(torch, dist)
    nx, ny = int(sys.argv[2]), int(sys.argv[3]); kind = sys.argv[4] if len(sys.argv) > 4 else "gaussNewtonGPU"
    sp = {"nIterations": 3, "lIterations": 30}
    full = workloads.arap_mesh(nx, ny, n_handles=12)
    N = nx * ny
    slab = VertexSlab(N, world, rank, full[7], full[8])
    dev = workloads.to_device(workloads.clone_params(full))           # every rank binds the global arrays
    s = OptSolver([N], spec_path("arap_mesh_deformation"), kind)
    connect_plan(s, slab)
    s.set_solver_parameters(sp)
    s.init(dev); costs = [s.current_cost()]
    while s.step(dev):
        costs.append(s.current_cost())
    torch.cuda.synchronize()
    off = slab.gather(dev[2].cpu().numpy()); ang = slab.gather(dev[3].cpu().numpy())
    ok = True
    if rank == 0:
        ref = workloads.to_device(workloads.clone_params(full))
        s1 = OptSolver([N], spec_path("arap_mesh_deformation"), kind); s1.set_solver_parameters(sp)
        s1.init(ref); c1 = [s1.current_cost()]
        while s1.step(ref):
            c1.append(s1.current_cost())
        o1, a1 = ref[2].cpu().numpy(), ref[3].cpu().numpy()
        rel_c = max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(costs, c1))
        e_off = np.linalg.norm(off - o1) / np.linalg.norm(o1); e_ang = np.linalg.norm(ang - a1) / max(np.linalg.norm(a1), 1e-12)
        print(f"multi_gpu_check ARAP world={world} {nx}x{ny} {kind} halo={slab.halo}: costs {costs[0]:.6f} -> {costs[-1]:.6f}; vs 1 GPU: cost rel {rel_c:.2e}, Offset {e_off:.2e}, Angle {e_ang:.2e}")
        print("  multi :", [f"{c:.6g}" for c in costs]); print("  single:", [f"{c:.6g}" for c in c1])
        ok = rel_c < 1e-4 and e_off < 1e-5 and e_ang < 2e-3 and len(costs) == len(c1)
    flag = torch.tensor([1 if ok else 0], device=flag_dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def main_rna():
    """vertex-slab decomposition of robust_nonrigid_alignment (fp64, LM: BASELINE.json config 5) vs the single-GPU solve"""
    import torch
    import torch.distributed as dist
    from opt_b200 import OptSolver, spec_path, workloads
    from opt_b200.distributed import VertexSlab, connect_plan
    rank, world, local, flag_dev = _setup(torch, dist)
    nx, ny = int(sys.argv[2]), int(sys.argv[3]); kind = sys.argv[4] if len(sys.argv) > 4 else "LMGPU"
    dp = (sys.argv[5] if len(sys.argv) > 5 else "f64") == "f64"
    sp = {"nIterations": 3, "lIterations": 30}
    full = workloads.robust_nonrigid(nx, ny, dtype=np.float64 if dp else np.float32, perturb=0.05)
    N = nx * ny
    slab = VertexSlab(N, world, rank, full[9], full[10])
    dev = workloads.to_device(workloads.clone_params(full))           # every rank binds the global arrays
    s = OptSolver([N], spec_path("robust_nonrigid_alignment"), kind, double_precision=dp)
    connect_plan(s, slab)
    s.set_solver_parameters(sp)
    s.init(dev); costs = [s.current_cost()]
    while s.step(dev):
        costs.append(s.current_cost())
    torch.cuda.synchronize()
    got = [slab.gather(dev[k].cpu().numpy()) for k in (2, 3, 4)]
    ok = True
    if rank == 0:
        ref = workloads.to_device(workloads.clone_params(full))
        s1 = OptSolver([N], spec_path("robust_nonrigid_alignment"), kind, double_precision=dp); s1.set_solver_parameters(sp)
        s1.init(ref); c1 = [s1.current_cost()]
        while s1.step(ref):
            c1.append(s1.current_cost())
        want = [ref[k].cpu().numpy() for k in (2, 3, 4)]
        rel_c = max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(costs, c1)) if len(costs) == len(c1) else float("inf")
        errs = [float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12)) for a, b in zip(got, want)]
        print(f"multi_gpu_check RNA world={world} {nx}x{ny} {kind} {'f64' if dp else 'f32'} halo={slab.halo}: costs {costs[0]:.8f} -> {costs[-1]:.8f}; vs 1 GPU: cost rel {rel_c:.2e}, "
              f"Offset {errs[0]:.2e}, Angle {errs[1]:.2e}, RobustWeights {errs[2]:.2e}")
        print("  multi :", [f"{c:.8g}" for c in costs]); print("  single:", [f"{c:.8g}" for c in c1])
        tol = (1e-9, 1e-8) if dp else (1e-4, 2e-3)
        ok = rel_c < tol[0] and max(errs) < tol[1] and len(costs) == len(c1)
    flag = torch.tensor([1 if ok else 0], device=flag_dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "arap":
        return main_arap()
    if len(sys.argv) > 1 and sys.argv[1] == "rna":
        return main_rna()
    import torch
    import torch.distributed as dist
    from opt_b200 import OptSolver, spec_path, workloads
    from opt_b200.distributed import Slab, scatter_image_params, connect_plan, gather_owned_rows
    rank, world, local, flag_dev = _setup(torch, dist)
    W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 96)
    kind = sys.argv[3] if len(sys.argv) > 3 else "gaussNewtonGPU"
    sp = {"nIterations": 3, "lIterations": 40}
    full = workloads.image_warping(W, H, mask="blob", seed=21)
    slab = Slab(H, world, rank)
    local_host = scatter_image_params(full, (0, 1, 2, 3, 4), W, slab)
    dev = workloads.to_device(local_host)
    s = OptSolver([W, slab.local_rows], spec_path("image_warping"), kind)
    connect_plan(s, slab)
    s.set_solver_parameters(sp)
    s.init(dev)
    costs = [s.current_cost()]
    while s.step(dev):
        costs.append(s.current_cost())
    torch.cuda.synchronize()
    launches = s.launches_last_step()
    max_launches = int(os.environ.get("OPTB_TEST_MAX_LAUNCHES", "0"))    # > 0: the caller expects the one-launch persistent loop
    off = gather_owned_rows(dev[0].cpu().numpy().reshape(-1, 2), W, slab).reshape(-1, 2)
    ang = gather_owned_rows(dev[1].cpu().numpy(), W, slab).reshape(-1)
    all_costs = [None] * world
    dist.all_gather_object(all_costs, costs)
    ok = True
    if rank == 0:
        assert all(c == all_costs[0] for c in all_costs), "ranks disagree on the cost history (must be bitwise equal)"
        ref = workloads.to_device(workloads.clone_params(full))
        s1 = OptSolver([W, H], spec_path("image_warping"), kind)
        s1.set_solver_parameters(sp)
        s1.init(ref); c1 = [s1.current_cost()]
        while s1.step(ref):
            c1.append(s1.current_cost())
        o1, a1 = ref[0].cpu().numpy(), ref[1].cpu().numpy()
        rel_c = max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(costs, c1))
        e_off = np.linalg.norm(off - o1) / np.linalg.norm(o1); e_ang = np.linalg.norm(ang - a1) / max(np.linalg.norm(a1), 1e-12)
        print(f"multi_gpu_check world={world} {W}x{H} {kind}: costs {costs[0]:.6f} -> {costs[-1]:.6f}; vs 1 GPU: cost rel {rel_c:.2e}, Offset {e_off:.2e}, Angle {e_ang:.2e}")
        print("  multi :", [f"{c:.6g}" for c in costs]); print("  single:", [f"{c:.6g}" for c in c1])
        tol = float(os.environ.get("OPTB_TEST_TOL_SCALE", "1"))       # > 1: paths whose iterates are not the reference's bit for bit (single-reduction loop)
        ok = rel_c < 1e-4 * tol and e_off < 1e-5 * tol and e_ang < 2e-3 * tol and len(costs) == len(c1)
        print(f"  launches in the last step: {launches}")
        if max_launches > 0 and launches > max_launches:
            print(f"  expected at most {max_launches} launches: the persistent multi-GPU loop was not taken"); ok = False
    flag = torch.tensor([1 if ok else 0], device=flag_dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
