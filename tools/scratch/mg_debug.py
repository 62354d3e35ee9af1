import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from opt_b200 import OptSolver, spec_path, workloads
from opt_b200.distributed import Slab, scatter_image_params, connect_plan
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
W, H = 128, 96
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1
full = workloads.image_warping(W, H, mask="blob", seed=21)
slab = Slab(H, world, rank)
dev = workloads.to_device(scatter_image_params(full, (0, 1, 2, 3, 4), W, slab))
s = OptSolver([W, slab.local_rows], spec_path("image_warping"))
connect_plan(s, slab)
s.set_solver_parameters({"nIterations": 1, "lIterations": 40}); s.set_max_linear_iterations(K)
s.init(dev); s.step(dev)
ref = workloads.to_device(workloads.clone_params(full))
s1 = OptSolver([W, H], spec_path("image_warping"))
s1.set_solver_parameters({"nIterations": 1, "lIterations": 40}); s1.set_max_linear_iterations(K)
s1.init(ref); s1.step(ref)
for name in ("r", "p", "z", "delta", "Ap_X"):
    a = s.unflatten(s.read_vector(name)); b = s1.unflatten(s1.read_vector(name))
    for k, (x, y) in enumerate(zip(a, b)):
        x = x.reshape(slab.local_rows, W, -1); y = y.reshape(H, W, -1)[slab.lo:slab.hi]
        err = np.abs(x - y).max(axis=(1, 2)); sc = np.abs(y).max() + 1e-30
        rows = [(i, f"{e/sc:.1e}") for i, e in enumerate(err) if e / sc > 1e-4]
        print(f"rank {rank} {name}[{k}] owned=[{slab.owned_begin},{slab.owned_end}) bad rows: {rows[:12]}", flush=True)
dist.destroy_process_group()
