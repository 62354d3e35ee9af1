"""multi-rank (slab) parity against the single-GPU solve, one process per rank under torchrun.  With fewer GPUs than ranks (the driver's
test box has one) the ranks SHARE a GPU (OPTB_SHARE_GPU=1, tests/multi_gpu_check.py): CUDA IPC, the peer-memory mailboxes, the halo
stores and the collective loop kernel run unchanged between processes on one device -- the processes time-slice, so it is slower, but
it is the same code path as over NVLink, which the builder's multi-GPU runs and bench.py's parity_vs_n1 record cover."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,args", [(2, ["128", "96", "gaussNewtonGPU"]), (2, ["256", "130", "LMGPU"]), (4, ["192", "64", "gaussNewtonGPU"]),
                                        (2, ["arap", "48", "40", "gaussNewtonGPU"]), (2, ["arap", "36", "50", "LMGPU"]), (4, ["arap", "32", "96", "gaussNewtonGPU"]),
                                        (2, ["rna", "40", "48", "LMGPU", "f64"]), (2, ["rna", "36", "50", "gaussNewtonGPU", "f32"]), (4, ["rna", "32", "96", "LMGPU", "f64"])])
def test_slab_solve_matches_single_gpu(built, world, args):
    import torch
    env = dict(os.environ)
    if torch.cuda.device_count() < world:
        env["OPTB_SHARE_GPU"] = "1"
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
