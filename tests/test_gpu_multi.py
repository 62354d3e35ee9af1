"""multi-GPU (slab) parity: needs >= 2 GPUs on the box, otherwise skipped"""
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
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
