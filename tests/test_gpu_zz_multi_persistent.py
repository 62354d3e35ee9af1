"""multi-GPU: the multi-rank form of the state-on-chip PCG loop (k_iwPcgLoopTile<.., MULTI>, OPTB_PERSISTENT=1 after
OptB200_PlanConnect) against the single-GPU solve.  Needs >= 2 GPUs, otherwise skipped.  Sorts last and is xfail-marked: the
kernel, gridAllReduceLLMulti and the host glue were written after the round's GPU budget was spent and have not run on
hardware; every case runs under torchrun in child processes, with a timeout."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="the multi-GPU persistent loop has not been run on hardware yet; remove this mark once it has passed")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,args,tile", [(2, ["128", "96"], None), (2, ["256", "130"], "64x8"), (4, ["192", "64"], None)])
def test_persistent_slab_solve_matches_single_gpu(built, world, args, tile):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + (os.getpid() % 400)
    env = dict(os.environ); env["OPTB_PERSISTENT"] = "1"; env["OPTB_TEST_MAX_LAUNCHES"] = "12"
    if tile:
        env["OPTB_RESIDENT_TILE"] = tile
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")] + args + ["gaussNewtonGPU"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
