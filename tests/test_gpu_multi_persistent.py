"""multi-rank form of the state-on-chip PCG loop (k_iwPcgLoopTile<.., MULTI> + gridAllReduceLLMulti: one collective cooperative kernel per
rank for the whole inner loop, taken by default after OptB200_PlanConnect when EVERY rank's slab fits on chip) against the single-GPU
solve; the check also asserts that the step really took the one-kernel path (launch count).  Ranks share a GPU when the box has
fewer GPUs than ranks (tests/test_gpu_multi.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cg", ["0", "1"])          # the two-reduction kernel (default: the reference's recurrences) and the opt-in single-reduction kernel
@pytest.mark.parametrize("world,args,tile", [(2, ["128", "96"], None), (2, ["256", "130"], "64x8"), (4, ["192", "64"], None)])
def test_persistent_slab_solve_matches_single_gpu(built, world, args, tile, cg):
    import torch
    port = 29500 + (os.getpid() % 400)
    env = dict(os.environ); env["OPTB_TEST_MAX_LAUNCHES"] = "12"       # the collective loop kernel is the default when every rank's slab fits
    env["OPTB_CG_LOOP"] = cg
    if cg == "1":
        env["OPTB_TEST_TOL_SCALE"] = "30"     # A p carried by a recurrence: equal to the reference's iterates in exact arithmetic, ~1e-3 apart in fp32 after 40 un-converged steps
    if torch.cuda.device_count() < world:
        env["OPTB_SHARE_GPU"] = "1"                                   # ranks time-slice one GPU (see tests/test_gpu_multi.py)
    if tile:
        env["OPTB_RESIDENT_TILE"] = tile
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")] + args + ["gaussNewtonGPU"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
