"""GPU: the hand-written shape_from_shading kernel set (opt_b200/csrc/energy_sfs.cu, OPTB_SFS_HAND=1) against the oracle:
kernel level (cost, J^T F, J^T J x) and full LM / GN solves.  The per-pixel math is already pinned on the CPU
(tests/test_sfs_host_parity.py); this adds the launches and reductions.  Sorts last and runs each case in a child process
on purpose: the kernels were written after the round's GPU budget was spent, so their first run on hardware is the
round-end run, and the library answers a CUDA error with exit(code)."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="energy_sfs.cu has not been run on hardware yet (written after the round's GPU budget was "
                                                     "spent); remove this mark once it has passed on a B200")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _child(body):
    code = ("import os, sys\n"
            f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
            "os.environ['OPTB_SFS_HAND'] = '1'\n"
            "import test_gpu_generic_energies as g\n" + body + "\nprint('SFS_HAND_OK')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "SFS_HAND_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_hand_written_set_matches_oracle(built):
    """kernel level (cost, J^T F, J^T J x) in both precisions, then full solves -- all in ONE child process (each child pays the
    import of torch again; the GPU tier of the driver has a time limit)"""
    lines = []
    for dp in (False, True):
        lines.append(f"g.test_kernel_parity(None, 'shape_from_shading', {dp})")
        for kind in ['LMGPU', 'gaussNewtonGPU']:
            lines.append(f"g.test_solve_parity(None, 'shape_from_shading', {kind!r}, {dp})")
    _child("\n".join(lines))


def test_uses_fewer_launches_than_the_generic_engine(built):
    _child("""
import numpy as np
from opt_b200 import OptSolver, spec_path, workloads
host = workloads.shape_from_shading(64, 48)
def launches(hand):
    if hand: os.environ['OPTB_SFS_HAND'] = '1'
    else: os.environ.pop('OPTB_SFS_HAND', None)
    dev = workloads.to_device(workloads.clone_params(host))
    s = OptSolver([64, 48], spec_path('shape_from_shading'), 'LMGPU')
    s.set_solver_parameters({'nIterations': 2, 'lIterations': 10})
    s.init(dev); s.step(dev)
    n = s.launches_last_step(); c = s.current_cost(); s.close()
    return n, c
nh, ch = launches(True); ng, cg = launches(False)
assert nh < ng, (nh, ng)                       # one gather per J^T J instead of row pass + column pass
assert abs(ch - cg) <= 2e-4 * abs(cg), (ch, cg)
""")
