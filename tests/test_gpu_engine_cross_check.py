"""GPU: shape_from_shading and robust_nonrigid_alignment have TWO implementations in libOpt.so -- the hand-written kernel sets
(energy_sfs.cu / energy_rna.cu, the default since they passed on a B200) and the generic cached-Jacobian engine (OPTB_GENERIC=1,
what every front-end generated energy runs on).  tests/test_gpu_generic_energies.py checks the default against the oracle; here the
generic engine is checked against the oracle for the same two energies (child process: the switch is read at plan time and the
library answers a CUDA error with exit(code)) and the two implementations are compared with each other on one solve."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _child(body, tag, switch="OPTB_GENERIC"):
    code = ("import os, sys\n"
            f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
            f"os.environ[{switch!r}] = '1'\n"
            "import test_gpu_generic_energies as g\n" + body + f"\nprint({tag!r})\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and tag in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.parametrize("energy,kinds", [("shape_from_shading", ["LMGPU", "gaussNewtonGPU"]), ("robust_nonrigid_alignment", ["LMGPU"])])
def test_generic_engine_matches_oracle(built, energy, kinds):
    lines = []
    for dp in (False, True):
        lines.append(f"g.test_kernel_parity(None, {energy!r}, {dp})")
        for kind in kinds:
            lines.append(f"g.test_solve_parity(None, {energy!r}, {kind!r}, {dp})")
    _child("\n".join(lines), "GENERIC_OK")


def test_rna_unmerged_gather_matches_oracle(built):
    """robust_nonrigid_alignment's J^T J gather merges an edge with its reverse when the graph is symmetric (every mesh the reference
    app builds); OPTB_RNA_NO_SYMMETRIC=1 keeps the two-loop form that arbitrary graphs take -- both must match the oracle"""
    lines = []
    for dp in (False, True):
        lines.append(f"g.test_kernel_parity(None, 'robust_nonrigid_alignment', {dp})")
        lines.append(f"g.test_solve_parity(None, 'robust_nonrigid_alignment', 'LMGPU', {dp})")
    _child("\n".join(lines), "UNMERGED_OK", switch="OPTB_RNA_NO_SYMMETRIC")


def test_arap_two_loop_gather_matches_oracle(built):
    """arap_mesh_deformation's gather merges an edge with its reverse on symmetric graphs; OPTB_ARAP_NO_SYMMETRIC=1 keeps the two-loop form
    that arbitrary graphs take (the default is checked by tests/test_gpu_generic_energies.py)"""
    lines = []
    for dp in (False, True):
        lines.append(f"g.test_kernel_parity(None, 'arap_mesh_deformation', {dp})")
        lines.append(f"g.test_solve_parity(None, 'arap_mesh_deformation', 'gaussNewtonGPU', {dp})")
    _child("\n".join(lines), "TWO_LOOP_OK", switch="OPTB_ARAP_NO_SYMMETRIC")


def test_hand_written_and_generic_agree(built):
    _child("""
import numpy as np
from opt_b200 import OptSolver, spec_path, workloads
host = workloads.shape_from_shading(64, 48)
def solve(generic):
    if generic: os.environ['OPTB_GENERIC'] = '1'
    else: os.environ.pop('OPTB_GENERIC', None)
    dev = workloads.to_device(workloads.clone_params(host))
    s = OptSolver([64, 48], spec_path('shape_from_shading'), 'LMGPU')
    s.set_solver_parameters({'nIterations': 3, 'lIterations': 10})
    c = s.solve({}, dev); x = dev[16].cpu().numpy().copy(); s.close()
    return c, x
ch, xh = solve(False); cg, xg = solve(True)
assert abs(ch - cg) <= 2e-4 * abs(cg), (ch, cg)            # fp32, different summation orders
assert np.linalg.norm(xh - xg) <= 1e-4 * np.linalg.norm(xg)
""", "CROSS_OK")
