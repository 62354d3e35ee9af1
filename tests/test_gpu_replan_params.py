"""GPU: re-solving on ONE plan with different problem parameters.  The inner PCG loop is a cached CUDA graph whose kernel nodes hold
their arguments by value (weights, caller pointers, neighbour-list buffers the energy reallocates when the graph grows); the plan
re-captures when that fingerprint changes (EnergyOps::captureKey, solver.cu runLoop).  Each case solves A then B on one plan and
compares B with a fresh plan that launches its kernels directly (OPTB_NO_GRAPH=1) -- the reference re-reads problemparams at every
Init/Step (util.t:664-692), so the two must agree exactly."""
import os

import numpy as np
import pytest

from opt_b200 import workloads

pytestmark = pytest.mark.gpu


def _solve(plan_env, energy, kind, dims, problems, sp):
    """solve every problem in `problems` in turn on one plan; returns (cost, unknown arrays) of the last"""
    from opt_b200 import OptSolver, spec_path
    old = {k: os.environ.get(k) for k in plan_env}
    os.environ.update(plan_env)
    try:
        s = OptSolver(dims, spec_path(energy), kind)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    out = None
    for host, upos in problems:
        dev = workloads.to_device(workloads.clone_params(host))
        cost = s.solve(sp, dev)
        out = (cost, [dev[p].cpu().numpy().copy() for p in upos])
    s.close()
    return out


@pytest.mark.parametrize("kind", ["gaussNewtonGPU", "LMGPU"])
def test_image_warping_weights_change_between_solves(built, kind):
    W, H = 96, 80
    a = workloads.image_warping(W, H, mask="blob", seed=3, perturb=0.3)
    b = workloads.clone_params(a); b[5] = float(np.sqrt(7.0)); b[6] = float(np.sqrt(0.5))      # other w_fit, w_reg
    sp = {"nIterations": 3, "lIterations": 30}
    env = {"OPTB_PERSISTENT": "0"}                                                               # the CUDA-graph loop (the on-chip loop takes its arguments per launch)
    cost2, x2 = _solve(env, "image_warping", kind, [W, H], [(a, (0, 1)), (b, (0, 1))], sp)
    cost1, x1 = _solve(dict(env, OPTB_NO_GRAPH="1"), "image_warping", kind, [W, H], [(b, (0, 1))], sp)
    assert cost2 == cost1 and all(np.array_equal(u, v) for u, v in zip(x2, x1))
    costa, _ = _solve(env, "image_warping", kind, [W, H], [(a, (0, 1))], sp)
    assert abs(costa - cost1) > 1e-3 * abs(cost1)           # the two problems really differ


def test_arap_graph_grows_between_solves(built):
    nx, ny = 40, 36
    full = workloads.arap_mesh(nx, ny, n_handles=10, perturb=0.05)
    small = workloads.clone_params(full)
    keep = (np.arange(full[6]) % 5) != 0                     # drop every fifth directed edge: fewer edges, other neighbour lists
    small[7] = np.ascontiguousarray(full[7][keep]); small[8] = np.ascontiguousarray(full[8][keep]); small[6] = int(keep.sum())
    sp = {"nIterations": 3, "lIterations": 25}
    N = nx * ny
    cost2, x2 = _solve({}, "arap_mesh_deformation", "gaussNewtonGPU", [N], [(small, (2, 3)), (full, (2, 3))], sp)
    cost1, x1 = _solve({"OPTB_NO_GRAPH": "1"}, "arap_mesh_deformation", "gaussNewtonGPU", [N], [(full, (2, 3))], sp)
    assert cost2 == cost1 and all(np.array_equal(u, v) for u, v in zip(x2, x1))


def test_generic_engine_graph_grows_between_solves(built):
    """same through the cached-Jacobian engine (its Jacobian buffers are reallocated when the edge count grows)"""
    nx, ny = 24, 20
    full = workloads.arap_mesh(nx, ny, n_handles=6, perturb=0.05)
    small = workloads.clone_params(full)
    keep = (np.arange(full[6]) % 3) != 0
    small[7] = np.ascontiguousarray(full[7][keep]); small[8] = np.ascontiguousarray(full[8][keep]); small[6] = int(keep.sum())
    sp = {"nIterations": 2, "lIterations": 20}
    N = nx * ny
    cost2, x2 = _solve({"OPTB_GENERIC": "1"}, "arap_mesh_deformation", "LMGPU", [N], [(small, (2, 3)), (full, (2, 3))], sp)
    cost1, x1 = _solve({"OPTB_GENERIC": "1", "OPTB_NO_GRAPH": "1"}, "arap_mesh_deformation", "LMGPU", [N], [(full, (2, 3))], sp)
    assert cost2 == cost1 and all(np.array_equal(u, v) for u, v in zip(x2, x1))
