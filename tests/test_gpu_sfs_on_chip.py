"""GPU: shape_from_shading's tile kernels (energy_sfs.cu / sfs_tile.cuh) -- the fused LM head k_sfsInitTiled (PCGInit1 + CtC +
FinalizeDiagonal), the whole-loop kernel k_sfsLoopResident (state on chip, residual resets and the zeta early-out inside) and the cost
folded into precompute -- against the oracle and against the kernel-per-step path of the same library (OPTB_PERSISTENT=0,
OPTB_SFS_NO_FUSED_INIT=1, OPTB_SFS_NO_FUSED_COST=1).  The per-thread functions are also run on the CPU by tests/test_sfs_host_parity.py;
what only the GPU can show is the synchronisation (tile-edge exchange released by the all-reduce, cooperative launch inside a graph)."""
import contextlib
import os

import numpy as np
import pytest

from opt_b200 import workloads

pytestmark = pytest.mark.gpu
PER_STEP = {"OPTB_PERSISTENT": "0", "OPTB_SFS_NO_FUSED_INIT": "1", "OPTB_SFS_NO_FUSED_COST": "1"}


@contextlib.contextmanager
def _env(**env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _plan(W, H, kind, env, dp=False):
    from opt_b200 import OptSolver, spec_path
    with _env(**env):
        return OptSolver([W, H], spec_path("shape_from_shading"), kind, double_precision=dp)


def _state_after(host, W, H, kind, k, env, dp=False):
    dev = workloads.to_device(workloads.clone_params(host))
    s = _plan(W, H, kind, env, dp)
    s.set_solver_parameter("nIterations", 1); s.set_solver_parameter("lIterations", 50)
    s.set_max_linear_iterations(k)
    s.init(dev); s.step(dev)
    out = {n: s.read_vector(n).copy() for n in ("delta", "r", "z", "p")}
    scal = {n: s.read_scalar(n) for n in ("q", "flags")}
    launches = s.launches_last_step()
    s.close()
    return out, scal, launches


@pytest.mark.parametrize("W,H", [(64, 48), (45, 37), (200, 150), (640, 480)])
@pytest.mark.parametrize("kind", ["LMGPU", "gaussNewtonGPU"])
def test_loop_on_chip_matches_per_step_path_and_oracle(built, W, H, kind):
    from oracle.oracle import Oracle
    host = workloads.shape_from_shading(W, H)
    lm = kind == "LMGPU"
    for k in ((3, 9, 10, 25, 45) if lm else (3, 12)):
        ref, sref, n_ref = _state_after(host, W, H, kind, k, PER_STEP)
        got, sgot, n_got = _state_after(host, W, H, kind, k, {})
        assert n_got < n_ref and n_got <= 4, (n_got, n_ref)      # head (LM: one launch) + the whole loop in one launch
        o = Oracle("shape_from_shading", kind, [W, H])
        oh = workloads.clone_params(host)
        o.set_solver_parameter("nIterations", 1); o.set_solver_parameter("lIterations", 50)
        o.set_max_linear_iterations(k)
        r0max = float(np.abs(o.eval_jtf(oh)[1]).max())
        o.init(oh); o.step(oh)
        # delta is what the solve uses; r and p are compared while the two fp32 CG trajectories (different summation orders, and the oracle's
        # scalar arithmetic) have not drifted apart yet: measured on a B200 at k = 25 / 45, |r - r_oracle| reaches 2e-3 / 1.4e-2 of |r|_max
        # (200x150 / 640x480) with delta still inside 5e-4, so beyond k = 12 only delta is checked
        for name in (("delta", "r", "p") if k <= 12 else ("delta",)):
            b = o.get_vector(name)
            scale = max(1e-30, float(np.abs(b).max()))
            # r = b - A delta after a reset: a few ulp of |J^T F| (measured 1.1e-5 at 64x48, k = 25; 2.4e-5 at 640x480 GN, k = 12); p = z + beta p carries the same absolute noise
            floor = 5e-5 * r0max if name in ("r", "p") else 0.0
            rel = (5e-4 if k < 25 else 2e-3) if name == "delta" else 2e-3
            assert np.abs(got[name][:b.size] - b).max() <= max(rel * scale, floor), (name, k, "oracle")
            assert np.abs(got[name] - ref[name]).max() <= max(rel * scale, floor), (name, k, "per-step path")
        if lm:
            # the same early-out iteration (zeta is a difference of nearly equal q's: one iteration of slack for the two summation orders)
            assert abs((sgot["flags"] or k + 1) - (sref["flags"] or k + 1)) <= (1 if k <= 12 else 3), (k, sgot, sref)
            assert abs(sgot["q"] - sref["q"]) <= 1e-3 * max(1e-30, abs(sref["q"]))


@pytest.mark.parametrize("dp", [False, True])
def test_small_images_in_double_and_float(built, dp):
    """sizes that cut tiles on both axes; fp64 takes the kernel too while its tile fits the shared memory"""
    from oracle.oracle import Oracle
    for W, H in ((33, 9), (5, 4), (70, 19), (96, 80)):
        host = workloads.shape_from_shading(W, H, dtype=np.float64 if dp else np.float32)
        got, sgot, n_got = _state_after(host, W, H, "LMGPU", 12, {}, dp)
        o = Oracle("shape_from_shading", "LMGPU", [W, H], double_precision=dp)
        oh = workloads.clone_params(host)
        o.set_solver_parameter("nIterations", 1); o.set_solver_parameter("lIterations", 50)
        o.set_max_linear_iterations(12)
        o.init(oh); o.step(oh)
        for name in ("delta", "p"):
            b = o.get_vector(name)
            assert np.abs(got[name][:b.size] - b).max() <= (1e-8 if dp else 5e-4) * max(1e-30, float(np.abs(b).max())), (W, H, name)


@pytest.mark.parametrize("W,H", [(64, 48), (640, 480)])
def test_full_solve_matches_oracle_and_per_step_path(built, W, H):
    """whole LM solves: the iteration graph replays the cooperative loop kernel (its reduction sequence lives in device memory)"""
    from oracle.oracle import Oracle
    host = workloads.shape_from_shading(W, H)
    sp = {"nIterations": 8, "lIterations": 10}

    def solve(env):
        dev = workloads.to_device(workloads.clone_params(host))
        s = _plan(W, H, "LMGPU", env)
        c = s.solve(sp, dev)
        x = dev[16].cpu().numpy().copy()
        s.close()
        return c, x

    c1, x1 = solve({})
    c0, x0 = solve(PER_STEP)
    for env in ({"OPTB_SFS_NO_HEAD_IN_LOOP": "1"}, {"OPTB_SFS_NO_FUSED_PRECOMPUTE": "1"}, {"OPTB_SFS_LOOP_NT": "1024"}):
        c2, x2 = solve(env)
        assert abs(c2 - c0) <= 5e-4 * abs(c0), (env, c2, c0)
    o = Oracle("shape_from_shading", "LMGPU", [W, H])
    for k, v in sp.items():
        o.set_solver_parameter(k, v)
    oh = workloads.clone_params(host)
    oc = o.solve(oh)
    assert abs(c1 - oc) <= 5e-4 * abs(oc) and abs(c1 - c0) <= 5e-4 * abs(c0), (c1, c0, oc)
    upd = np.linalg.norm(oh[16] - host[16])
    assert np.linalg.norm(x1 - oh[16]) <= 2e-2 * upd and np.linalg.norm(x1 - x0) <= 2e-2 * upd


def test_two_solves_on_one_plan(built):
    """the second solve replays the first one's graphs: the loop kernel's sequence word and arrival counter carry over"""
    W, H = 200, 150
    host = workloads.shape_from_shading(W, H)
    s = _plan(W, H, "LMGPU", {})
    sp = {"nIterations": 5, "lIterations": 10}
    dev = workloads.to_device(workloads.clone_params(host))
    c1 = s.solve(sp, dev)
    dev2 = workloads.to_device(workloads.clone_params(host))
    c2 = s.solve(sp, dev2)
    assert c1 == c2, (c1, c2)          # same inputs, same launches, deterministic reductions
    assert np.array_equal(dev[16].cpu().numpy(), dev2[16].cpu().numpy())
    s.close()


@pytest.mark.parametrize("W,H", [(45, 37), (640, 480)])
def test_model_cost_from_loop_kernel(built, W, H):
    """computeModelCost folded into the loop kernel's tail against the stand-alone kernel, after one and after three LM iterations"""
    host = workloads.shape_from_shading(W, H)
    vals = {}
    for tag, env in (("fused", {}), ("head_separate", {"OPTB_SFS_NO_HEAD_IN_LOOP": "1"}), ("separate", {"OPTB_SFS_NO_FUSED_MODEL": "1"}), ("per_step", PER_STEP)):
        dev = workloads.to_device(workloads.clone_params(host))
        s = _plan(W, H, "LMGPU", env)
        s.set_solver_parameters({"nIterations": 5, "lIterations": 10})
        s.init(dev)
        out = []
        for _ in range(3):
            s.step(dev)
            out.append((s.read_scalar("modelCost"), s.current_cost(), s.read_scalar("trust_region_radius")))
        vals[tag] = out
        s.close()
    for tag in ("head_separate", "separate", "per_step"):
        for (m1, c1, t1), (m0, c0, t0) in zip(vals["fused"], vals[tag]):
            assert abs(m1 - m0) <= 1e-4 * abs(m0) and abs(c1 - c0) <= 2e-4 * abs(c0) and abs(t1 - t0) <= 1e-3 * t0, (tag, vals["fused"], vals[tag])
