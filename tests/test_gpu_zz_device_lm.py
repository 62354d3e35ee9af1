"""GPU: Levenberg-Marquardt with the accept / reject / trust-region decision ON THE DEVICE (OPTB_DEVICE_LM=1; SURVEY.md 8(f) rank 2:
"cost, model cost, accept/reject, radius update ... without host syncs").  Opt_ProblemSolve then enqueues every nonlinear iteration after
the first back to back -- loop kernel, precompute + cost, k_lmDecide, revert-if-rejected, precompute-if-rejected -- and waits once per chunk of eight iterations.  The
kernels are the default path's (they only gained a gate argument) and the decision is the arithmetic of Plan::decide (lm_decide.cuh,
checked bit for bit against the reference's lines on the CPU: tests/test_lm_decide_host.py), so a solve must end where the default solve
ends: same number of iterations, same radius, same cost, same unknowns.

Written after this round's GPU budget was spent: never run on hardware, hence the non-strict xfail mark -- an XPASS in the driver's run is
the evidence, a failure does not hide the verified suite.  The mode is opt-in and touches no default launch."""
import contextlib
import os
import subprocess
import sys

import numpy as np
import pytest

from opt_b200 import workloads

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="added after the round's GPU budget was spent; first hardware run is the driver's")]


@pytest.fixture
def child(request):
    """These tests have never run on hardware, and a CUDA failure inside libOpt.so ends the process (printf + exit, like the reference,
    util.t:739-768): so each one runs in its own python.  In the parent this fixture runs the very same test id in a child (xfail marks
    off) and hands its result to the body, which only checks the exit code; in the child it returns None and the body runs."""
    if os.environ.get("OPTB_TEST_CHILD") == "1":
        return None
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "--runxfail", "-m", "gpu", "-p", "no:cacheprovider", request.node.nodeid],
                          cwd=root, env=dict(os.environ, OPTB_TEST_CHILD="1"), capture_output=True, text=True, timeout=900)


@contextlib.contextmanager
def _env(**env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        yield
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)


def _solve(W, H, sp, device_lm, host, twice=False):
    from opt_b200 import OptSolver, spec_path
    s = OptSolver([W, H], spec_path("shape_from_shading"), "LMGPU")
    out = []
    with _env(OPTB_DEVICE_LM="1" if device_lm else "0"):
        for _ in range(2 if twice else 1):
            dev = workloads.to_device(workloads.clone_params(host))
            c = s.solve(sp, dev)
            out.append((c, dev[16].cpu().numpy().copy(), int(s.read_scalar("nIter")), s.read_scalar("trust_region_radius"), int(s.read_scalar("deviceIterations"))))
    s.close()
    return out


CASES = {
    "plain": {"nIterations": 8, "lIterations": 10},
    # steps of quality below 0.9 are rejected: reverts, shrinking radius, precompute redone (:1147-1156)
    "rejections": {"nIterations": 10, "lIterations": 10, "min_relative_decrease": 0.9},
    # the solve ends on the function tolerance well before nIterations (:1126-1130): the remaining enqueued iterations must do nothing
    "function_tolerance": {"nIterations": 12, "lIterations": 10, "function_tolerance": 0.05},
    # every step rejected until the radius falls under the minimum (:1152-1155)
    "radius_too_small": {"nIterations": 12, "lIterations": 10, "min_relative_decrease": 1e9, "min_trust_region_radius": 1e2},
}


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("W,H", [(64, 48), (640, 480)])
def test_device_decision_ends_where_the_host_decision_ends(built, W, H, case, child):
    if child is not None:
        assert child.returncode == 0, (child.stdout + child.stderr)[-3000:]
        return
    host = workloads.shape_from_shading(W, H)
    sp = CASES[case]
    (c0, x0, n0, r0, d0), = _solve(W, H, sp, False, host)
    (c1, x1, n1, r1, d1), = _solve(W, H, sp, True, host)
    assert d0 == 0
    if n0 > 1:
        assert d1 == n1 - 1, (d1, n1, n0)      # everything after the first iteration ran without the host
    assert n1 == n0, (n1, n0)
    assert abs(r1 - r0) <= 1e-6 * abs(r0), (r1, r0)
    assert abs(c1 - c0) <= 1e-6 * abs(c0), (c1, c0)
    upd = np.linalg.norm(x0 - host[16])
    assert np.linalg.norm(x1 - x0) <= 1e-5 * max(upd, 1e-6 * np.linalg.norm(x0))
    if case == "rejections":
        assert r0 < 1e4          # at least one rejection shrank the radius from its default (the case exercises the revert)
    if case == "function_tolerance":
        assert n0 < sp["nIterations"]
    if case == "radius_too_small":
        assert n0 < sp["nIterations"] and np.array_equal(x0, host[16]) and np.array_equal(x1, host[16])


def test_device_decision_against_the_oracle_and_twice_on_one_plan(built, child):
    if child is not None:
        assert child.returncode == 0, (child.stdout + child.stderr)[-3000:]
        return
    from oracle.oracle import Oracle
    W, H = 200, 150
    host = workloads.shape_from_shading(W, H)
    sp = {"nIterations": 8, "lIterations": 10}
    runs = _solve(W, H, sp, True, host, twice=True)
    assert runs[0][0] == runs[1][0] and np.array_equal(runs[0][1], runs[1][1])      # the second solve replays the first one's graphs
    o = Oracle("shape_from_shading", "LMGPU", [W, H])
    for k, v in sp.items():
        o.set_solver_parameter(k, v)
    oh = workloads.clone_params(host)
    oc = o.solve(oh)
    assert abs(runs[0][0] - oc) <= 5e-4 * abs(oc), (runs[0][0], oc)
    assert np.linalg.norm(runs[0][1] - oh[16]) <= 2e-2 * np.linalg.norm(oh[16] - host[16])
