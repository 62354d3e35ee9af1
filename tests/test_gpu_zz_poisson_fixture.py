"""GPU parity on the reference app's own poisson_image_editing run (BASELINE.json configs[0]; examples/poisson_image_editing/src/main.cpp:
poisson0.png + poisson1.png + poisson_mask.png, 448x320, 22 828 solved pixels, nIterations = 1, lIterations = 100) through the C ABI against
the CPU oracle -- whose poisson energy is itself pinned to the reference's hand-written equations on the same images
(tests/test_oracle.py::test_oracle_poisson_on_the_shipped_images_vs_reference_hand_equations).

Written after this round's GPU budget was spent: it has never run on hardware, hence the non-strict xfail mark (the same energy passes
at 21x17 in test_gpu_generic_energies.py; what is new here is the size, the real image statistics and the app's schedule).  An XPASS in the
driver's run is the evidence; a failure does not hide the verified suite."""
import os
import subprocess
import sys

import numpy as np
import pytest

from opt_b200 import examples, workloads

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


F = examples.FIXTURES


def _problem():
    return examples.poisson_from_files(os.path.join(F, "poisson0.png"), os.path.join(F, "poisson1.png"), os.path.join(F, "poisson_mask.png"))


def _flat(s, vec):
    parts, o = [], 0
    for off, el, ch in s.unknown_layout():
        parts.append(vec[o:o + el * ch].reshape(el, ch)); o += el * ch
    return s.flatten(parts)


@pytest.mark.parametrize("route", ["registry", "frontend"])
def test_poisson_shipped_images_kernel_level(built, route, child):
    if child is not None:
        assert child.returncode == 0, (child.stdout + child.stderr)[-3000:]
        return
    from opt_b200 import OptSolver, spec_path
    from oracle.oracle import Oracle
    host, (W, H) = _problem()
    os.environ.pop("OPTB_FRONTEND", None)
    if route == "frontend":
        os.environ["OPTB_FRONTEND"] = "1"
    try:
        s = OptSolver([W, H], spec_path("poisson_image_editing"), "gaussNewtonGPU")
    finally:
        os.environ.pop("OPTB_FRONTEND", None)
    o = Oracle("poisson_image_editing", "gaussNewtonGPU", [W, H])
    dev = workloads.to_device(host)
    cost, r0, pre_inv, rdp = s.eval_jtf(dev)
    ocost, g, diag = o.eval_jtf(host)
    assert abs(cost - ocost) <= 2e-5 * abs(ocost), (cost, ocost)
    gf = _flat(s, g)
    assert np.abs(r0 + gf).max() <= 3e-5 * np.abs(gf).max()
    x = np.random.default_rng(9).normal(size=len(g)).astype(np.float32)
    out, dot = s.apply_jtj(dev, _flat(s, x))
    ref = _flat(s, o.apply_jtj(host, x))
    assert np.abs(out - ref).max() <= 3e-5 * max(1.0, np.abs(ref).max())
    s.close()


def test_poisson_shipped_images_app_schedule(built, child):
    """nIterations = 1, lIterations = 100 as the app runs it; fp32 on both sides, different summation orders"""
    if child is not None:
        assert child.returncode == 0, (child.stdout + child.stderr)[-3000:]
        return
    from opt_b200 import OptSolver, spec_path
    from oracle.oracle import Oracle
    host, (W, H) = _problem()
    s = OptSolver([W, H], spec_path("poisson_image_editing"), "gaussNewtonGPU")
    o = Oracle("poisson_image_editing", "gaussNewtonGPU", [W, H])
    sp = {"nIterations": 1, "lIterations": 100}
    for k, v in sp.items():
        o.set_solver_parameter(k, v)
    ref = workloads.clone_params(host); dev = workloads.to_device(workloads.clone_params(host))
    oc = o.solve(ref); c = s.solve(sp, dev)
    assert abs(c - oc) <= 1e-3 * abs(oc), (c, oc)
    a, b = dev[0].cpu().numpy(), ref[0]
    inc = host[2] == 0
    assert np.array_equal(a[~inc], host[0][~inc])                      # excluded pixels untouched
    assert np.linalg.norm(a - b) <= 2e-3 * np.linalg.norm(b - host[0]), (float(np.linalg.norm(a - b)), float(np.linalg.norm(b - host[0])))
    s.close()
