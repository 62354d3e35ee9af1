"""GPU: the opt-in 2-D tile variant of the state-on-chip PCG loop (OPTB_RESIDENT_TILE=WxH, energy_image_warping.cu
k_iwPcgLoopTile) against the default loops.  Sorts last on purpose: the kernel was written after the round's GPU budget
was spent, so its first run on hardware is the round-end run; the validated default path does not depend on it."""
import contextlib
import os

import numpy as np
import pytest

from opt_b200 import workloads

pytestmark = pytest.mark.gpu


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


def _state_after(W, H, k, env, dp=False):
    from opt_b200 import OptSolver, spec_path
    host = workloads.image_warping(W, H, mask="random", seed=3, perturb=0.3)
    if dp:
        host = [a.astype(np.float64) if isinstance(a, np.ndarray) and a.dtype == np.float32 else a for a in host]
    dev = workloads.to_device(host)
    with _env(**env):
        s = OptSolver([W, H], spec_path("image_warping"), "gaussNewtonGPU", double_precision=dp)
    s.set_solver_parameter("nIterations", 1); s.set_solver_parameter("lIterations", 50)
    s.set_max_linear_iterations(k)
    s.init(dev); s.step(dev)
    out = {n: s.read_vector(n).copy() for n in ("delta", "r", "z", "p", "Ap_X")}
    launches = s.launches_last_step()
    s.close()
    return out, launches


def _compare(W, H, tile):
    k = 6
    ref, n_ref = _state_after(W, H, k, {"OPTB_PERSISTENT": "0"})
    got, n_got = _state_after(W, H, k, {"OPTB_PERSISTENT": "1", "OPTB_RESIDENT_TILE": tile})
    assert n_got < n_ref          # one launch for the whole loop instead of two per iteration
    for name in ref:
        scale = max(1e-30, np.abs(ref[name]).max())
        assert np.abs(got[name] - ref[name]).max() <= 1e-3 * scale, name


CASES = [(48, 40, "16x8"), (37, 29, "16x8"), (64, 64, "64x4"), (50, 33, "7x5")]


def test_tile_variant_matches_per_step_loop(built):
    """runs in ONE child process: the library answers a CUDA error with exit(code) like the reference does, and an unproven kernel
    must not be able to take the test session down with it"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (f"import sys; sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r}); import test_gpu_resident_tile as t\n"
            + "\n".join(f"t._compare({W}, {H}, {tile!r}); print('case {W}x{H} tile {tile} ok')" for W, H, tile in CASES) + "\nprint('TILE_OK')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0 and "TILE_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])
