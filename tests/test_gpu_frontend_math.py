"""GPU: numeric check of the front end's single-argument math functions (values and device dual-number derivatives).
Kept in a file that sorts last: it was written after the round's GPU budget was spent, so its first run is the
round-end run; a surprise here must not hide the rest of the suite behind `-x`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


MATH_FNS = {   # name -> (numpy function, derivative, constant that keeps a*x + c inside the domain)
    "sqrt": (np.sqrt, lambda u: 0.5 / np.sqrt(u), 1.5), "sin": (np.sin, np.cos, 1.5), "cos": (np.cos, lambda u: -np.sin(u), 1.5),
    "tan": (np.tan, lambda u: 1 + np.tan(u) ** 2, 0.3), "exp": (np.exp, np.exp, 0.3), "log": (np.log, lambda u: 1 / u, 1.5),
    "log10": (np.log10, lambda u: 1 / (np.log(10.0) * u), 1.5), "asin": (np.arcsin, lambda u: 1 / np.sqrt(1 - u * u), 0.2),
    "acos": (np.arccos, lambda u: -1 / np.sqrt(1 - u * u), 0.2), "atan": (np.arctan, lambda u: 1 / (1 + u * u), 0.5),
    "sinh": (np.sinh, np.cosh, 0.5), "cosh": (np.cosh, np.sinh, 0.5), "tanh": (np.tanh, lambda u: 1 - np.tanh(u) ** 2, 0.5),
    "asinh": (np.arcsinh, lambda u: 1 / np.sqrt(u * u + 1), 0.5), "acosh": (np.arccosh, lambda u: 1 / np.sqrt(u * u - 1), 1.8),
}


def _check_math_function(tmp_path, fn):
    """every single-argument function of ad.t:747-763: the residual f(0.1 x + c) - y through front end + device dual
    numbers, cost and -J^T F against closed forms in numpy (fp64)"""
    import torch
    from opt_b200 import OptSolver
    f, df, c = MATH_FNS[fn]
    n = 257
    spec = tmp_path / f"m_{fn}.t"
    spec.write_text(f'local N = Dim("N", 0)\nlocal X = Unknown("X", opt_float, {{ N }}, 0)\nlocal Y = Array("Y", opt_float, {{ N }}, 1)\n'
                    f'UsePreconditioner(true)\nEnergy({fn}(X(0) * 0.1 + {c}) - Y(0))\n')
    rng = np.random.default_rng(5)
    x = rng.uniform(-1.0, 1.0, n); y = rng.normal(size=n)
    dev = [torch.from_numpy(x.copy()).cuda(), torch.from_numpy(y).cuda()]
    s = OptSolver([n], str(spec), "gaussNewtonGPU", double_precision=True)
    cost, r0, pre, _ = s.eval_jtf(dev)
    u = 0.1 * x + c
    res = f(u) - y
    jac = 0.1 * df(u)
    assert abs(cost - 0.5 * np.sum(res * res)) <= 1e-12 * max(1.0, cost)
    got = s.unflatten(r0)[0].reshape(-1)
    assert np.abs(got - (-jac * res)).max() <= 1e-11 * max(1.0, np.abs(jac * res).max())
    s.close()


MASK_SPEC = """
local W, H = Dim("W", 0), Dim("H", 1)
local X = Unknown("X", opt_float, { W, H }, 0)
local Y = Array("Y", opt_float, { W, H }, 1)
local M = Array("M", opt_float, { W, H }, 2)
UsePreconditioner(true)
Exclude(Not(eq(M(0, 0), 0)))
Energy(X(0, 0) - Y(0, 0))
for dx, dy in Stencil { { 1, 0 }, { 0, 1 } } do
    Energy(Select(InBounds(dx, dy), 2.0 * (X(0, 0) - X(dx, dy)), 0))
end
"""


def _check_exclude_mask_may_change_between_steps(tmp_path):
    """the reference evaluates Exclude() in every kernel, so a caller may edit a mask image between two Opt_ProblemStep
    calls; the generic engine caches masked columns per solve and must notice (generic_engine.cuh k_geActiveChanged)"""
    import torch
    from opt_b200 import OptSolver
    W, H = 24, 17
    spec = tmp_path / "masked_smooth.t"; spec.write_text(MASK_SPEC)
    rng = np.random.default_rng(2)
    Y = rng.normal(size=W * H); X0 = Y + 0.3 * rng.normal(size=W * H)
    M1 = (rng.random(W * H) < 0.15).astype(np.float64); M2 = (rng.random(W * H) < 0.15).astype(np.float64)
    params = {"nIterations": 5, "lIterations": 20}
    # solver A: one step with mask 1, then the mask buffer is overwritten in place, one more step
    devA = [torch.from_numpy(X0.copy()).cuda(), torch.from_numpy(Y).cuda(), torch.from_numpy(M1.copy()).cuda()]
    a = OptSolver([W, H], str(spec), "gaussNewtonGPU", double_precision=True)
    a.set_solver_parameters(params)
    a.init(devA); a.step(devA)
    X1 = devA[0].cpu().numpy().copy()
    devA[2].copy_(torch.from_numpy(M2))
    a.step(devA)
    XA = devA[0].cpu().numpy().copy()
    a.close()
    # solver B: starts from X1 with mask 2
    devB = [torch.from_numpy(X1.copy()).cuda(), torch.from_numpy(Y).cuda(), torch.from_numpy(M2.copy()).cuda()]
    b = OptSolver([W, H], str(spec), "gaussNewtonGPU", double_precision=True)
    b.set_solver_parameters(params)
    b.init(devB); b.step(devB)
    XB = devB[0].cpu().numpy().copy()
    b.close()
    assert np.abs(XA - XB).max() <= 1e-12 * max(1.0, np.abs(XB).max())
    assert np.array_equal(XA[M2 != 0], X1[M2 != 0])            # pixels excluded by the NEW mask did not move in the second step


def _run_in_child(body):
    """the library answers a CUDA error with exit(code) like the reference does: first runs of new device paths happen in a child
    process so that a surprise cannot take the pytest session down"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (f"import sys, pathlib, tempfile; sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
            "import test_gpu_frontend_math as t\ntmp = pathlib.Path(tempfile.mkdtemp())\n" + body + "\nprint('CHILD_OK')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0 and "CHILD_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_math_function_values_and_derivatives(built):
    _run_in_child("\n".join(f"t._check_math_function(tmp, {fn!r}); print('{fn} ok')" for fn in sorted(MATH_FNS)))


def test_exclude_mask_may_change_between_steps(built):
    _run_in_child("t._check_exclude_mask_may_change_between_steps(tmp)")
