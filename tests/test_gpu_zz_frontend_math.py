"""GPU: numeric check of the front end's single-argument math functions (values and device dual-number derivatives).
Kept in a file that sorts last: it was written after the round's GPU budget was spent, so its first run is the
round-end run; a surprise here must not hide the rest of the suite behind `-x`."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first run on hardware is the round-end run (written after the round's GPU budget was "
                                                     "spent); remove this mark once it has passed on a B200")]


MATH_FNS = {   # name -> (numpy function, derivative, constant that keeps a*x + c inside the domain)
    "sqrt": (np.sqrt, lambda u: 0.5 / np.sqrt(u), 1.5), "sin": (np.sin, np.cos, 1.5), "cos": (np.cos, lambda u: -np.sin(u), 1.5),
    "tan": (np.tan, lambda u: 1 + np.tan(u) ** 2, 0.3), "exp": (np.exp, np.exp, 0.3), "log": (np.log, lambda u: 1 / u, 1.5),
    "log10": (np.log10, lambda u: 1 / (np.log(10.0) * u), 1.5), "asin": (np.arcsin, lambda u: 1 / np.sqrt(1 - u * u), 0.2),
    "acos": (np.arccos, lambda u: -1 / np.sqrt(1 - u * u), 0.2), "atan": (np.arctan, lambda u: 1 / (1 + u * u), 0.5),
    "sinh": (np.sinh, np.cosh, 0.5), "cosh": (np.cosh, np.sinh, 0.5), "tanh": (np.tanh, lambda u: 1 - np.tanh(u) ** 2, 0.5),
    "asinh": (np.arcsinh, lambda u: 1 / np.sqrt(u * u + 1), 0.5), "acosh": (np.arccosh, lambda u: 1 / np.sqrt(u * u - 1), 1.8),
}


@pytest.mark.parametrize("fn", sorted(MATH_FNS))
def test_math_function_values_and_derivatives(built, tmp_path, fn):
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
