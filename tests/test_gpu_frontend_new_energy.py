"""GPU: an energy that exists in NO registry goes through the general front end (interpreter -> generated CUDA ->
NVRTC -> generic engine) and is checked against an independent numpy restatement with explicit finite-difference
gradients.  Also the reference's own example specs without a built-in kernel set (volumetric / cotangent / embedded
mesh energies) must plan and reduce their cost."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SPEC = """
local W, H = Dim("W", 0), Dim("H", 1)
local X = Unknown("X", opt_float2, { W, H }, 0)
local Y = Array("Y", opt_float2, { W, H }, 1)
local M = Array("M", opt_float, { W, H }, 2)
local lambda = Param("lambda", float, 3)
UsePreconditioner(true)
Exclude(Not(eq(M(0, 0), 0)))
Energy(X(0, 0) - Y(0, 0))
for dx, dy in Stencil { { 1, 0 }, { 0, 1 } } do
    local g = X(0, 0) - X(dx, dy)
    local w = 1.0 / sqrt(sqrt(g:dot(g) + 0.01))
    Energy(Select(InBounds(dx, dy), lambda * w * g, 0))
end
"""


def _cost(X, Y, M, lam, W, H):
    X = X.reshape(H, W, 2); Y = Y.reshape(H, W, 2); act = (M.reshape(H, W) == 0)
    c = 0.5 * (((X - Y) ** 2).sum(-1) * act).sum()
    for dx, dy in ((1, 0), (0, 1)):
        g = X[:H - dy, :W - dx] - X[dy:, dx:]
        w = 1.0 / np.sqrt(np.sqrt((g * g).sum(-1) + 0.01))
        c += 0.5 * (((lam * w[..., None] * g) ** 2).sum(-1) * act[:H - dy, :W - dx]).sum()
    return c


def test_new_energy_matches_numpy(built, tmp_path):
    import torch
    from opt_b200 import OptSolver
    W, H = 20, 14
    spec = tmp_path / "robust_smooth.t"; spec.write_text(SPEC)
    rng = np.random.default_rng(0)
    Y = rng.normal(size=(W * H, 2)); X0 = Y + 0.1 * rng.normal(size=Y.shape); M = (rng.random(W * H) < 0.1).astype(np.float64); lam = 0.7
    dev = [torch.from_numpy(X0.copy()).cuda(), torch.from_numpy(Y).cuda(), torch.from_numpy(M).cuda(), lam]
    s = OptSolver([W, H], str(spec), "gaussNewtonGPU", double_precision=True)
    cost, r0, pre, _ = s.eval_jtf(dev)
    assert abs(cost - _cost(X0, Y, M, np.float32(lam), W, H)) < 1e-9 * cost
    # J^T F of included unknowns == finite-difference gradient of the cost w.r.t. included unknowns, with excluded pixels frozen
    # (for this energy rows centred at excluded pixels also feed their neighbours: compare only pixels whose upper/left neighbours are included)
    g = -s.unflatten(r0)[0].reshape(H, W, 2)
    act = (M.reshape(H, W) == 0)
    ok = act.copy(); ok[:, 1:] &= act[:, :-1]; ok[1:, :] &= act[:-1, :]
    checked = 0
    for (y, x) in zip(*np.nonzero(ok)):
        if checked >= 25:
            break
        for c in range(2):
            Xp = X0.copy().reshape(H, W, 2); Xm = Xp.copy(); h = 1e-6
            Xp[y, x, c] += h; Xm[y, x, c] -= h
            fd = (_cost(Xp.reshape(-1, 2), Y, M, np.float32(lam), W, H) - _cost(Xm.reshape(-1, 2), Y, M, np.float32(lam), W, H)) / (2 * h)
            assert abs(fd - g[y, x, c]) < 1e-5 * max(1.0, abs(fd)), (y, x, c, fd, g[y, x, c])
        checked += 1
    c0 = cost
    c1 = s.solve({"nIterations": 5, "lIterations": 30}, dev)
    assert c1 < c0
    Xf = dev[0].cpu().numpy()
    assert np.array_equal(Xf.reshape(H, W, 2)[~act], X0.reshape(H, W, 2)[~act])     # excluded unknowns untouched
    assert abs(c1 - _cost(Xf, Y, M, np.float32(lam), W, H)) < 1e-9 * max(c1, 1e-12)


@pytest.mark.parametrize("rel", ["examples/volumetric_mesh_deformation/volumetric_mesh_deformation.t", "examples/cotangent_mesh_smoothing/cotangent_mesh_smoothing.t",
                                 "examples/embedded_mesh_deformation/embedded_mesh_deformation.t"])
def test_reference_specs_without_builtin_kernels(built, rel):
    import torch
    from opt_b200 import OptSolver
    path = os.path.join("/root/reference", rel)
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted on this machine")
    rng = np.random.default_rng(1)
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
    if "volumetric" in rel:
        n = 6; N = n ** 3
        g = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(N, 3)[:, ::-1].astype(np.float32)
        cons = np.full((N, 3), -np.inf, np.float32); cons[0] = g[0]; cons[-1] = g[-1] + [1.0, 0.5, 0.0]
        dims, params = [n, n, n], [f32(g), f32(np.zeros((N, 3))), f32(g), f32(cons), 2.0, 1.0]
    elif "cotangent" in rel:
        n = 40; X = rng.normal(size=(n, 3))
        v0 = np.arange(n, dtype=np.int32); v1 = (v0 + 1) % n; v2 = (v0 + 2) % n; v3 = (v0 + 3) % n
        I = lambda a: torch.from_numpy(a.astype(np.int32)).cuda()
        dims, params = [n], [1.0, 0.5, f32(X), f32(X + 0.05 * rng.normal(size=X.shape)), int(n), I(v0), I(v1), I(v2), I(v3)]
    else:
        n = 30; P = rng.normal(size=(n, 3)).astype(np.float32)
        R = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (n, 1)) + 0.05 * rng.normal(size=(n, 9)).astype(np.float32)
        cons = np.full((n, 3), -np.inf, np.float32); cons[:5] = P[:5] + 0.3
        v0 = np.repeat(np.arange(n), 2).astype(np.int32); v1 = ((v0 + np.tile([1, 5], n)) % n).astype(np.int32)
        I = lambda a: torch.from_numpy(a).cuda()
        dims, params = [n], [2.0, 1.0, 1.5, f32(P), f32(R), f32(P), f32(cons), int(len(v0)), I(v0), I(v1)]
    s = OptSolver(dims, path, "gaussNewtonGPU")
    s.set_solver_parameters({"nIterations": 4, "lIterations": 20})
    s.init(params); c0 = s.current_cost()
    while s.step(params):
        pass
    assert np.isfinite(s.current_cost()) and s.current_cost() < c0


def test_optical_flow_sampled_image(built):
    """SampledImage (bilinear read at unknown-dependent coordinates, partials from derivative images): cost equals a numpy
    restatement, a GN solve with exact derivative images reduces it"""
    import torch
    from opt_b200 import OptSolver, spec_path
    W, H = 48, 36
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    f = lambda x, y: np.sin(0.21 * x) * np.cos(0.17 * y) + 0.3 * np.sin(0.065 * x + 0.11 * y)
    fx = lambda x, y: 0.21 * np.cos(0.21 * x) * np.cos(0.17 * y) + 0.3 * 0.065 * np.cos(0.065 * x + 0.11 * y)
    fy = lambda x, y: -0.17 * np.sin(0.21 * x) * np.sin(0.17 * y) + 0.3 * 0.11 * np.cos(0.065 * x + 0.11 * y)
    I_hat, I_dx, I_dy = f(xs, ys), fx(xs, ys), fy(xs, ys)
    I = f(xs + 0.6, ys - 0.4)                                   # the scene moved by (0.6, -0.4)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dev = [1.0, 0.3, t(np.zeros((W * H, 2))), t(I.reshape(-1)), t(I_hat.reshape(-1)), t(I_dx.reshape(-1)), t(I_dy.reshape(-1))]
    s = OptSolver([W, H], spec_path("optical_flow"), "gaussNewtonGPU", double_precision=True)
    s.set_solver_parameters({"nIterations": 6, "lIterations": 40})
    s.init(dev)
    c0 = s.current_cost()
    assert abs(c0 - 0.5 * ((I - I_hat) ** 2).sum()) < 1e-9 * c0      # zero flow: the sample is the pixel itself, regulariser is zero
    while s.step(dev):
        pass
    c1 = s.current_cost()
    flow = dev[2].cpu().numpy().reshape(H, W, 2)[8:-8, 8:-8].mean(axis=(0, 1))
    # un-preconditioned PCG on this smooth image converges slowly: require a real decrease and a flow that points the right way
    assert c1 < 0.95 * c0 and flow[0] > 0.02 and flow[1] < -0.01, (c0, c1, flow)

