"""Synthetic problem instances for the benchmark configurations (SURVEY.md section 8d,
BASELINE.json configs).  Pure numpy; each generator returns the positional parameter list of the
corresponding .t spec with HOST numpy arrays (float Params as python floats, graph edge counts as
ints).  `to_device` turns such a list into the CUDA-tensor form OptSolver takes."""
import numpy as np


def image_warping(W, H, dtype=np.float32, seed=0, mask="none", n_constraints=9, w_fit=100.0, w_reg=0.01,
                  perturb=0.0):
    """image_warping.t inputs, built the way the reference app does (examples/image_warping/src/
    CombinedSolver.h:161-207, main.cpp:98-108): UrShape = pixel grid, Offset0 = UrShape, Angle0 = 0,
    border pixels pinned to themselves, a few interior constraints displaced, weights passed as
    square roots (CombinedSolver.h:126-130).
    mask: "none" (all active) | "random" (5% interior pixels excluded) | "blob" (a disc excluded).
    perturb > 0 adds noise to Offset/Angle so that every Jacobian entry is exercised."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    ur = np.stack([xs, ys], -1).astype(dtype).reshape(-1, 2).copy()
    off = ur.copy()
    ang = np.zeros(W * H, dtype)
    cons = -np.ones((W * H, 2), dtype)
    m = np.zeros(W * H, dtype)
    border = ((xs == 0) | (ys == 0) | (xs == W - 1) | (ys == H - 1)).reshape(-1)
    if mask == "random":
        m[(rng.random(W * H) < 0.05) & ~border] = 1
    elif mask == "blob":
        cx, cy, rad = 0.3 * W, 0.6 * H, 0.12 * min(W, H)
        m[(((xs - cx) ** 2 + (ys - cy) ** 2) < rad * rad).reshape(-1) & ~border] = 1
    cons[border] = ur[border]
    for _ in range(n_constraints):
        cx, cy = int(rng.integers(W // 8, W - W // 8)), int(rng.integers(H // 8, H - H // 8))
        if m[cy * W + cx] == 0:
            s = 0.04 * min(W, H)
            cons[cy * W + cx] = [cx + rng.uniform(-s, s), cy + rng.uniform(-s, s)]
    if perturb > 0:
        off += rng.normal(0, perturb, off.shape).astype(dtype)
        ang += rng.normal(0, 0.3 * perturb, ang.shape).astype(dtype)
    return [off, ang, ur, cons, m, float(np.sqrt(w_fit)), float(np.sqrt(w_reg))]


def clone_params(params):
    return [p.copy() if isinstance(p, np.ndarray) else p for p in params]


def to_device(params, device="cuda"):
    import torch
    out = []
    for p in params:
        out.append(torch.from_numpy(np.ascontiguousarray(p)).to(device) if isinstance(p, np.ndarray) else p)
    return out


def to_host(params):
    return [p.detach().cpu().numpy() if hasattr(p, "detach") else p for p in params]


def grid_mesh_edges(nx, ny):
    """directed edge lists (v0, v1) of a triangulated nx x ny grid, built like the reference's
    createGraphFromNeighborLists (examples/shared/OptGraph.h:64-75): for every vertex, one directed edge
    to each neighbour (valence <= 6)."""
    idx = np.arange(nx * ny).reshape(ny, nx)
    pairs = []
    for dy, dx in ((0, 1), (1, 0), (1, 1)):
        a = idx[:ny - dy, :nx - dx].reshape(-1); b = idx[dy:, dx:].reshape(-1)
        pairs.append(np.stack([a, b], 1)); pairs.append(np.stack([b, a], 1))
    e = np.concatenate(pairs, 0)
    e = e[np.lexsort((e[:, 1], e[:, 0]))]           # grouped by head vertex, as the neighbour-list walk produces
    return e[:, 0].astype(np.int32).copy(), e[:, 1].astype(np.int32).copy()


def arap_mesh(nx, ny, dtype=np.float32, seed=230948, n_handles=32, w_fit=4.0, w_reg=1.0, perturb=0.0):
    """arap_mesh_deformation.t inputs on a jittered triangulated grid: rest positions =
    grid + N(0, 0.01) jitter, Offset0 = rest, Angle0 = 0, `n_handles` vertices pinned/moved, all others
    unconstrained (Constraints.x = -inf as the reference app writes, CombinedSolver.h)."""
    rng = np.random.default_rng(seed)
    N = nx * ny
    xs, ys = np.meshgrid(np.arange(nx), np.arange(ny))
    ur = np.stack([xs, ys, np.zeros_like(xs)], -1).reshape(N, 3).astype(np.float64)
    ur += rng.normal(0, 0.01, ur.shape)
    ur = ur.astype(dtype)
    off = ur.copy()
    ang = np.zeros((N, 3), dtype)
    cons = np.full((N, 3), -np.inf, dtype)
    hs = rng.choice(N, size=min(n_handles, N), replace=False)
    for k, h in enumerate(hs):
        cons[h] = ur[h] + (rng.normal(0, 0.05 * max(nx, ny), 3) if k % 2 else 0.0)
    if perturb > 0:
        off += rng.normal(0, perturb, off.shape).astype(dtype)
        ang += rng.normal(0, 0.3 * perturb, ang.shape).astype(dtype)
    v0, v1 = grid_mesh_edges(nx, ny)
    return [float(np.sqrt(w_fit)), float(np.sqrt(w_reg)), off, ang, ur, cons, int(len(v0)), v0, v1]


def poisson(W, H, dtype=np.float32, seed=0):
    """poisson_image_editing.t inputs: X0 = base image, T = inserted image, M = 0 inside the blend region."""
    rng = np.random.default_rng(seed)
    X = rng.random((W * H, 4)).astype(dtype)
    T = rng.random((W * H, 4)).astype(dtype)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    inside = ((xs > W * 0.2) & (xs < W * 0.8) & (ys > H * 0.25) & (ys < H * 0.75)).reshape(-1)
    M = np.where(inside, 0, 1).astype(dtype)
    return [X, T, M]


def laplacian(W, H, dtype=np.float32, seed=0):
    rng = np.random.default_rng(seed)
    return [np.zeros(W * H, dtype), rng.random(W * H).astype(dtype)]


def curve_fitting(n=512, dtype=np.float64, a=100.0, b=102.0, start=(99.7, 101.6)):
    """tests/minimal_graph_only/main.cpp:43-61: noiseless samples of y = a cos(bx) + b sin(ax)"""
    x = np.arange(n) * 2.0 * 3.141592653589 / n
    y = a * np.cos(b * x) + b * np.sin(a * x)
    data = np.stack([x, y], 1).astype(dtype).copy()
    unk = np.array([[np.float32(start[0]), np.float32(start[1])]], dtype=dtype)
    d = np.arange(n, dtype=np.int32); p = np.zeros(n, dtype=np.int32)
    return [unk, data, int(n), d, p]     # dims = [n, 1]


def shape_from_shading(W, H, dtype=np.float32, seed=0):
    """shape_from_shading.t inputs, synthetic: a smooth depth map with an invalid (D_i <= 0) corner region,
    intensity rendered from a noisy copy, camera and SH lighting like the shipped parameter file
    (w_p=100 w_s=100 w_g=1, f~574, principal point at the image centre)."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    depth = 1.0 + 0.15 * np.sin(xs / W * 3.1) * np.cos(ys / H * 2.3) + 0.001 * rng.normal(size=xs.shape)
    D = depth.copy()
    D[(xs < W * 0.15) & (ys < H * 0.2)] = -10000.0          # loader clamps -inf to -10000 (SimpleBuffer.cpp:30-40)
    X = (D + 0.002 * rng.normal(size=D.shape)).astype(dtype).reshape(-1).copy()
    Im = (0.5 + 0.3 * np.sin(xs * 0.11) * np.cos(ys * 0.07)).astype(dtype).reshape(-1).copy()
    eR = (rng.random(W * H) < 0.9).astype(np.uint8); eC = (rng.random(W * H) < 0.9).astype(np.uint8)
    L = [0.6, -0.1, 0.5, 0.2, 0.05, -0.03, 0.08, 0.02, -0.04]
    f = 574.05 * W / 640.0
    return [100.0, 100.0, 1.0, float(f), float(f), W / 2.0, H / 2.0] + [float(v) for v in L] + \
           [X, D.astype(dtype).reshape(-1).copy(), Im, eR, eC]


def robust_nonrigid(nx, ny, dtype=np.float64, seed=1, w_fit=10.0, w_reg=1.0, perturb=0.0):
    """robust_nonrigid_alignment.t inputs on a jittered grid mesh: target = rest surface displaced along z"""
    rng = np.random.default_rng(seed)
    N = nx * ny
    xs, ys = np.meshgrid(np.arange(nx), np.arange(ny))
    ur = np.stack([xs, ys, np.zeros_like(xs)], -1).reshape(N, 3).astype(np.float64) + rng.normal(0, 0.01, (N, 3))
    off = ur.copy(); ang = np.zeros((N, 3)); rw = np.ones((N, 1))
    cons = ur + np.stack([np.zeros(N), np.zeros(N), 0.3 * np.sin(ur[:, 0] * 0.5)], 1)
    cons[rng.random(N) < 0.2] = -np.inf                      # no correspondence found
    nrm = np.tile(np.array([0.0, 0.0, 1.0]), (N, 1)) + rng.normal(0, 0.05, (N, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    if perturb > 0:
        off += rng.normal(0, perturb, off.shape); ang += rng.normal(0, 0.3 * perturb, ang.shape); rw += rng.normal(0, 0.1 * perturb, rw.shape)
    v0, v1 = grid_mesh_edges(nx, ny)
    c = lambda a: np.ascontiguousarray(a.astype(dtype))
    return [float(np.sqrt(w_fit)), float(np.sqrt(w_reg)), c(off), c(ang), c(rw), c(ur), c(cons), c(nrm), int(len(v0)), v0, v1]
