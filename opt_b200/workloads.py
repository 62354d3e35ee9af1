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
