"""No GPU needed: a numpy emulation of the DATA FLOW of the state-on-chip PCG loop in its 2-D tile / multi-rank form
(opt_b200/csrc/energy_image_warping.cu, k_iwPcgLoopTile<T, K, MULTI>) against a plain global PCG.

The CUDA kernel cannot run here, but its scheme can be checked: per-tile state with a one-cell ring, p_ring = z_ring + beta * p_ring
from the z that neighbouring tiles (or the neighbouring RANK, through its halo row) published before the second all-reduce, edge
cells publishing z, rank-boundary tiles pushing their edge row into the neighbour's halo row, and the "shadow delta" that keeps the
update of X on halo rows equal to the owner's.  The emulation executes the phases in the kernel's order (phase A of every tile, all-reduce,
phase B of every tile, all-reduce) with the image_warping J^T J operator written in numpy; the result must equal a global PCG on the
undecomposed image to rounding.  This pins the protocol, not the CUDA code (that is tests/test_gpu_zz_*.py)."""
import numpy as np
import pytest


def make_problem(W, H, seed):
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:H, 0:W]
    U = np.stack([xs, ys], -1).astype(np.float64) + 0.1 * rng.normal(size=(H, W, 2))
    ang = 0.4 * rng.normal(size=(H, W))
    active = rng.random((H, W)) > 0.08
    fit = (rng.random((H, W)) < 0.2) & active
    geom = np.stack([np.cos(ang), np.sin(ang), U[..., 0], U[..., 1]], -1)
    # pair validity bits as k_iwInit computes them: neighbour in bounds and both pixels active; bit 4 fit, bit 5 active
    flags = np.zeros((H, W), np.int64)
    for k, (dx, dy) in enumerate(((1, 0), (-1, 0), (0, 1), (0, -1))):
        nb = np.zeros((H, W), bool)
        ys0, ys1 = max(0, -dy), H - max(0, dy); xs0, xs1 = max(0, -dx), W - max(0, dx)
        nb[ys0:ys1, xs0:xs1] = active[ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]
        flags |= ((nb & active).astype(np.int64) << k)
    flags |= fit.astype(np.int64) << 4
    flags |= active.astype(np.int64) << 5
    return geom, flags, 0.7, 1.3       # w (rigidity), wf (fit): the square-rooted weights of the spec


def apply_padded(pO, pA, geom, flags, w, wf):
    """(J^T J p) on the interior of arrays padded by one cell on every side (energy_image_warping.cu header formulas); flags for the interior only"""
    h, wd = flags.shape
    c = lambda a: a[1:-1, 1:-1]
    gi = c(geom); pOi = c(pO); pAi = c(pA)
    outO = np.zeros_like(pOi); outA = np.zeros_like(pAi)
    for k, (dx, dy) in enumerate(((1, 0), (-1, 0), (0, 1), (0, -1))):
        sl = (slice(1 + dy, 1 + dy + h), slice(1 + dx, 1 + dx + wd))
        gj = geom[sl]; pOj = pO[sl]; pAj = pA[sl]
        v = ((flags >> k) & 1).astype(np.float64)
        ddx = gi[..., 2] - gj[..., 2]; ddy = gi[..., 3] - gj[..., 3]
        eix = -gi[..., 1] * ddx - gi[..., 0] * ddy; eiy = gi[..., 0] * ddx - gi[..., 1] * ddy
        ejx = -gj[..., 1] * ddx - gj[..., 0] * ddy; ejy = gj[..., 0] * ddx - gj[..., 1] * ddy
        px = pOi[..., 0] - pOj[..., 0]; py = pOi[..., 1] - pOj[..., 1]
        outO[..., 0] += v * (2 * px - eix * pAi - ejx * pAj)
        outO[..., 1] += v * (2 * py - eiy * pAi - ejy * pAj)
        outA += v * ((eix * eix + eiy * eiy) * pAi - (eix * px + eiy * py))
    outO *= w * w; outA *= w * w
    f = ((flags >> 4) & 1).astype(np.float64) * wf * wf
    outO += f[..., None] * pOi
    act = ((flags >> 5) & 1).astype(np.float64)
    return outO * act[..., None], outA * act


def pad(a):
    return np.pad(a, [(1, 1), (1, 1)] + [(0, 0)] * (a.ndim - 2))


def global_pcg(geom, flags, w, wf, r0O, r0A, preO, preA, L):
    rO, rA = r0O.copy(), r0A.copy()
    pO, pA = preO * rO, preA * rA
    dO, dA = np.zeros_like(rO), np.zeros_like(rA)
    rz = (rO * pO).sum() + (rA * pA).sum()
    for _ in range(L):
        ApO, ApA = apply_padded(pad(pO), pad(pA), pad(geom), flags, w, wf)
        den = (pO * ApO).sum() + (pA * ApA).sum()
        alpha = rz / den if den > 0 else 0.0
        dO += alpha * pO; dA += alpha * pA
        rO -= alpha * ApO; rA -= alpha * ApA
        zO, zA = preO * rO, preA * rA
        rz_new = (zO * rO).sum() + (zA * rA).sum()
        beta = rz_new / rz if rz > 0 else 0.0
        pO, pA = zO + beta * pO, zA + beta * pA
        rz = rz_new
    return dO, dA


class Tile:
    pass


def tiled_pcg(geom, flags, w, wf, r0O, r0A, preO, preA, L, world, tw, th):
    """the kernel's scheme: `world` ranks own equal row slabs (+1 halo row per neighbour), tiles of tw x th over the owned rows"""
    H, W = flags.shape
    assert H % world == 0
    rows = H // world
    ranks = []
    for rk in range(world):
        g0 = rk * rows                                   # first owned global row
        up, dn = rk > 0, rk + 1 < world
        lo = g0 - (1 if up else 0); hi = g0 + rows + (1 if dn else 0)      # local domain = global rows [lo, hi)
        R = Tile(); R.lo, R.hi, R.y0, R.y1, R.up, R.dn = lo, hi, g0 - lo, g0 - lo + rows, up, dn
        R.z = np.zeros((hi - lo, W, 3))                  # the rank's z vector (local rows incl. halo rows)
        R.delta = np.zeros((hi - lo, W, 3))
        # p incl. halo rows: what exchangeHalo(haloP, p) leaves before the launch
        p0 = np.concatenate([preO * r0O, (preA * r0A)[..., None]], -1)
        R.p_init = p0[lo:hi].copy()
        R.tiles = []
        for ty in range(R.y0, R.y1, th):
            for tx in range(0, W, tw):
                t = Tile(); t.tx0, t.ty0 = tx, ty; t.twE = min(tw, W - tx); t.thE = min(th, R.y1 - ty)
                Hl = hi - lo
                def grab(arr, fill=0.0):                 # ring-included copy of a local-domain array, zero outside the local domain
                    out = np.full((t.thE + 2, t.twE + 2) + arr.shape[2:], fill)
                    for lr in range(t.thE + 2):
                        gy = ty - 1 + lr
                        if not (0 <= gy < Hl):
                            continue
                        x0, x1 = max(tx - 1, 0), min(tx + t.twE + 1, W)
                        out[lr, x0 - (tx - 1):x1 - (tx - 1)] = arr[gy, x0:x1]
                    return out
                t.p = grab(R.p_init)
                t.geom = grab(geom[lo:hi])
                own = (slice(ty, ty + t.thE), slice(tx, tx + t.twE))
                gown = (slice(lo + ty, lo + ty + t.thE), slice(tx, tx + t.twE))
                t.flags = flags[gown]
                t.r = np.concatenate([r0O[gown], r0A[gown][..., None]], -1)
                t.pre = np.concatenate([preO[gown], preA[gown][..., None]], -1)
                t.delta = np.zeros((t.thE, t.twE, 3)); t.z = np.zeros((t.thE, t.twE, 3)); t.Ap = np.zeros((t.thE, t.twE, 3))
                t.own = own
                t.shadowUp = R.up and ty == R.y0; t.shadowDn = R.dn and ty + t.thE == R.y1
                t.hUp = np.zeros((t.twE, 3)); t.hDn = np.zeros((t.twE, 3))
                R.tiles.append(t)
        ranks.append(R)
    rz = sum((t.r * t.p[1:-1, 1:-1]).sum() for R in ranks for t in R.tiles)          # r.p of PCGInit1, all-reduced
    rz_prev = 0.0
    for l in range(L):
        # ---------------- phase A on every tile of every rank
        acc = 0.0
        for R in ranks:
            Hl = R.hi - R.lo
            for t in R.tiles:
                if l > 0:
                    beta = rz / rz_prev if rz_prev > 0 else 0.0
                    t.p[1:-1, 1:-1] = t.z + beta * t.p[1:-1, 1:-1]
                    for lr, lc in ([(0, c + 1) for c in range(t.twE)] + [(t.thE + 1, c + 1) for c in range(t.twE)]
                                   + [(r + 1, 0) for r in range(t.thE)] + [(r + 1, t.twE + 1) for r in range(t.thE)]):
                        gx, gy = t.tx0 - 1 + lc, t.ty0 - 1 + lr
                        if 0 <= gx < W and 0 <= gy < Hl:
                            t.p[lr, lc] = R.z[gy, gx] + beta * t.p[lr, lc]
                ApO, ApA = apply_padded(t.p[..., :2], t.p[..., 2], t.geom, t.flags, w, wf)
                t.Ap = np.concatenate([ApO, ApA[..., None]], -1)
                acc += (t.p[1:-1, 1:-1] * t.Ap).sum()
        pAp = acc                                                              # all-reduce #1 (every tile of every rank)
        alpha = rz / pAp if pAp > 0 else 0.0
        # ---------------- phase B on every tile of every rank
        acc = 0.0
        pushes = []
        for ri, R in enumerate(ranks):
            for t in R.tiles:
                if t.shadowUp:
                    t.hUp += alpha * t.p[0, 1:-1]
                if t.shadowDn:
                    t.hDn += alpha * t.p[t.thE + 1, 1:-1]
                own = t.p[1:-1, 1:-1]
                t.delta += alpha * own
                t.r -= alpha * t.Ap
                t.z = t.pre * t.r
                acc += (t.z * t.r).sum()
                # edge cells publish z into the rank's z vector; rank-boundary rows also go to the neighbour's halo row
                for ly in range(t.thE):
                    for lx in range(t.twE):
                        if ly in (0, t.thE - 1) or lx in (0, t.twE - 1):
                            R.z[t.ty0 + ly, t.tx0 + lx] = t.z[ly, lx]
                            if R.up and t.ty0 + ly == R.y0:
                                pushes.append((ri - 1, ranks[ri - 1].hi - ranks[ri - 1].lo - 1, t.tx0 + lx, t.z[ly, lx].copy()))
                            if R.dn and t.ty0 + ly == R.y1 - 1:
                                pushes.append((ri + 1, 0, t.tx0 + lx, t.z[ly, lx].copy()))
        for (rk, row, x, val) in pushes:                                       # NVLink stores, visible after all-reduce #2
            ranks[rk].z[row, x] = val
        rz_prev, rz = rz, acc                                                  # all-reduce #2
    # write-back: delta of the owned pixels and the shadow delta of the neighbours' boundary rows
    for R in ranks:
        for t in R.tiles:
            R.delta[t.own] = t.delta
            if t.shadowUp:
                R.delta[R.y0 - 1, t.tx0:t.tx0 + t.twE] = t.hUp
            if t.shadowDn:
                R.delta[R.y1, t.tx0:t.tx0 + t.twE] = t.hDn
    return ranks


@pytest.mark.parametrize("W,H,world,tw,th", [(12, 8, 1, 12, 2), (12, 8, 1, 5, 3), (16, 12, 2, 16, 2), (16, 12, 2, 7, 4), (10, 12, 4, 4, 2), (9, 6, 3, 9, 1)])
def test_tile_scheme_equals_global_pcg(W, H, world, tw, th):
    geom, flags, w, wf = make_problem(W, H, seed=W * 100 + H)
    rng = np.random.default_rng(1)
    act = ((flags >> 5) & 1).astype(np.float64)
    r0O = rng.normal(size=(H, W, 2)) * act[..., None]; r0A = rng.normal(size=(H, W)) * act
    preO = (0.2 + rng.random((H, W, 1))) * np.ones((1, 1, 2)) * act[..., None]; preA = (0.2 + rng.random((H, W))) * act
    L = 7
    dO, dA = global_pcg(geom, flags, w, wf, r0O, r0A, preO, preA, L)
    ranks = tiled_pcg(geom, flags, w, wf, r0O, r0A, preO, preA, L, world, tw, th)
    ref = np.concatenate([dO, dA[..., None]], -1)
    scale = np.abs(ref).max()
    for R in ranks:
        # owned rows AND halo rows (shadow delta) must equal the global solution's delta on the same global rows
        got = R.delta
        want = ref[R.lo:R.hi]
        assert np.abs(got - want).max() <= 1e-11 * scale, (R.lo, R.hi, np.abs(got - want).max(), scale)


def test_single_reduction_recurrences_equal_standard_pcg():
    """the arrangement of k_iwCgLoopTile (matvec on z, gamma = r.z and delta = w.z in one reduction, alpha = gamma / (delta - beta gamma /
    alpha_old), A p carried as s = w + beta s, the loop ending after the L-th update) reproduces the reference's PCG iterates
    (solverGPUGaussNewton.t:421-550, same guards) in exact arithmetic -- checked in fp64 on a random SPD system with a Jacobi preconditioner"""
    rng = np.random.default_rng(11)
    n, L = 120, 25
    B = rng.normal(size=(n, n)); A = B @ B.T + n * np.eye(n)
    b = rng.normal(size=n); M = 1.0 / (1.0 + np.sqrt(np.diag(A))) ** 2        # the CERES-style guarded inverse the solver uses
    # reference arrangement
    x = np.zeros(n); r = b.copy(); z = M * r; p = z.copy(); rz = r @ z
    for _ in range(L):
        Ap = A @ p; den = p @ Ap
        alpha = rz / den if den > 0 else 0.0
        x = x + alpha * p; r = r - alpha * Ap; z = M * r; rz_new = r @ z
        beta = rz_new / rz if rz > 0 else 0.0
        p = z + beta * p; rz = rz_new
    # single-reduction arrangement
    xc = np.zeros(n); rc = b.copy(); u = M * rc; w = A @ u
    gam, dlt = rc @ u, w @ u
    alpha = gam / dlt if dlt > 0 else 0.0; beta = 0.0
    pc = np.zeros(n); s = np.zeros(n)
    for it in range(L):
        pc = u + beta * pc; s = w + beta * s
        xc = xc + alpha * pc; rc = rc - alpha * s
        u = M * rc
        if it + 1 == L:
            break
        w = A @ u
        g_new, d_new = rc @ u, w @ u
        beta = g_new / gam if gam > 0 else 0.0
        den = d_new - beta * g_new / alpha if alpha != 0 else d_new
        alpha = g_new / den if den > 0 else 0.0
        gam = g_new
    assert np.linalg.norm(xc - x) <= 1e-9 * np.linalg.norm(x)
    assert np.linalg.norm(rc - r) <= 1e-9 * np.linalg.norm(b)
