"""GPU: the example harnesses' pre-steps on the device (include/OptB200Examples.h) against the host loops they replace, restated in numpy
(image_warping/src/CombinedSolver.h:181-207, arap_mesh_deformation/src/CombinedSolver.h:90-108, robust_nonrigid_alignment/src/
CombinedSolver.h:261-346 with a brute-force nearest-neighbour search in place of nanoflann)."""
import os

import numpy as np
import pytest

from opt_b200 import examples

pytestmark = pytest.mark.gpu
F = examples.FIXTURES


@pytest.mark.parametrize("alpha", [1.0, 0.25])
@pytest.mark.parametrize("dp", [False, True])
def test_image_warping_constraint_image(built, alpha, dp):
    import torch
    dt = np.float64 if dp else np.float32
    host, (W, H) = examples.image_warping_from_files(os.path.join(F, "cat512_mask.png"), os.path.join(F, "cat512.constraints"), alpha=alpha, dtype=dt)
    rows = [list(r) for r in examples.io.read_pair_constraints(os.path.join(F, "cat512.constraints"))]
    rows += [[x, y, x, y] for y in range(H) for x in range(W) if y == 0 or x == 0 or y == H - 1 or x == W - 1]
    d_rows = torch.tensor(rows, dtype=torch.int32, device="cuda")
    d_mask = torch.from_numpy(host[4]).cuda()
    d_cons = torch.full((W * H, 2), 7.0, dtype=d_mask.dtype, device="cuda")
    examples.image_warping_constraints_gpu(d_cons, d_mask, W, H, d_rows, alpha)
    assert np.array_equal(d_cons.cpu().numpy(), host[3])


def test_arap_constraints(built):
    import torch
    host, N = examples.arap_from_files(os.path.join(F, "small_armadillo.ply"), os.path.join(F, "small_armadillo.mrk"), alpha=0.3)
    pos, _, idx = examples.io.read_markers(os.path.join(F, "small_armadillo.mrk"))
    d_cur = torch.from_numpy(host[2]).cuda()
    d_cons = torch.zeros((N, 3), dtype=torch.float32, device="cuda")
    examples.arap_constraints_gpu(d_cons, d_cur, torch.from_numpy(idx).cuda(), torch.from_numpy(pos).cuda(), 0.3)
    got, want = d_cons.cpu().numpy(), host[5]
    assert np.array_equal(np.isfinite(got), np.isfinite(want)) and int(np.isfinite(want[:, 0]).sum()) == 3
    fin = np.isfinite(want)
    assert np.all(np.isneginf(got[~fin])) and np.allclose(got[fin], want[fin], rtol=1e-6, atol=0)     # (1-a) x + a t: one rounding apart at most (FMA contraction)


def _normals_ref(v, f):
    n = np.zeros_like(v, dtype=np.float64)
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]).astype(np.float64)
    fn /= np.maximum(np.linalg.norm(fn, axis=1, keepdims=True), 1e-30)
    for k in range(3):
        np.add.at(n, f[:, k], fn)
    return n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)


def test_vertex_normals(built):
    import torch
    v, f = examples.io.read_mesh(os.path.join(F, "small_armadillo.ply"))
    n = examples.vertex_normals_gpu(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda()).cpu().numpy()
    assert np.abs(n - _normals_ref(v, f)).max() < 2e-5
    n64 = examples.vertex_normals_gpu(torch.from_numpy(v.astype(np.float64)).cuda(), torch.from_numpy(f).cuda()).cpu().numpy()
    assert np.abs(n64 - n).max() < 2e-6


@pytest.mark.parametrize("dp", [False, True])
def test_rna_correspondences_match_brute_force(built, dp):
    import torch
    rng = np.random.default_rng(7)
    N, M, thr, cos_thr, K = 3000, 4000, 0.12, 0.7, 20
    src = rng.random((N, 3)).astype(np.float32); tgt = rng.random((M, 3)).astype(np.float32)
    sn = rng.normal(size=(N, 3)); sn = (sn / np.linalg.norm(sn, axis=1, keepdims=True)).astype(np.float32)
    tn = rng.normal(size=(M, 3)); tn = (tn / np.linalg.norm(tn, axis=1, keepdims=True)).astype(np.float32)
    # reference: K nearest (ties by index), stop at the first beyond thr, first with a compatible normal
    want = np.full((N, 3), -np.inf, np.float32); want_n = np.zeros((N, 3), np.float32); hit = np.zeros(N, bool)
    for i in range(N):
        d2 = ((tgt - src[i]) ** 2).sum(1, dtype=np.float32)
        order = np.lexsort((np.arange(M), d2))[:K]
        for t in order:
            if np.sqrt(d2[t]) > thr:
                break
            if float(np.dot(tn[t], sn[i])) > cos_thr:
                want[i] = tgt[t]; want_n[i] = tn[t]; hit[i] = True
                break
    assert 100 < hit.sum() < N - 100            # both outcomes are exercised
    dt = torch.float64 if dp else torch.float32
    d_tgt, d_tn = torch.from_numpy(tgt).cuda(), torch.from_numpy(tn).cuda()
    d_src = torch.from_numpy(src).cuda().to(dt); d_sn = torch.from_numpy(sn).cuda()
    cons = torch.zeros((N, 3), dtype=dt, device="cuda"); cn = torch.zeros((N, 3), dtype=dt, device="cuda")
    w = torch.full((N,), 0.5, dtype=dt, device="cuda"); prev = torch.full((N, 3), -9.0e13, dtype=torch.float32, device="cuda")
    g = examples.PointGrid(d_tgt, thr)
    changed = g.correspondences(d_tn, d_src, d_sn, cons, cn, w, prev, thr, cos_thr, K)
    got = cons.cpu().numpy().astype(np.float32)
    # distances in fp32 on both sides, but summed in a different order: allow a different pick only at an exact near-tie
    same = np.all((got == want) | (np.isinf(got) & np.isinf(want)), axis=1)
    assert same.mean() > 0.999, same.mean()
    assert np.array_equal(cn.cpu().numpy().astype(np.float32)[same & hit], want_n[same & hit])
    assert changed == N and bool((w == 1).all())          # every constraint is new the first time
    w.fill_(0.25)
    changed2 = g.correspondences(d_tn, d_src, d_sn, cons, cn, w, prev, thr, cos_thr, K)
    assert changed2 == 0 and bool((w == 0.25).all())      # nothing moved: no constraint changes, weights untouched
    g.close()
