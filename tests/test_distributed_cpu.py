"""world_size-2/3 gloo tests (CPU) of the host-side multi-GPU plumbing: slab partition, scatter with halo rows,
IPC-handle all-gather wiring, gather of owned rows -- and of the DECOMPOSITION itself: the row-slab PCG the kernels implement
(one halo row, p / delta recomputed redundantly on the halo, z of the boundary rows exchanged, dot products over owned rows only,
two all-reduces per step) is run here with the CPU oracle as the per-slab operator and gloo as the transport,
and must reproduce the single-domain reference iterates (solverGPUGaussNewton.t:361-550).  The kernels themselves need GPUs
(tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_geometry():
    from opt_b200.distributed import Slab, slab_bounds
    for H, world in ((2048, 8), (10, 3), (7, 7), (5, 1)):
        b = slab_bounds(H, world)
        assert b[0] == 0 and b[-1] == H and all(b[i + 1] > b[i] for i in range(world))
        for r in range(world):
            s = Slab(H, world, r)
            assert s.local_rows == (s.g1 - s.g0) + s.top + s.bot
            assert s.owned_end - s.owned_begin == s.g1 - s.g0
            if r > 0:
                up = Slab(H, world, r - 1)
                assert s.upper_rows == up.local_rows and up.lower_rows == s.local_rows
                # my first owned global row is the row the upper neighbour keeps as its bottom halo
                assert up.hi - 1 == s.g0
    with pytest.raises(ValueError):
        Slab(3, 4, 0)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    import torch.distributed as dist
    from opt_b200.distributed import Slab, scatter_image_params, connect_plan, gather_owned_rows
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H = 8, 11
    full = [np.arange(H * W * 2, dtype=np.float32).reshape(-1, 2), np.arange(H * W, dtype=np.float32), 3.0]
    slab = Slab(H, world, rank)
    loc = scatter_image_params(full, (0, 1), W, slab)
    assert loc[0].shape == (slab.local_rows * W, 2) and loc[2] == 3.0
    assert np.array_equal(loc[1].reshape(slab.local_rows, W)[slab.owned_begin], full[1].reshape(H, W)[slab.g0])

    class FakeSolver:   # records what the C ABI would receive
        def comm_handles(self):
            return bytes([rank + 1]) * 256
        def connect(self, *a):
            self.args = a
    fs = FakeSolver()
    connect_plan(fs, slab)
    r_, w_, blob, ob, oe, up, dn = fs.args
    assert (r_, w_, ob, oe, up, dn) == (rank, world, slab.owned_begin, slab.owned_end, slab.upper_rows, slab.lower_rows)
    assert len(blob) == 256 * world and all(blob[256 * r] == r + 1 for r in range(world))
    loc[1][:] = -1.0
    slab.owned(loc[1], W)[:] = rank                     # only owned rows count
    g = gather_owned_rows(loc[1], W, slab)
    exp = np.concatenate([np.full((Slab(H, world, r).g1 - Slab(H, world, r).g0, W), r, np.float32) for r in range(world)])
    assert np.array_equal(g, exp)
    dist.destroy_process_group()
    q.put(rank)


def test_gloo_world2_plumbing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]


def test_vertex_slab_geometry():
    from opt_b200 import workloads
    from opt_b200.distributed import VertexSlab
    v0, v1 = workloads.grid_mesh_edges(12, 9)
    N = 12 * 9
    reach = int(np.abs(v0.astype(np.int64) - v1).max())            # grid order: one row + one vertex
    assert reach == 13
    for world in (1, 2, 3):
        slabs = [VertexSlab(N, world, r, v0, v1) for r in range(world)]
        assert slabs[0].g0 == 0 and slabs[-1].g1 == N
        for r, s in enumerate(slabs):
            assert s.halo == 16 and s.g0 % 4 == 0 and s.g1 > s.g0
            assert s.upper_rows == (16 if r > 0 else 0) and s.lower_rows == (16 if r + 1 < world else 0)
            if r > 0:
                assert slabs[r - 1].g1 == s.g0
            # every edge with an owned endpoint stays inside owned range +- halo: what the rank streams is all it needs
            touch = (v0 >= s.g0) & (v0 < s.g1)
            assert v1[touch].min() >= s.g0 - s.halo and v1[touch].max() < s.g1 + s.halo
    with pytest.raises(ValueError):
        VertexSlab(N, 8, 0, v0, v1)                                  # 13 vertices per rank < the graph's index bandwidth


def _pcg_worker(rank, world, port, q, W, H, k):
    sys.path.insert(0, ROOT)
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world), "OMP_NUM_THREADS": "1"})
    import torch
    import torch.distributed as dist
    from opt_b200 import workloads
    from opt_b200.distributed import Slab, scatter_image_params, gather_owned_rows
    from oracle.oracle import Oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = workloads.image_warping(W, H, dtype=np.float64, seed=11, mask="random", perturb=0.3)
    slab = Slab(H, world, rank)
    loc = scatter_image_params(full, (0, 1, 2, 3, 4), W, slab)
    Hl = slab.local_rows; Nl = W * Hl
    o = Oracle("image_warping", "gaussNewtonGPU", [W, Hl], True)
    o.set_solver_parameter("nIterations", 1); o.set_solver_parameter("lIterations", 50)
    o.set_max_linear_iterations(0)
    work = workloads.clone_params(loc)
    o.init(work); o.step(work)                                      # PCGInit1 on the slab: r = -J^T F, preconditioner (exact on the owned rows)
    r = o.get_vector("r"); pre = o.get_vector("preconditioner")
    own = np.zeros(Hl, bool); own[slab.owned_begin:slab.owned_end] = True
    ownF = np.concatenate([np.repeat(own, W * 2), np.repeat(own, W)])     # flat layout: Offset (N x 2), then Angle (N)
    r = np.where(ownF, r, 0.0)

    def rows(v):        # (Hl, W*3) view: row y = [Offset row | Angle row], so that one row slice is a halo message
        return np.concatenate([v[:2 * Nl].reshape(Hl, 2 * W), v[2 * Nl:].reshape(Hl, W)], 1)

    def unrows(m):
        return np.concatenate([m[:, :2 * W].reshape(-1), m[:, 2 * W:].reshape(-1)])

    def exchange(v):    # boundary rows of an owned-valid vector -> the neighbours' halo rows
        m = rows(v).copy()
        msgs = [None] * world
        dist.all_gather_object(msgs, (m[slab.owned_begin].copy(), m[slab.owned_end - 1].copy()))
        if slab.top:
            m[slab.owned_begin - 1] = msgs[rank - 1][1]
        if slab.bot:
            m[slab.owned_end] = msgs[rank + 1][0]
        return unrows(m)

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64); dist.all_reduce(t); return float(t[0])

    z = exchange(np.where(ownF, pre * r, 0.0))
    p = z.copy(); delta = np.zeros_like(p)
    rz = allsum(float((z * r)[ownF].sum()))
    for _ in range(k):
        Ap = np.where(ownF, o.apply_jtj(loc, p), 0.0)              # needs p on the halo rows; valid on the owned rows
        pAp = allsum(float((p * Ap)[ownF].sum()))
        alpha = rz / pAp if pAp > 0 else 0.0
        delta = delta + alpha * p                                   # also on the halo rows (redundantly, same alpha everywhere)
        r = r - alpha * Ap
        z = exchange(np.where(ownF, pre * r, 0.0))
        rz1 = allsum(float((z * r)[ownF].sum()))
        beta = rz1 / rz if rz > 0 else 0.0
        p = z + beta * p                                            # halo rows: neighbour's z + beta x my redundant copy of its p
        rz = rz1
    out = {}
    for name, v in (("delta", delta), ("r", r), ("p", p)):
        m = rows(v)
        out[name] = gather_owned_rows(np.ascontiguousarray(m), 1, slab).reshape(H, 3 * W)      # "image" of width 1 whose pixels are whole rows
    # a halo row of delta / p must equal its owner's copy bitwise-close (same arithmetic on both sides)
    halo_ok = True
    for name, v in (("delta", delta), ("p", p)):
        m = rows(v); msgs = [None] * world
        dist.all_gather_object(msgs, (m[slab.owned_begin].copy(), m[slab.owned_end - 1].copy()))
        if slab.top:
            halo_ok = halo_ok and np.allclose(m[slab.owned_begin - 1], msgs[rank - 1][1], rtol=1e-12, atol=1e-14)
        if slab.bot:
            halo_ok = halo_ok and np.allclose(m[slab.owned_end], msgs[rank + 1][0], rtol=1e-12, atol=1e-14)
    if rank == 0:
        g = Oracle("image_warping", "gaussNewtonGPU", [W, H], True)
        g.set_solver_parameter("nIterations", 1); g.set_solver_parameter("lIterations", 50)
        g.set_max_linear_iterations(k)
        fw = workloads.clone_params(full)
        g.init(fw); g.step(fw)
        N = W * H
        err = {}
        for name in ("delta", "r", "p"):
            ref = g.get_vector(name)
            refm = np.concatenate([ref[:2 * N].reshape(H, 2 * W), ref[2 * N:].reshape(H, W)], 1)
            err[name] = float(np.abs(out[name] - refm).max() / max(1e-300, np.abs(refm).max()))
        q.put(("err", err))
    dist.barrier()
    dist.destroy_process_group()
    q.put(("halo", rank, bool(halo_ok)))


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_slab_pcg_reproduces_single_domain_iterates(built, world):
    """k = 12 PCG iterations of image_warping (fp64, masked, perturbed) on `world` row slabs over gloo against the oracle's
    single-domain state after the same 12 iterations: delta, r, p to 1e-9 of their inf-norms (different summation order only)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 300) + world
    procs = [ctx.Process(target=_pcg_worker, args=(r, world, port, q, 20, 17, 12)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    got = [q.get() for _ in range(world + 1)]
    err = [g[1] for g in got if g[0] == "err"][0]
    assert all(v <= 1e-9 for v in err.values()), err
    assert all(g[2] for g in got if g[0] == "halo")
