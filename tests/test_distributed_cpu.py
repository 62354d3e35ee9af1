"""world_size-2 gloo test (CPU) of the host-side multi-GPU plumbing: slab partition, scatter with halo rows,
IPC-handle all-gather wiring, gather of owned rows.  The kernels themselves need GPUs (tests/test_gpu_multi.py)."""
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
