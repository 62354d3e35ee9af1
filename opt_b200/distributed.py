"""Host-side plumbing for the multi-GPU (row-slab) decomposition -- SURVEY.md section 8(e).

One process per GPU (`torch.distributed`, NCCL on GPUs, gloo in the CPU tests).  torch.distributed is used
for plumbing only: agreeing on the partition and all-gathering the CUDA IPC handles once per plan.  All
per-iteration traffic (halo rows of z / p, the two PCG scalars) is issued from inside the solver's kernels
over NVLink peer memory (opt_b200/csrc/solver.cu, common.cuh).
"""
import numpy as np


def slab_bounds(H, world):
    """owned global rows [b[r], b[r+1]) of each rank: contiguous, as equal as possible"""
    base, rem = divmod(H, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


class Slab:
    """Geometry of one rank's LOCAL image: owned rows plus one halo row towards each neighbour."""

    def __init__(self, H, world, rank, halo=1):
        b = slab_bounds(H, world)
        self.H, self.world, self.rank, self.halo = H, world, rank, halo
        self.g0, self.g1 = b[rank], b[rank + 1]                       # owned global rows
        self.top = halo if rank > 0 else 0                            # halo rows above / below
        self.bot = halo if rank + 1 < world else 0
        self.lo, self.hi = self.g0 - self.top, self.g1 + self.bot     # global rows held locally
        self.local_rows = self.hi - self.lo
        self.owned_begin, self.owned_end = self.top, self.top + (self.g1 - self.g0)

        def rows_of(r):
            if r < 0 or r >= world:
                return 0
            return (b[r + 1] - b[r]) + (halo if r > 0 else 0) + (halo if r + 1 < world else 0)
        self.upper_rows, self.lower_rows = rows_of(rank - 1), rows_of(rank + 1)
        if min(b[r + 1] - b[r] for r in range(world)) < 1:
            raise ValueError("more ranks than rows")

    def take(self, image, W):
        """rows [lo, hi) of a (H*W, C) or (H*W,) host array -> contiguous local copy"""
        a = np.asarray(image)
        return np.ascontiguousarray(a.reshape((self.H, W) + a.shape[1:])[self.lo:self.hi].reshape((self.local_rows * W,) + a.shape[1:]))

    def owned(self, local_image, W):
        a = np.asarray(local_image)
        return a.reshape((self.local_rows, W) + a.shape[1:])[self.owned_begin:self.owned_end]


def scatter_image_params(params, image_positions, W, slab):
    """local parameter list of one rank: image arrays cut to the slab (+halo), everything else unchanged"""
    return [slab.take(p, W) if i in image_positions else p for i, p in enumerate(params)]


def connect_plan(solver, slab, group=None):
    """all-gather the plans' IPC handles and wire the peer-memory mailboxes / halo targets (collective)"""
    import torch
    import torch.distributed as dist
    mine = solver.comm_handles()
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(device)
    out = [torch.empty_like(t) for _ in range(slab.world)]
    dist.all_gather(out, t, group=group)
    blob = b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)
    solver.connect(slab.rank, slab.world, blob, slab.owned_begin, slab.owned_end, slab.upper_rows, slab.lower_rows)
    dist.barrier(group=group)


def gather_owned_rows(local_image, W, slab, group=None):
    """rank 0 receives the full (H, W, ...) array assembled from every rank's owned rows (host numpy)"""
    import torch.distributed as dist
    mine = np.ascontiguousarray(slab.owned(local_image, W))
    parts = [None] * slab.world
    dist.all_gather_object(parts, mine, group=group)
    return np.concatenate(parts, 0)


class VertexSlab:
    """1-D decomposition of a graph problem over its (locality-ordered) vertex index -- SURVEY.md 8(e) "graph partition".
    Every rank binds the GLOBAL arrays (they are small: 12 B per vertex per array) but owns the vertices [g0, g1) and
    streams only that range plus `halo` vertices on each side, where halo = the graph's index bandwidth max|v0 - v1|
    rounded up to 4 (so a locality-preserving vertex order -- grid order, RCM, a space-filling curve -- is what makes
    the decomposition efficient).  Same wiring as the image slabs: OptB200_PlanConnect(owned range, halo depths)."""

    def __init__(self, n_vertices, world, rank, v0, v1):
        self.N, self.world, self.rank = int(n_vertices), world, rank
        b = [min(self.N, int(round(self.N * r / world / 4.0)) * 4) for r in range(world)] + [self.N]
        reach = int(np.abs(np.asarray(v0, np.int64) - np.asarray(v1, np.int64)).max()) if len(v0) else 0
        self.halo = (reach + 3) // 4 * 4
        if world > 1 and min(b[r + 1] - b[r] for r in range(world)) < max(self.halo, 4):
            raise ValueError(f"graph index bandwidth {reach} exceeds the per-rank vertex range: reorder the vertices or use fewer ranks")
        self.g0, self.g1 = b[rank], b[rank + 1]
        self.owned_begin, self.owned_end = self.g0, self.g1
        self.upper_rows = self.halo if rank > 0 else 0
        self.lower_rows = self.halo if rank + 1 < world else 0

    def gather(self, array, group=None):
        """assemble the global per-vertex array from every rank's owned rows (host numpy, all ranks receive it)"""
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, np.ascontiguousarray(np.asarray(array)[self.g0:self.g1]), group=group)
        return np.concatenate(parts, 0)
