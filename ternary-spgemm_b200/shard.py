"""N-column sharding across the GPUs of one box (north-star subsystem 4).

Column n of Y depends only on column n of W (reference cpp_impl/comp.h:39-63), and TCSC is
column-major, so rank r of G owns the contiguous column slice [N·r/G, N·(r+1)/G): a contiguous
range of all four TCSC arrays.  Per input there is exactly ONE collective — the broadcast of X
from rank 0 (M·K floats over NVLink/NVSwitch via NCCL) — and no reduction: every rank writes its
own Y[:, lo:hi].  `gather_columns` exists for verification / callers that want Y assembled; it
is not on the timed path.

torch.distributed is plumbing here (process group, NCCL/gloo); the compute is whatever
`compute(X, lo, hi)` the caller passes — libtsg's kernels on the GPU box, the checker in the
CPU gloo tests.
"""
from __future__ import annotations

from . import shard_columns  # noqa: F401  (re-export: the partition rule lives with the format)


def broadcast_x(X, src: int = 0, group=None):
    """Replicate the activation batch: the only data-path collective."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(X, src=src, group=group)
    return X


class PeerX:
    """X replicated WITHOUT a broadcast: rank `src` owns the activation buffer in symmetric memory
    (NVLink / NVSwitch peer mapping) and every other rank's SpMM kernel reads it in place through
    its peer pointer — the transfer happens inside the consuming kernel (for the tensor-core path
    it is the one read of X by its split kernel; for the decode kernels a few KB per CTA), so there
    is no collective launch and no second copy of X in HBM.

    Two buffers alternate, so one device-side barrier per step is enough: a rank enqueues barrier i
    after its kernel of step i-1, hence when barrier i has completed everywhere, buffer (i-1)%2 is
    free again for the owner's write of step i+1.

        px = PeerX(M, K, device)                 # collective: all ranks
        x = px.stage(host_or_device_X)           # rank src copies, everyone barriers; returns the view
        matrix.spmm_dev(x, b, Y, M, ...)         # kernel pulls X over NVLink
    """

    def __init__(self, M: int, K: int, device, *, group=None, src: int = 0):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.src, self.rank = src, dist.get_rank(group)
        g = group if group is not None else dist.group.WORLD
        self.local, self.hdl, self.view = [], [], []
        for _ in range(2):
            t = symm.empty((M, K), dtype=torch.float32, device=device)
            h = symm.rendezvous(t, g)
            self.local.append(t)
            self.hdl.append(h)
            self.view.append(t if self.rank == src else h.get_buffer(src, (M, K), torch.float32))
        self.step = 0

    def stage(self, X):
        """Publish this step's X (only rank `src` reads its argument).  Enqueued on the current
        stream; returns the tensor every rank passes to its kernel."""
        i = self.step & 1
        self.step += 1
        if self.rank == self.src:
            self.local[i].copy_(X, non_blocking=True)
        self.hdl[i].barrier(channel=0)
        return self.view[i]


def sharded_spmm(X, N: int, compute, *, group=None, src: int = 0):
    """Run one step of the sharded path on this rank: broadcast X, compute the local column
    slice.  Returns (Y_local, (lo, hi))."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_columns(N, world, rank)
    broadcast_x(X, src=src, group=group)
    return compute(X, lo, hi), (lo, hi)


def gather_columns(Y_local, N: int, group=None):
    """Assemble the full M×N result on every rank from the per-rank column slices (uneven
    slices allowed).  Verification helper — not part of the timed path."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return Y_local
    world = dist.get_world_size(group)
    M = Y_local.shape[0]
    widths = [shard_columns(N, world, r)[1] - shard_columns(N, world, r)[0] for r in range(world)]
    wmax = max(widths)
    pad = torch.zeros(M, wmax, dtype=Y_local.dtype, device=Y_local.device)
    pad[:, : Y_local.shape[1]] = Y_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:, :w] for p, w in zip(parts, widths)], dim=1)
