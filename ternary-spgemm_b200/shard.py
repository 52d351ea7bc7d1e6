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
