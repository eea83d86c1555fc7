"""Page-level data parallelism: pages are independent units, so a book is
sharded page -> rank with NO collective inside the pixel path.  torch.distributed is used only
to move bytes around it: scatter page batches from a root, gather byte results (NCCL over
NVLink when the tensors are on GPUs, gloo on CPU for the tests).

One process per GPU (torchrun); every function takes the process group implicitly.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_pages, rank, world):
    """Pages of this rank: i -> rank i mod world (all pages of a config share a shape: balanced)."""
    return list(range(rank, n_pages, world))


def shard_counts(n_pages, world):
    return [len(range(r, n_pages, world)) for r in range(world)]


def scatter_pages(pages, page_shape, n_pages, src=0, device=None):
    """Root holds `pages` (uint8 tensor [n_pages, *page_shape]); every rank receives its shard
    [len(shard_indices), *page_shape].  Point-to-point sends grouped per destination."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = shard_indices(n_pages, rank, world)
    out = torch.empty((len(mine),) + tuple(page_shape), dtype=torch.uint8, device=device)
    if rank == src:
        ops, keep = [], []
        for r in range(world):
            idx = shard_indices(n_pages, r, world)
            if not idx:
                continue
            chunk = pages[r::world].contiguous()             # pages i = r (mod world), in page order
            if r == src:
                out.copy_(chunk)
            else:
                keep.append(chunk)
                ops.append(dist.P2POp(dist.isend, chunk, r))
        if ops:                                              # one grouped launch: all sends in flight at once (NCCL group)
            for q in dist.batch_isend_irecv(ops):
                q.wait()
    elif mine:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.irecv, out, src)]):
            q.wait()
    return out


def gather_results(local, n_pages, dst=0):
    """Inverse of scatter_pages for a per-page result tensor [n_local, ...]: rank `dst` gets
    [n_pages, ...] in page order, the others None."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == dst:
        full = torch.empty((n_pages,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        ops, bufs = [], {}
        for r in range(world):
            cnt = len(range(r, n_pages, world))
            if not cnt:
                continue
            if r == dst:
                full[r::world] = local
            else:
                bufs[r] = torch.empty((cnt,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
                ops.append(dist.P2POp(dist.irecv, bufs[r], r))
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        for r, buf in bufs.items():
            full[r::world] = buf
        return full
    if local.shape[0]:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst)]):
            q.wait()
    return None


def max_over_ranks(value, device=None):
    """Device-timed seconds -> the slowest rank's (how multi-GPU numbers are reported)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
