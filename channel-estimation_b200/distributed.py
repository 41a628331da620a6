"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL on the GPU box, gloo in CPU
tests).  Realizations are independent (DS.m:350 "Preallocate for Parfor"), so the i_rep axis is cut
into contiguous shards; no collective runs inside the loop body.  The only exchange is the final
sum of the integer error counters (and, when the per-realization BER arrays of DS.m:322-345 are
wanted, a gather of each rank's slice)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total, rank, world):
    """Contiguous block [lo, hi) of realization indices owned by `rank`; blocks differ by at most 1."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def run_sharded(run_fn, n_total, device=None, gather=True):
    """run_fn(first_rep, n_rep) -> err[n_rep, n_snr, n_it, 3, 2, 2] (uint32) for global realization
    indices first_rep .. first_rep + n_rep - 1 (results must depend on the index only, which the
    counter-based generator guarantees).  Returns (counters, err_all): the error counters summed
    over every realization of every rank, and -- on every rank when gather=True -- the
    per-realization array in global index order."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    lo, hi = shard_bounds(n_total, rank, world)
    local = np.asarray(run_fn(lo, hi - lo)) if hi > lo else None
    shapes = [None] * world
    if world > 1:
        dist.all_gather_object(shapes, None if local is None else local.shape[1:])
    else:
        shapes[0] = local.shape[1:]
    tail = next(s for s in shapes if s is not None)
    if local is None:
        local = np.zeros((0,) + tuple(tail), dtype=np.uint32)
    dev = device if device is not None else torch.device("cpu")
    counters = torch.from_numpy(local.astype(np.int64).sum(axis=0)).to(dev)
    if world > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)          # the path's only collective
    err_all = None
    if gather:
        if world > 1:
            n_max = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world))
            pad = np.zeros((n_max,) + tuple(tail), dtype=np.int64)
            pad[: local.shape[0]] = local
            mine = torch.from_numpy(pad).to(dev)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            chunks = []
            for r in range(world):
                a, b = shard_bounds(n_total, r, world)
                chunks.append(parts[r][: b - a].cpu().numpy())
            err_all = np.concatenate(chunks, axis=0).astype(np.uint32)
        else:
            err_all = local
    return counters.cpu().numpy(), err_all
