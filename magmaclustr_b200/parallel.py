"""Multi-GPU sharding: tasks are partitioned contiguously across ranks; the only
exchange is the sum all-reduce of the union-grid precision sums (U^2 + U doubles) per E step and
of scalars (log-likelihood, shared-HP objective/gradient).  torch.distributed does the plumbing."""
import numpy as np


def shard_tasks(offs, rank, world):
    """Contiguous range [lo, hi) of tasks for `rank`, balanced by sum N_i^3 (ragged tasks)."""
    offs = np.asarray(offs, dtype=np.int64)
    m = len(offs) - 1
    if world <= 1:
        return 0, m
    cost = np.diff(offs).astype(np.float64) ** 3
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    bounds = [int(np.searchsorted(cum, cum[-1] * r / world, side="left")) for r in range(world + 1)]
    bounds[0], bounds[-1] = 0, m
    for r in range(1, world + 1):          # keep the partition monotone and non-empty when possible
        bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds[rank], bounds[rank + 1]


class _CudaView:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 2}


def make_allreduce(dist):
    """All-reduce hook for Context.set_allreduce backed by torch.distributed (NCCL)."""
    import torch

    def hook(dev_ptr, count):
        t = torch.as_tensor(_CudaView(dev_ptr, count), device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
    return hook


def init_library_nccl(ctx, dist, rank, world):
    """The library's own NCCL communicator (mb_nccl_init): rank 0 creates the id, torch.distributed only carries the
    128 bytes to the other ranks."""
    box = [ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.nccl_init(box[0], rank, world)
