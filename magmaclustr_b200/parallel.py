"""Multi-GPU sharding: tasks are partitioned contiguously across ranks; the only
exchange is the sum all-reduce of the union-grid precision sums (U^2 + U doubles) per E step and
of scalars (log-likelihood, shared-HP objective/gradient).  torch.distributed does the plumbing."""
import numpy as np


def shard_tasks(offs, rank, world, n_grid=0):
    """Contiguous range [lo, hi) of tasks for `rank`.

    The boundaries are those of the library's E-step leaves (mb_shard_range): every sum over tasks is taken leaf by leaf
    and then in a fixed tree over the leaves, so a rank has to hold whole leaves -- that is what makes post_mean, post_cov,
    the log-likelihood and the EM step count bit-identical on 1, 2, 4 and 8 GPUs.  `n_grid` = size U of the union grid
    (it bounds the number of leaves through the accumulator memory).  Leaves hold equal task COUNTS; for strongly ragged
    data the shards are balanced in tasks, not in sum N_i^3."""
    import ctypes as C
    from . import _lib as L
    offs = np.asarray(offs, dtype=np.int64)
    m = len(offs) - 1
    if world <= 1:
        return 0, m
    lo, hi = C.c_int64(), C.c_int64()
    L.check(L.lib().mb_shard_range(m, int(n_grid), int(rank), int(world), C.byref(lo), C.byref(hi)))
    return int(lo.value), int(hi.value)


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
