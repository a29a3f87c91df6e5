"""Multi-GPU plumbing of the VQ path: rows sharded by rank, codebook replicated,
ONE exchange step — the sum over ranks of the usage histogram ``hist[K]`` (uint64) and the
squared-error sum ``sse`` (float64) — between ``dvq_vq_forward`` and ``dvq_vq_finalize``.
The integer histogram makes perplexity independent of the number of ranks exactly; the loss
differs only by fp64 summation order.  The inference path and PointNet need no collective.

One process per GPU, launched by torchrun; ``torch.distributed`` provides rendezvous and the
communicator.  By default the reduction is two ``torch.distributed.all_reduce`` calls on views of the
packed stats tensor (hist as int64, sse as float64) — inside ProcessGroupNCCL's own sequencing, watchdog
and flight recorder.  ``use_raw_nccl(True)`` opts into the C-ABI hook instead (``dvq_allreduce_stats``:
both buffers in ONE NCCL group enqueued on the caller's current stream with torch's ``ncclComm_t``, no
stream hop — ~20 us less per step, but outside torch's bookkeeping and dependent on the private
``_comm_ptr()`` accessor).
"""
from __future__ import annotations

import torch
import torch.distributed as tdist


def shard_bounds(n_rows: int, rank: int, world: int):
    """Contiguous row block [lo, hi) of rank ``rank``: sizes differ by at most one row."""
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


_RAW_NCCL = False


def use_raw_nccl(on: bool = True) -> None:
    """Opt in to / out of the raw-NCCL hook (see the module docstring)."""
    global _RAW_NCCL
    _RAW_NCCL = bool(on)


def _nccl_comm_ptr(group, device):
    try:
        backend = group._get_backend(torch.device(device))
        ptr = backend._comm_ptr()
        return int(ptr) if ptr else None
    except Exception:
        return None


def allreduce_stats(stats: torch.Tensor, n_e: int, n_local: int, group=None) -> int:
    """In-place sum over ranks of ``stats = [hist (n_e x int64 bit-pattern of uint64) | sse (float64 bits)]``.
    Returns the row count to hand to ``dvq_vq_finalize``: ``n_local`` on one rank, else 0 ("use the
    histogram total" — every row is counted once, so no second collective and no host sync)."""
    if group is None:
        group = tdist.group.WORLD
    world = tdist.get_world_size(group)
    if world == 1:
        return int(n_local)
    hist = stats[:n_e]
    sse = stats[n_e:n_e + 1].view(torch.float64)
    done = False
    if _RAW_NCCL and stats.is_cuda and tdist.get_backend(group) == "nccl":
        comm = _nccl_comm_ptr(group, stats.device)
        if comm is not None:
            from . import _cabi
            with torch.cuda.device(stats.device):
                _cabi.check(_cabi.lib.dvq_allreduce_stats(
                    comm, hist.data_ptr(), sse.data_ptr(), n_e,
                    torch.cuda.current_stream(stats.device).cuda_stream), "dvq_allreduce_stats")
            done = True
    if not done:
        tdist.all_reduce(hist, op=tdist.ReduceOp.SUM, group=group)
        tdist.all_reduce(sse, op=tdist.ReduceOp.SUM, group=group)
    return 0


def shard_module(module, group=None):
    """Mark every ``dvq.VectorQuantizer`` under ``module`` as row-sharded over ``group``."""
    from .quantizer import VectorQuantizer
    if group is None:
        group = tdist.group.WORLD
    for m in module.modules():
        if isinstance(m, VectorQuantizer):
            m.process_group = group
    return module


def world_size(group=None) -> int:
    return tdist.get_world_size(group if group is not None else tdist.group.WORLD)


def all_reduce_sum(t: torch.Tensor, group=None) -> None:
    tdist.all_reduce(t, op=tdist.ReduceOp.SUM, group=group if group is not None else tdist.group.WORLD)


def broadcast0(t: torch.Tensor, group=None) -> None:
    """Broadcast from the group's rank 0 (replicated state such as re-seeded codebook rows)."""
    g = group if group is not None else tdist.group.WORLD
    tdist.broadcast(t, src=tdist.get_global_rank(g, 0), group=g)
