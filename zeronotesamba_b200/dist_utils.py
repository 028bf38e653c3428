"""Data-parallel plumbing for the pretext step (one process per GPU, torch.distributed).

Partitioning: rank r owns source clips r, r + world, r + 2 world, ...; a step's 16
crops are shifts of ONE clip (pretext.py:312-321), so the NT-Xent loss never mixes clips across
ranks and the only exchange is one gradient all-reduce (sum) per step over the flat fp32 gradient
buffer; Adam then applies grad_scale = 1/world on every rank, which keeps replicas identical.
The reference itself has no data parallelism (it splits the two branches over two GPUs,
pretext.py:199-200); this is new functionality asked for by the north star.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def world_info() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_clips(n_clips: int, rank: int, world: int) -> List[int]:
    """Indices of the source clips rank ``rank`` processes; every rank gets the same count (the tail
    that does not divide evenly is dropped so that all ranks run the same number of steps)."""
    per_rank = n_clips // world
    return [rank + i * world for i in range(per_rank)]


def allreduce_gradients(flat_g: torch.Tensor) -> float:
    """Sum-all-reduce the flat gradient buffer in place; returns the scale Adam must apply."""
    if not is_distributed():
        return 1.0
    dist.all_reduce(flat_g, op=dist.ReduceOp.SUM)
    return 1.0 / dist.get_world_size()


def broadcast_parameters(flat_p: torch.Tensor, src: int = 0) -> None:
    if is_distributed():
        dist.broadcast(flat_p, src)


class SymmetricFlat:
    """Flat fp32 parameter and gradient buffers in NVLink-addressable symmetric memory
    (torch.distributed._symmetric_memory): ``p`` / ``g`` are this rank's tensors, ``p_ptrs`` / ``g_ptrs``
    ctypes arrays with every rank's device address of the same buffers, ``barrier(channel)`` a
    device-side cross-rank barrier enqueued on the current stream."""

    def __init__(self, numel: int, device: torch.device):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        group = dist.group.WORLD
        self.p = symm_mem.empty(numel, dtype=torch.float32, device=device)
        self.g = symm_mem.empty(numel, dtype=torch.float32, device=device)
        self.p.zero_()
        self.g.zero_()
        self._hp = symm_mem.rendezvous(self.p, group)
        self._hg = symm_mem.rendezvous(self.g, group)
        self.p_ptrs = (ctypes.c_void_p * self.world)(*[int(a) for a in self._hp.buffer_ptrs])
        self.g_ptrs = (ctypes.c_void_p * self.world)(*[int(a) for a in self._hg.buffer_ptrs])
        assert int(self._hp.buffer_ptrs[self.rank]) == self.p.data_ptr(), "symmetric buffer is not where rendezvous says"
        assert int(self._hg.buffer_ptrs[self.rank]) == self.g.data_ptr()
        torch.cuda.synchronize()
        dist.barrier()

    def barrier(self, channel: int = 0) -> None:
        self._hg.barrier(channel=channel)
