"""Multi-GPU plumbing for the sigma build.

sigma = A.X is linear in the auxiliary index P of the density-fitting tensor and in the grid
points g, so every rank owns a contiguous aux block and a contiguous grid batch, computes a partial
sigma from them, and ONE all-reduce(sum, fp64) of the MO-space [nvec, dim] partial per `vind` call
combines them.  The Fock / spin-adaptation terms are local and replicated, added after the
reduction identically on every rank, so the Davidson state stays replicated without further
communication.  The reference has no distributed code; this is new.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple


def split_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of range(n): the first n % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) when absent."""
    return (int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


class SigmaReducer:
    """All-reduce of the partial sigma block over the process group (NCCL on GPU, gloo in CPU tests).

    One process per GPU; `torch.distributed` owns the communicator.  `world == 1` is a no-op.
    """

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.calls = 0
        self.bytes = 0

    def allreduce_(self, t):
        """In-place sum over ranks of a torch tensor (fp64).  Returns the tensor."""
        if self.enabled:
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
            self.calls += 1
            self.bytes += t.numel() * t.element_size()
        return t


def init_process_group_from_env(backend: Optional[str] = None):
    """Initialise torch.distributed from torchrun's environment (MASTER_ADDR defaults to 127.0.0.1)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank_world()
    if world == 1 or dist.is_initialized():
        return rank, local_rank, world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world
