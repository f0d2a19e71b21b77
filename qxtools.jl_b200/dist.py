"""Multi-GPU partitioning of the hot path: one process per GPU.

The reference's two levels of parallelism (docs/src/users_guide.md:11-20; flags
``--mpi`` / ``--sub-comm-size`` of bin/qxrun.jl:40-46): output bitstrings are
divided among sub-communicators, and inside a sub-communicator the SLICES of each
contraction are divided among its ranks, whose partial sums are combined with one
reduction.  Here the communicator is ``torch.distributed`` (NCCL over NVLink on
the GPUs, gloo in the CPU tests); the only data that ever crosses the link is the
``[n_amp]`` complex vector of partial amplitudes.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def partition_range(n: int, parts: int, i: int) -> Tuple[int, int]:
    """Contiguous, balanced split of range(n): part i of ``parts``.  Contiguous so that a
    rank's slices form few aligned blocks (high slice variables constant)."""
    if parts < 1 or not (0 <= i < parts):
        raise ValueError("bad partition index")
    return (n * i) // parts, (n * (i + 1)) // parts


class Distribution:
    """Which bitstrings and which slice range this rank computes."""

    def __init__(self, n_bitstrings: int, n_slices: int, world: int = 1, rank: int = 0,
                 sub_comm_size: Optional[int] = None):
        if sub_comm_size is None:
            sub_comm_size = world                 # all ranks share the slices of every bitstring
        if sub_comm_size < 1 or world % sub_comm_size:
            raise ValueError("sub_comm_size must divide the number of ranks")
        self.world, self.rank, self.sub_comm_size = world, rank, sub_comm_size
        self.n_groups = world // sub_comm_size
        self.group = rank // sub_comm_size
        self.rank_in_group = rank % sub_comm_size
        self.amp_begin, self.amp_end = partition_range(n_bitstrings, self.n_groups, self.group)
        self.slice_begin, self.slice_end = partition_range(n_slices, sub_comm_size, self.rank_in_group)
        self.group_ranks = list(range(self.group * sub_comm_size, (self.group + 1) * sub_comm_size))
        self.n_bitstrings, self.n_slices = n_bitstrings, n_slices

    def amp_ranges(self) -> List[Tuple[int, int]]:
        return [partition_range(self.n_bitstrings, self.n_groups, g) for g in range(self.n_groups)]


def reduce_partial_amplitudes(partial, dist_module=None, group=None):
    """Sum the per-rank partial amplitudes in place (one all-reduce of 2*n_amp reals)."""
    import torch
    import torch.distributed as td
    dist_module = dist_module or td
    if not dist_module.is_initialized() or dist_module.get_world_size(group) == 1:
        return partial
    view = torch.view_as_real(partial) if partial.is_complex() else partial
    dist_module.all_reduce(view, group=group)
    return partial
