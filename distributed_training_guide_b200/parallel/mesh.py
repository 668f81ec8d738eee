"""2-D (dp, tp) device mesh.

Reference: ``init_device_mesh("cuda", (dp, tp), mesh_dim_names=("dp","tp"))`` with TP =
the whole node in chapter 06 (``06-tensor-parallel/train_llm.py:37-55``) and ``(world/tp,
tp)`` from ``--tensor-parallel`` in chapter 07 (``07-2d-parallel/train_llm.py:47-53``).  TP
ranks are contiguous (NVLink/NVSwitch neighbours), DP ranks strided.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import torch.distributed as dist


@dataclasses.dataclass
class Mesh2D:
    world_size: int
    rank: int
    dp_size: int
    tp_size: int
    dp_rank: int
    tp_rank: int
    dp_group: Optional[object] = None  # ProcessGroup over ranks with equal tp_rank (strided)
    tp_group: Optional[object] = None  # ProcessGroup over ranks with equal dp_rank (contiguous)
    dp_ranks: tuple = ()
    tp_ranks: tuple = ()

    def __repr__(self):
        return (f"Mesh2D(dp={self.dp_size}, tp={self.tp_size}, rank={self.rank} -> "
                f"dp_rank={self.dp_rank}, tp_rank={self.tp_rank})")


def build_mesh(world_size: int, rank: int, tp_size: int, create_groups: bool = True) -> Mesh2D:
    assert world_size % tp_size == 0, f"world size {world_size} not divisible by tp {tp_size}"
    dp_size = world_size // tp_size
    dp_rank, tp_rank = divmod(rank, tp_size)
    mesh = Mesh2D(world_size, rank, dp_size, tp_size, dp_rank, tp_rank)
    mesh.tp_ranks = tuple(dp_rank * tp_size + t for t in range(tp_size))
    mesh.dp_ranks = tuple(d * tp_size + tp_rank for d in range(dp_size))
    if create_groups and dist.is_initialized() and world_size > 1:
        # every rank must create every group, in the same order
        for d in range(dp_size):
            ranks = [d * tp_size + t for t in range(tp_size)]
            g = dist.new_group(ranks) if tp_size > 1 else None
            if d == dp_rank:
                mesh.tp_group = g
        for t in range(tp_size):
            ranks = [d * tp_size + t for d in range(dp_size)]
            g = dist.new_group(ranks) if dp_size > 1 else None
            if t == tp_rank:
                mesh.dp_group = g
    return mesh
