"""Pretrained-weight loading for sharded engines (reference: rank-0 ``from_pretrained`` on CPU +
``set_model_state_dict(broadcast_from_rank0=True)`` + manual buffer broadcasts,
``05-training-llama-405b/train_llm.py:76-145``).

Because every group is a flat buffer, distribution is per group instead of per tensor: rank 0 copies the HF
tensors of one group into a staging flat tensor (names are identical to ours), broadcasts it, and every rank
keeps the slice it owns.  Peak host memory on rank 0 is one checkpoint file + one group; nothing like the
reference's 764 GB resident model is needed.  There are no non-persistent buffers to broadcast (RoPE tables
are recomputed from the config).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist


def load_into_fsdp(engine, get_tensor: Optional[Callable[[str], torch.Tensor]], src_rank: int = 0):
    """``get_tensor(name)`` is only called on ``src_rank`` (e.g. a safetensors ``get_tensor``)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    for g, sh in zip(engine.groups, engine.shards):
        staging = torch.zeros(g.padded_numel, dtype=g.param.dtype, device=g.param.device)
        if rank == src_rank:
            for name, off, shape in zip(g.names, g.offsets, g.shapes):
                t = get_tensor(name)
                staging[off:off + t.numel()].copy_(t.reshape(-1).to(staging.dtype))
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(staging, src=src_rank)
        per = sh.padded_numel
        sh.param.copy_(staging[engine.rank * per:(engine.rank + 1) * per])
        del staging


def load_state_dict_into_flat(model, state_dict: Dict[str, torch.Tensor]):
    """Replicated engines (chapters 01/02): in-place copy into the flat views."""
    own = model.state_dict()
    with torch.no_grad():
        for k, v in state_dict.items():
            own[k].copy_(v)
