"""Turn a sharded (DCP) checkpoint of chapters 04-07 into one HF-named ``model.pt``.

    python -m distributed_training_guide_b200.tools.consolidate <exp_dir> -m meta-llama/Llama-2-7b-hf [--world N]

The sharded checkpoint stores, per flat group (``embed``, ``layer{i}``, ``head``), the concatenation of the
ranks' 1-D shards; this tool loads them in a single process, cuts the groups back into named parameters
(group layouts are derived from the model config) and writes ``<exp_dir>/model.pt``.  (The reference points to
``torch.distributed.checkpoint.format_utils.dcp_to_torch_save`` for the same job.)
"""
from __future__ import annotations

import argparse
from pathlib import Path

import torch

from ..models import build_model, get_config
from ..parallel.flat import build_groups


def consolidate(exp_dir: str, model_name: str, world: int) -> Path:
    import torch.distributed.checkpoint as dcp

    cfg = get_config(model_name)
    model = build_model(cfg, dtype=torch.bfloat16, device="cpu", init=False)
    groups = build_groups(model, "cpu", torch.bfloat16, world_size=world, with_grad=False)
    state = {"model": {g.name: torch.zeros(g.padded_numel, dtype=torch.bfloat16) for g in groups}}
    dcp.load(state, checkpoint_id=str(Path(exp_dir) / "checkpoint"))
    for g in groups:
        g.param.copy_(state["model"][g.name])
    out = Path(exp_dir) / "model.pt"
    torch.save(model.state_dict(), out)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("exp_dir")
    ap.add_argument("-m", "--model-name", required=True)
    ap.add_argument("--world", type=int, required=True, help="number of ranks that wrote the checkpoint")
    a = ap.parse_args()
    print(f"wrote {consolidate(a.exp_dir, a.model_name, a.world)}")


if __name__ == "__main__":
    main()
