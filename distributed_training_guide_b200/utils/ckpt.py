"""Checkpoint / resume with the reference's on-disk layout:

    <save-dir>/<experiment-name>/
        state.json        {"epoch","global_step","epoch_step","running_loss"}
        lr_scheduler.pt
        model.pt          chapters 01 / 02 (full state dict, rank 0)
        optimizer.pt      chapter 01 (reference chapter 02 drops it; we keep per-rank shards
                          ``optimizer.rank{r}.pt`` so Adam moments survive a resume, fixing §8 #8)
        checkpoint/       chapters 04-07: torch.distributed.checkpoint directory
                          (``.metadata`` + ``__{rank}_{i}.distcp``), keys "model"/"optimizer"
        rng.pt            only with --deterministic (reference related-topics/determinism)

Resume is decided only by the existence of ``state.json`` (reference ``01-single-gpu/
train_llm.py:94``).  Keys in ``model.pt`` are plain HF names (no ``_orig_mod.`` / ``module.``
wrapper prefixes, fixing §8 #9).
"""
from __future__ import annotations

import json
import os
import random
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

STATE_KEYS = ("epoch", "global_step", "epoch_step", "running_loss")


def new_state():
    return {"epoch": 0, "global_step": 0, "epoch_step": 0, "running_loss": 0}


def experiment_dir(args):
    """(is_experiment, exp_dir) — experiment mode only when ``-e`` is given (``01:80-84``)."""
    exp_dir = Path(args.save_dir)
    if args.experiment_name is None:
        return False, exp_dir
    return True, exp_dir / args.experiment_name


def can_resume(exp_dir: Path) -> bool:
    return (Path(exp_dir) / "state.json").exists()


def _atomic_save(obj, path: Path):
    tmp = Path(str(path) + ".tmp")
    torch.save(obj, tmp)
    os.replace(tmp, path)


def save_state_json(exp_dir: Path, state: dict):
    tmp = exp_dir / "state.json.tmp"
    with open(tmp, "w") as fp:
        json.dump({k: state[k] for k in STATE_KEYS}, fp)
    os.replace(tmp, exp_dir / "state.json")  # written last: its presence marks a complete checkpoint


def load_state_json(exp_dir: Path) -> dict:
    with open(Path(exp_dir) / "state.json") as fp:
        return json.load(fp)


# -- RNG (determinism recipe) ---------------------------------------------------------------
def rng_state():
    st = {"np": np.random.get_state(), "random": random.getstate(), "torch": torch.get_rng_state()}
    if torch.cuda.is_available():
        st["cuda"] = torch.cuda.get_rng_state_all()
    return st


def set_rng_state(st):
    np.random.set_state(st["np"])
    random.setstate(st["random"])
    torch.set_rng_state(st["torch"])
    if "cuda" in st and torch.cuda.is_available():
        torch.cuda.set_rng_state_all(st["cuda"])


def _restore_lr(lr_scheduler):
    """``LRScheduler.load_state_dict`` restores the schedule position but not the optimizer's current ``lr`` (torch
    relies on ``optimizer.load_state_dict`` for that); our flat optimizers persist only moments and step counters,
    so put the scheduled value back — otherwise the first resumed step runs at the initial learning rate."""
    opt = getattr(lr_scheduler, "optimizer", None)
    last = getattr(lr_scheduler, "_last_lr", None)
    if opt is not None and last is not None:
        for g, lr in zip(opt.param_groups, last):
            g["lr"] = lr


# -- full (unsharded) checkpoints: chapters 01 / 02 -------------------------------------------
def save_full(exp_dir: Path, model, optimizer, lr_scheduler, state, rank=0, world_size=1,
              save_optimizer=True, deterministic=False):
    """Rank 0 writes model/lr_scheduler/state; every rank writes its optimizer shard."""
    exp_dir = Path(exp_dir)
    # every rank: on node-local disks the other nodes' ranks write their optimizer shard into their own copy of the
    # directory (reference chapters 02/03 cover that layout); exist_ok makes the shared-mount case a no-op
    exp_dir.mkdir(parents=True, exist_ok=True)
    if world_size > 1:
        dist.barrier()
    if save_optimizer and optimizer is not None:
        name = "optimizer.pt" if world_size == 1 else f"optimizer.rank{rank}.pt"
        _atomic_save(optimizer.state_dict(), exp_dir / name)
    if deterministic:
        _atomic_save(rng_state(), exp_dir / ("rng.pt" if world_size == 1 else f"rng.rank{rank}.pt"))
    if rank == 0:
        _atomic_save(model.state_dict(), exp_dir / "model.pt")
        _atomic_save(lr_scheduler.state_dict(), exp_dir / "lr_scheduler.pt")
    if world_size > 1:
        dist.barrier()
    if rank == 0:
        save_state_json(exp_dir, state)
    if world_size > 1:
        dist.barrier()


def load_full(exp_dir: Path, model, optimizer, lr_scheduler, device, rank=0, world_size=1, deterministic=False):
    exp_dir = Path(exp_dir)

    def _load(p):
        return torch.load(p, map_location=device, weights_only=True)

    sd = _load(exp_dir / "model.pt")
    with torch.no_grad():
        own = model.state_dict()
        for k, v in sd.items():
            own[k].copy_(v)  # in place: parameters are views of flat buffers
    opt_path = exp_dir / ("optimizer.pt" if world_size == 1 else f"optimizer.rank{rank}.pt")
    if optimizer is not None and opt_path.exists():
        optimizer.load_state_dict(_load(opt_path))
    elif optimizer is not None:
        import logging

        logging.getLogger("dtg_b200").warning(
            "%s not found (different world size or node-local disk?): optimizer state starts from zero", opt_path)
    lr_scheduler.load_state_dict(_load(exp_dir / "lr_scheduler.pt"))
    _restore_lr(lr_scheduler)
    rng_path = exp_dir / ("rng.pt" if world_size == 1 else f"rng.rank{rank}.pt")
    if deterministic and rng_path.exists():
        set_rng_state(torch.load(rng_path, weights_only=False))
    return load_state_json(exp_dir)


# -- sharded checkpoints (DCP layout): chapters 04-07 ---------------------------------------------
def save_sharded(exp_dir: Path, shards: dict, lr_scheduler, state, rank, world_size, extra_rank0=None):
    """``shards``: {"model": {key: 1-D local tensor}, "optimizer": {...}} — identical keys and
    local sizes on every rank of the sharding group.  All ranks write between barriers; rank 0
    adds ``lr_scheduler.pt`` and ``state.json`` (reference ``04-...:241-255``)."""
    import torch.distributed.checkpoint as dcp

    exp_dir = Path(exp_dir)
    if world_size > 1:
        dist.barrier()
    sd = {top: {k: _shard_wrap(v, rank, world_size) for k, v in d.items()} for top, d in shards.items()}
    dcp.save(sd, checkpoint_id=str(exp_dir / "checkpoint"))
    if rank == 0:
        _atomic_save(lr_scheduler.state_dict(), exp_dir / "lr_scheduler.pt")
        for name, obj in (extra_rank0 or {}).items():  # small JSON side files (e.g. AdamW step counters): written
            tmp = exp_dir / (name + ".tmp")             # atomically and BEFORE state.json, the completeness marker
            with open(tmp, "w") as fp:
                json.dump(obj, fp)
            os.replace(tmp, exp_dir / name)
        save_state_json(exp_dir, state)
    if world_size > 1:
        dist.barrier()


def load_json_side_file(exp_dir: Path, name: str):
    p = Path(exp_dir) / name
    return json.loads(p.read_text()) if p.exists() else None


def load_sharded(exp_dir: Path, shards: dict, lr_scheduler, device, rank, world_size):
    """In-place load into freshly built local shards (reference ``04-...:135-157``)."""
    import torch.distributed.checkpoint as dcp

    exp_dir = Path(exp_dir)
    wrapped = {top: {k: _shard_wrap(v, rank, world_size) for k, v in d.items()} for top, d in shards.items()}
    dcp.load(wrapped, checkpoint_id=str(exp_dir / "checkpoint"))
    for top, d in shards.items():
        for k, v in d.items():
            w = wrapped[top][k]
            local = w.to_local() if hasattr(w, "to_local") else w
            if local.data_ptr() != v.data_ptr():
                v.copy_(local)
    lr_scheduler.load_state_dict(torch.load(exp_dir / "lr_scheduler.pt", map_location=device, weights_only=True))
    _restore_lr(lr_scheduler)
    return load_state_json(exp_dir)


_MESH_CACHE = {}


def _shard_wrap(local: torch.Tensor, rank: int, world_size: int):
    """1-D local shard -> DTensor(Shard(0)) over the default group (plain tensor if world==1)."""
    if world_size == 1 or not dist.is_initialized():
        return local
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.tensor import DTensor, Shard

    key = (local.device.type, world_size)
    if key not in _MESH_CACHE:
        _MESH_CACHE[key] = init_device_mesh(local.device.type, (world_size,))
    return DTensor.from_local(local, _MESH_CACHE[key], [Shard(0)], run_check=False)
