"""Per-chapter strategies: how the model is placed, parallelised, stepped and checkpointed.

=====================================  =========================================================
chapter                                strategy
=====================================  =========================================================
01-single-gpu                          :class:`SingleDevice`
02-distributed-data-parallel           :class:`DataParallelZero1`   (DDP + ZeRO-1)
04-fully-sharded-data-parallel         :class:`FullyShardedDataParallel`
05-training-llama-405b                 :class:`FullyShardedDataParallel` (+offload, act-ckpt, prefetch)
06-tensor-parallel                     :class:`TensorParallel` (tp = world, dp = 1 per node)
07-2d-parallel                         :class:`TwoDParallel`   (FSDP x TP)
alternative-frameworks/deepspeed       :class:`ZeroConfigured` (ZeRO stage from a ds_config.json)
=====================================  =========================================================
"""
from __future__ import annotations

import contextlib
import logging
from pathlib import Path

import torch
import torch.distributed as dist

from ..models import build_model
from ..utils import ckpt as ckpt_utils
from ..utils.lr import cosine_schedule
from . import bootstrap
from .flat import build_groups
from .optim import FlatAdamW

LOGGER = logging.getLogger("dtg_b200")


class Strategy:
    chapter = "base"
    log_rank_prefix = True
    show_progress = True

    def __init__(self, args=None):
        self.args = args
        self.env = None
        self.dp_size = 1
        self.dp_rank = 0
        self.groups = None

    # -- process / device -----------------------------------------------------------------
    def setup(self, args):
        dev = getattr(args, "device", None)
        self.env = bootstrap.init_distributed(device_type=dev, local_rank=getattr(args, "local_rank", None))
        self.dp_size, self.dp_rank = self.env.world_size, self.env.rank
        return self.env

    def barrier(self):
        if self.env.distributed:
            dist.barrier()

    def data_guard(self):
        if self.env.distributed:
            return bootstrap.rank0_first()
        return contextlib.nullcontext()

    def teardown(self):
        pass

    # -- model / optimizer -----------------------------------------------------------------
    def dtype(self):
        return torch.bfloat16

    def build_model(self, args, config):
        raise NotImplementedError

    def num_parameters(self, model):
        return sum(p.numel() for p in model.parameters())

    def build_optimizer(self, args, model, lr):
        return FlatAdamW(self.groups, lr=lr)

    def build_lr_scheduler(self, args, optimizer, lr):
        return cosine_schedule(optimizer, lr)

    def build_tracker(self, args, exp_dir, resumed, config):
        from ..utils.tracking import build_tracker

        return build_tracker(args, self.env, exp_dir, resumed, config)

    # -- step hooks ---------------------------------------------------------------------------
    def pre_step(self, model):
        pass

    def prepare_batch(self, batch):
        return batch

    def grad_sync(self, model, enabled=True):
        return contextlib.nullcontext()

    def backward(self, model, loss):
        loss.backward()

    # -- checkpoints ----------------------------------------------------------------------------
    def make_experiment_dir(self, exp_dir: Path):
        if self.env.rank == 0:
            LOGGER.info("Creating experiment root directory")
            exp_dir.mkdir(parents=True, exist_ok=True)

    def save_checkpoint(self, exp_dir, model, optimizer, lr_scheduler, state):
        ckpt_utils.save_full(exp_dir, model, optimizer, lr_scheduler, state, rank=self.env.rank,
                             world_size=self.env.world_size if self.env.distributed else 1,
                             deterministic=getattr(self.args, "deterministic", False))

    def load_checkpoint(self, exp_dir, model, optimizer, lr_scheduler):
        return ckpt_utils.load_full(exp_dir, model, optimizer, lr_scheduler, self.env.device, rank=self.env.rank,
                                    world_size=self.env.world_size if self.env.distributed else 1,
                                    deterministic=getattr(self.args, "deterministic", False))


class SingleDevice(Strategy):
    """Chapter 01: one device, bf16, flat AdamW (reference ``01-single-gpu/train_llm.py``).
    Runs on CPU too (BASELINE.json config 01: GPT-2 124M plumbing)."""

    chapter = "01-single-gpu"
    log_rank_prefix = False

    def setup(self, args):
        dev = getattr(args, "device", None) or ("cuda" if torch.cuda.is_available() else "cpu")
        device = torch.device("cuda:0" if dev == "cuda" else dev)
        if device.type == "cuda":
            torch.cuda.set_device(device)
        self.env = bootstrap.DistEnv(rank=0, local_rank=0, world_size=1, device=device, distributed=False)
        return self.env

    def build_model(self, args, config):
        model = build_model(config, dtype=self.dtype(), device=self.env.device)
        self.groups = build_groups(model, self.env.device, self.dtype())
        return model
