"""ZeRO-config driven strategy (the guide's DeepSpeed alternative, reference
``alternative-frameworks/deepspeed/train_llm.py`` + ``ds_config.json``).

The JSON keys understood are the ones the reference config sets: ``train_micro_batch_size_per_gpu``,
``gradient_accumulation_steps``, ``optimizer`` (AdamW params), ``scheduler`` (WarmupCosineLR), ``bf16``,
``zero_optimization.stage`` and ``offload_optimizer.device``.  The stage selects an engine of this
repository instead of the DeepSpeed runtime.
"""
from __future__ import annotations

import json
import logging

from ..utils.lr import cosine_schedule, warmup_cosine_schedule
from .strategies import DataParallelZero1, FullyShardedDataParallel, Strategy

DEFAULT_CONFIG = {
    "train_micro_batch_size_per_gpu": 1,
    "optimizer": {"type": "AdamW", "params": {"lr": 3e-5}},
    "scheduler": {"type": "WarmupCosineLR", "params": {"total_num_steps": 1000, "warmup_num_steps": 0,
                                                        "warmup_min_ratio": 0.0, "cos_min_ratio": 1e-2}},
    "bf16": {"enabled": True},
    "zero_optimization": {"stage": 3},
}


KNOWN_KEYS = {"train_micro_batch_size_per_gpu", "gradient_accumulation_steps", "optimizer", "scheduler", "bf16",
              "zero_optimization", "train_batch_size", "steps_per_print", "wall_clock_breakdown"}


def load_zero_config(path):
    if not path:
        return dict(DEFAULT_CONFIG)
    with open(path) as fp:
        cfg = json.load(fp)
    unknown = sorted(set(cfg) - KNOWN_KEYS)
    if unknown:
        logging.getLogger("dtg_b200").warning(f"{path}: keys not understood by this front end are ignored: {unknown}")
    if not cfg.get("bf16", {"enabled": True}).get("enabled", True):
        raise ValueError(f"{path}: bf16.enabled must be true (every engine here computes in bf16)")
    if cfg.get("optimizer", {}).get("type", "AdamW") not in ("AdamW", "Adam"):
        raise ValueError(f"{path}: optimizer.type {cfg['optimizer']['type']!r} is not supported (AdamW)")
    out = dict(DEFAULT_CONFIG)
    out.update(cfg)
    return out


class ZeroConfigured(Strategy):
    """Delegates to DDP(+ZeRO-1) or FSDP according to ``zero_optimization.stage``."""

    chapter = "deepspeed"

    def __init__(self, args):
        super().__init__(args)
        self.cfg = load_zero_config(getattr(args, "deepspeed_config", None))
        zo = self.cfg.get("zero_optimization", {})
        stage = int(zo.get("stage", 0))
        # the JSON takes over batch size / lr / accumulation, like deepspeed.initialize does
        args.batch_size = int(self.cfg.get("train_micro_batch_size_per_gpu", args.batch_size))
        args.grad_accum_steps = int(self.cfg.get("gradient_accumulation_steps", getattr(args, "grad_accum_steps", 1)))
        opt = self.cfg.get("optimizer", {}).get("params", {})
        args.lr = float(opt.get("lr", args.lr))
        self.opt_params = opt
        if getattr(args, "wandb", "off") == "off":
            args.wandb = "rank0"  # the reference's deepspeed script logs to wandb from rank 0
        args.cpu_offload = zo.get("offload_optimizer", {}).get("device", "none") == "cpu"
        if stage >= 3:
            self.inner = FullyShardedDataParallel(args)
        else:
            self.inner = DataParallelZero1(args, zero1=stage >= 1)
        self.stage = stage

    def __getattr__(self, name):  # everything not overridden is the inner strategy's
        return getattr(self.__dict__["inner"], name)

    def setup(self, args):
        env = self.inner.setup(args)
        self.env, self.dp_size, self.dp_rank = env, self.inner.dp_size, self.inner.dp_rank
        return env

    def build_model(self, args, config):
        return self.inner.build_model(args, config)

    def num_parameters(self, model):
        return self.inner.num_parameters(model)

    def build_optimizer(self, args, model, lr):
        opt = self.inner.build_optimizer(args, model, lr)
        for pg in opt.param_groups:
            if "betas" in self.opt_params:
                pg["betas"] = tuple(self.opt_params["betas"])
            if "eps" in self.opt_params:
                pg["eps"] = float(self.opt_params["eps"])
            if "weight_decay" in self.opt_params:
                pg["weight_decay"] = float(self.opt_params["weight_decay"])
        return opt

    def build_lr_scheduler(self, args, optimizer, lr):
        sch = self.cfg.get("scheduler", {})
        if sch.get("type") == "WarmupCosineLR":
            p = sch.get("params", {})
            return warmup_cosine_schedule(optimizer, int(p.get("total_num_steps", 1000)),
                                          int(p.get("warmup_num_steps", 0)), float(p.get("warmup_min_ratio", 0.0)),
                                          float(p.get("cos_min_ratio", 1e-2)))
        return cosine_schedule(optimizer, lr)

    def pre_step(self, model):
        return self.inner.pre_step(model)

    def grad_sync(self, model, enabled=True):
        return self.inner.grad_sync(model, enabled)

    def backward(self, model, loss):
        return self.inner.backward(model, loss)

    def barrier(self):
        return self.inner.barrier()

    def data_guard(self):
        return self.inner.data_guard()

    def make_experiment_dir(self, exp_dir):
        return self.inner.make_experiment_dir(exp_dir)

    def save_checkpoint(self, *a, **k):
        return self.inner.save_checkpoint(*a, **k)

    def load_checkpoint(self, *a, **k):
        return self.inner.load_checkpoint(*a, **k)

    def build_tracker(self, args, exp_dir, resumed, config):
        return Strategy.build_tracker(self.inner, args, exp_dir, resumed, config)

    def teardown(self):
        return self.inner.teardown()

    def check_health(self):
        return self.inner.check_health()
