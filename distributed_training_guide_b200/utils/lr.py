"""LR schedule and LR-scaling rules.

Reference: ``CosineAnnealingLR(optimizer, T_max=1000, eta_min=lr*1e-2)`` (``01-single-gpu/
train_llm.py:75-78``; periodic — the LR climbs again after step 1000);
DeepSpeed ``WarmupCosineLR`` (``alternative-frameworks/deepspeed/ds_config.json:9-16``); linear
and square-root batch-size scaling rules (``related-topics/effective-batch-size-and-lr``).
"""
from __future__ import annotations

import math

import torch


def cosine_schedule(optimizer, lr, t_max=1000, eta_min_ratio=1e-2):
    return torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=t_max, eta_min=lr * eta_min_ratio)


def warmup_cosine_schedule(optimizer, total_num_steps, warmup_num_steps=0, warmup_min_ratio=0.0, cos_min_ratio=1e-2):
    """DeepSpeed's WarmupCosineLR: linear warm-up from ``warmup_min_ratio`` then cosine to ``cos_min_ratio``."""

    def f(step):
        if warmup_num_steps > 0 and step < warmup_num_steps:
            return warmup_min_ratio + (1 - warmup_min_ratio) * step / warmup_num_steps
        prog = min(1.0, (step - warmup_num_steps) / max(1, total_num_steps - warmup_num_steps))
        return cos_min_ratio + (1 - cos_min_ratio) * 0.5 * (1 + math.cos(math.pi * prog))

    return torch.optim.lr_scheduler.LambdaLR(optimizer, f)


def scale_lr(lr: float, dp_size: int, rule: str = "none") -> float:
    """'linear': lr * N (SGD-style);  'sqrt': lr * sqrt(N) (Adam-style);  'none': unchanged."""
    if rule == "linear":
        return lr * dp_size
    if rule == "sqrt":
        return lr * math.sqrt(dp_size)
    return lr
