"""Experiment tracking (wandb) in the three topologies of the reference
(``related-topics/wandb-configurations/README.md:5-63``; rank-0 usage in
``alternative-frameworks/deepspeed/train_llm.py:110-124,185-186``):

  rank0        one run, only the global rank 0 logs
  local-rank0  one run per node (grouped by experiment)
  all          one run per rank (grouped by experiment, per-rank dir ``rank-{r}``)

There is no network on the GPU boxes, so the tracker runs ``mode=offline`` unless
``WANDB_MODE`` says otherwise, and degrades to a JSONL file when wandb is not installed.
"""
from __future__ import annotations

import json
import os
from pathlib import Path


class JsonlTracker:
    def __init__(self, path: Path):
        path.parent.mkdir(parents=True, exist_ok=True)
        self.fp = open(path, "a")

    def log(self, info, step):
        self.fp.write(json.dumps({"step": step, **info}) + "\n")
        self.fp.flush()

    def finish(self):
        self.fp.close()


class WandbTracker:
    def __init__(self, run):
        self.run = run

    def log(self, info, step):
        self.run.log(info, step=step)

    def finish(self):
        self.run.finish()


def build_tracker(args, env, exp_dir, resumed, config):
    mode = getattr(args, "wandb", "off")
    if mode == "off":
        return None
    active = {"rank0": env.rank == 0, "local-rank0": env.local_rank == 0, "all": True}[mode]
    if not active:
        return None
    run_dir = Path(exp_dir) if exp_dir is not None else Path(args.save_dir)
    if mode == "all":
        run_dir = run_dir / f"rank-{env.rank}"
    run_dir.mkdir(parents=True, exist_ok=True)
    try:
        import wandb

        os.environ.setdefault("WANDB_MODE", "offline")
        run = wandb.init(
            project="distributed-training-guide-b200", dir=str(run_dir),
            name=args.experiment_name if mode == "rank0" else f"{args.experiment_name}-rank{env.rank}",
            id=(args.experiment_name if mode == "rank0" else None),
            group=None if mode == "rank0" else args.experiment_name,
            resume="must" if (resumed and mode == "rank0") else None,
            save_code=False,
            config={"args": vars(args), "model": config.to_dict(), "world_size": env.world_size},
        )
        return WandbTracker(run)
    except Exception:
        return JsonlTracker(run_dir / "metrics.jsonl")
