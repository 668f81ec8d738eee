"""Resume must continue the SAME trajectory: train 4 steps straight vs. 2 steps -> checkpoint -> fresh engine ->
load -> 2 more steps, and compare the weights (and the AdamW step counters).  Covers the sharded engines whose
resume tests used to check only ``Resumed=True``: FSDP with and without ``--cpu-offload`` (the host master copy
must be refreshed from the loaded shards), pure tensor parallel and 2-D (AdamW bias-correction steps persisted)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from dist_utils import run_distributed


def _local_params(eng):
    e = getattr(eng.strategy, "engine", None)
    if e is not None and hasattr(e, "shards"):
        return [s.param.detach().float().clone() for s in e.shards]
    return [g.param.detach().float().clone() for g in eng.strategy.groups]


def _steps(eng):
    e = getattr(eng.strategy, "engine", None)
    if e is not None and hasattr(e, "optimizer_steps"):
        return sorted(e.optimizer_steps().items())
    return sorted((g.name, int(eng.optimizer.state[g.param]["step"])) for g in eng.strategy.groups)


def _resume(rank, world, par, offload, tp, tmp):
    from distributed_training_guide_b200.engine import TrainEngine

    def make():
        torch.manual_seed(0)
        return TrainEngine.create("debug-llama", parallelism=par, batch_size=2, seq_length=32, device="cpu", lr=1e-2,
                                  cpu_offload=offload, tensor_parallel=tp)

    eng = make()
    init = _local_params(eng)
    batches = [eng.synthetic_batch(seed=i, pinned=False) for i in range(4)]
    for b in batches[:2]:
        eng.step(b)
    exp = Path(tmp)
    if rank == 0:
        exp.mkdir(parents=True, exist_ok=True)
    eng.strategy.barrier()
    eng.strategy.save_checkpoint(exp, eng.model, eng.optimizer, eng.lr_scheduler,
                                 {"epoch": 0, "global_step": 2, "epoch_step": 2, "running_loss": 0.0})
    for b in batches[2:]:
        eng.step(b)
    straight, straight_steps = _local_params(eng), _steps(eng)

    eng2 = make()
    st = eng2.strategy.load_checkpoint(exp, eng2.model, eng2.optimizer, eng2.lr_scheduler)
    assert st["global_step"] == 2
    for b in batches[2:]:
        eng2.step(b)
    resumed, resumed_steps = _local_params(eng2), _steps(eng2)
    assert straight_steps == resumed_steps, (straight_steps, resumed_steps)
    diff = max(float((a - b).abs().max()) for a, b in zip(straight, resumed))
    moved = max(float((a - b).abs().max()) for a, b in zip(straight, init))
    return diff, moved


@pytest.mark.parametrize("par,offload,tp", [("fsdp", False, None), ("fsdp", True, None), ("tp", False, 2),
                                            ("2d", False, 2), ("2d", True, 2)])
def test_resume_continues_the_uninterrupted_trajectory(tmp_path, par, offload, tp):
    world = 4 if par == "2d" else 2
    res = run_distributed(_resume, world=world, args=(par, offload, tp, str(tmp_path / "exp")), timeout=600)
    for diff, moved in res:
        assert moved > 1e-2, moved          # training did move the weights ...
        assert diff <= 1e-6 * max(1.0, moved) + 1e-7, (diff, moved)   # ... and the resumed run lands on the same ones
