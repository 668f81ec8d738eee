"""FSDP engine schedule on CPU (gloo, 2 processes) vs a single-process run on the concatenated batch."""
import numpy as np
import torch

from dist_utils import run_distributed
from test_ddp_cpu import _single_reference


def _train(rank, world, steps, ckpt_act):
    from distributed_training_guide_b200.engine import TrainEngine

    torch.manual_seed(0)
    eng = TrainEngine.create("debug-llama", parallelism="fsdp", batch_size=2, seq_length=32, device="cpu", lr=1e-3,
                             checkpoint_activations=ckpt_act)
    losses = [float(eng.step(eng.synthetic_batch(seed=i, pinned=False))) for i in range(steps)]
    sd = eng.strategy.engine.full_state_dict()
    return losses, {k: v.float() for k, v in sd.items()}


def _check(ckpt_act):
    steps, world = 3, 2
    res = run_distributed(_train, world=world, args=(steps, ckpt_act))
    ref_losses, ref_sd = _single_reference(steps, world)
    (l0, sd0), (l1, sd1) = res
    for i in range(steps):
        assert abs(0.5 * (l0[i] + l1[i]) - ref_losses[i]) < 2e-2, (i, l0[i], l1[i], ref_losses[i])
    for k in sd0:
        assert np.array_equal(sd0[k], sd1[k]), k
        assert np.abs(sd0[k] - ref_sd[k].numpy()).max() < 2e-2, k


def test_fsdp_matches_single_process():
    _check(False)


def test_fsdp_with_activation_checkpointing():
    _check(True)
