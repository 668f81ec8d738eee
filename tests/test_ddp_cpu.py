"""DDP + ZeRO-1 engine logic on CPU (gloo, 2 processes): the distributed run must track a
single-process run that sees the concatenated batch."""
import torch

from dist_utils import initial_weights, run_distributed, update_rel_err


def _train(rank, world, parallelism, steps, zero1):
    from distributed_training_guide_b200.engine import TrainEngine
    torch.manual_seed(0)
    eng = TrainEngine.create("debug-llama", parallelism=parallelism, batch_size=2, seq_length=32, device="cpu",
                             lr=1e-3)
    losses = []
    for i in range(steps):
        b = eng.synthetic_batch(seed=i, pinned=False)
        losses.append(float(eng.step(b)))
    sd = {k: v.detach().float().clone() for k, v in eng.model.state_dict().items()}
    return losses, sd, eng.strategy.dp_rank


def _single_reference(steps, world):
    """Single process, same init, global batch = concat of what each dp rank would draw."""
    from distributed_training_guide_b200.engine import TrainEngine

    torch.manual_seed(0)
    eng = TrainEngine.create("debug-llama", parallelism="single", batch_size=2, seq_length=32, device="cpu", lr=1e-3)
    losses = []
    for i in range(steps):
        parts = []
        for r in range(world):
            g = torch.Generator().manual_seed(1000 * i + r)
            parts.append(torch.randint(0, eng.config.vocab_size, (2, 32), generator=g))
        ids = torch.cat(parts)
        losses.append(float(eng.step({"input_ids": ids, "labels": ids.clone()})))
    return losses, {k: v.detach().float().clone() for k, v in eng.model.state_dict().items()}


import pytest


@pytest.mark.parametrize("parallelism", ["ddp", "ddp_allreduce"])
def test_ddp_matches_single_process(parallelism):
    steps, world = 3, 2
    res = run_distributed(_train, world=world, args=(parallelism, steps, True))
    ref_losses, ref_sd = _single_reference(steps, world)
    (l0, sd0, _), (l1, sd1, _) = res
    # replicas stay identical
    import numpy as np

    for k in sd0:
        assert np.array_equal(sd0[k], sd1[k]), k
    # mean of the per-rank losses == loss of the global batch
    for i in range(steps):
        assert abs(0.5 * (l0[i] + l1[i]) - ref_losses[i]) < 2e-2, (i, l0[i], l1[i], ref_losses[i])
    for k in sd0:
        assert np.abs(sd0[k] - ref_sd[k].numpy()).max() < 2e-2, k
    err = update_rel_err(initial_weights(), sd0, {k: v.numpy() for k, v in ref_sd.items()})
    assert err < 0.1, err  # the update itself (not just the weights) matches the single-process run


def _train_accum(rank, world, steps, K):
    from distributed_training_guide_b200.engine import TrainEngine

    torch.manual_seed(0)
    eng = TrainEngine.create("debug-llama", parallelism="ddp", batch_size=1, seq_length=32, device="cpu", lr=1e-3)
    s = eng.strategy
    for i in range(steps):
        for m in range(K):
            g = torch.Generator().manual_seed(10_000 * i + 100 * m + rank)
            ids = torch.randint(0, eng.config.vocab_size, (1, 32), generator=g)
            out = eng.model(**s.prepare_batch({"input_ids": ids, "labels": ids.clone()}))
            with s.grad_sync(eng.model, enabled=(m == K - 1)):   # no_sync() on the first K-1 micro-batches
                s.backward(eng.model, out.loss / K)
        eng.optimizer.step()
        eng.lr_scheduler.step()
        eng.optimizer.zero_grad()
    return {k: v.detach().float().clone() for k, v in eng.model.state_dict().items()}


def test_gradient_accumulation_with_no_sync_matches_big_batch():
    import numpy as np

    from distributed_training_guide_b200.engine import TrainEngine

    steps, world, K = 2, 2, 2
    sd0, sd1 = run_distributed(_train_accum, world=world, args=(steps, K))
    torch.manual_seed(0)
    eng = TrainEngine.create("debug-llama", parallelism="single", batch_size=world * K, seq_length=32, device="cpu",
                             lr=1e-3)
    for i in range(steps):
        parts = []
        for m in range(K):
            for r in range(world):
                g = torch.Generator().manual_seed(10_000 * i + 100 * m + r)
                parts.append(torch.randint(0, eng.config.vocab_size, (1, 32), generator=g))
        ids = torch.cat(parts)
        eng.step({"input_ids": ids, "labels": ids.clone()})
    ref_sd = {k: v.detach().float().numpy() for k, v in eng.model.state_dict().items()}
    for k in sd0:
        assert np.array_equal(sd0[k], sd1[k]), k
    # an optimizer step per micro-batch (or a dropped micro-batch) puts this at ~0.9; bf16 rounding at ~0.03
    err = update_rel_err(initial_weights(), sd0, ref_sd)
    assert err < 0.1, err


def _train_gpt2(rank, world, steps):
    from distributed_training_guide_b200.engine import TrainEngine

    torch.manual_seed(0)
    eng = TrainEngine.create("debug-gpt2", parallelism="ddp", batch_size=2, seq_length=32, device="cpu", lr=1e-3)
    eng.model.eval()   # dropout off: the two layouts draw different masks
    for i in range(steps):
        g = torch.Generator().manual_seed(1000 * i + rank)
        ids = torch.randint(0, eng.config.vocab_size, (2, 32), generator=g)
        eng.step({"input_ids": ids, "labels": ids.clone()})
    return {k: v.detach().float().clone() for k, v in eng.model.state_dict().items()}


def test_gpt2_under_ddp_matches_single_process():
    """The reference's canonical multi-GPU smoke model: tied embeddings, LayerNorm, biases — gradients come from
    autograd (PyTorch ops), the buckets and the sharded optimizer are the same engine as for Llama."""
    import numpy as np

    from distributed_training_guide_b200.engine import TrainEngine

    steps, world = 3, 2
    sd0, sd1 = run_distributed(_train_gpt2, world=world, args=(steps,))
    torch.manual_seed(0)
    eng = TrainEngine.create("debug-gpt2", parallelism="single", batch_size=4, seq_length=32, device="cpu", lr=1e-3)
    eng.model.eval()
    init = {k: v.detach().float().clone().numpy() for k, v in eng.model.state_dict().items()}
    for i in range(steps):
        parts = [torch.randint(0, eng.config.vocab_size, (2, 32), generator=torch.Generator().manual_seed(1000 * i + r))
                 for r in range(world)]
        ids = torch.cat(parts)
        eng.step({"input_ids": ids, "labels": ids.clone()})
    ref_sd = {k: v.detach().float().numpy() for k, v in eng.model.state_dict().items()}
    for k in sd0:
        assert np.array_equal(sd0[k], sd1[k]), k                     # replicas stay identical
    err = update_rel_err(init, sd0, ref_sd)
    assert err < 0.1, err                                            # and they actually trained, like one process
