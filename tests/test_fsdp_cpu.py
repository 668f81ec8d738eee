"""FSDP engine schedule on CPU (gloo, 2 processes) vs a single-process run on the concatenated batch."""
import numpy as np
import pytest
import torch

from dist_utils import initial_weights, run_distributed, update_rel_err
from test_ddp_cpu import _single_reference


def _train(rank, world, steps, ckpt_act, offload=False):
    from distributed_training_guide_b200.engine import TrainEngine

    torch.manual_seed(0)
    eng = TrainEngine.create("debug-llama", parallelism="fsdp", batch_size=2, seq_length=32, device="cpu", lr=1e-3,
                             checkpoint_activations=ckpt_act, cpu_offload=offload)
    losses = [float(eng.step(eng.synthetic_batch(seed=i, pinned=False))) for i in range(steps)]
    sd = eng.strategy.engine.full_state_dict()
    return losses, {k: v.float() for k, v in sd.items()}


def _check(ckpt_act, offload=False):
    steps, world = 3, 2
    res = run_distributed(_train, world=world, args=(steps, ckpt_act, offload))
    ref_losses, ref_sd = _single_reference(steps, world)
    (l0, sd0), (l1, sd1) = res
    for i in range(steps):
        assert abs(0.5 * (l0[i] + l1[i]) - ref_losses[i]) < 2e-2, (i, l0[i], l1[i], ref_losses[i])
    for k in sd0:
        assert np.array_equal(sd0[k], sd1[k]), k
        assert np.abs(sd0[k] - ref_sd[k].numpy()).max() < 2e-2, k
    err = update_rel_err(initial_weights(), sd0, {k: v.numpy() for k, v in ref_sd.items()})
    assert err < 0.1, err  # the update itself (not just the weights) matches the single-process run


def test_fsdp_matches_single_process():
    _check(False)


def test_fsdp_with_activation_checkpointing():
    _check(True)


def test_fsdp_cpu_offload_flag():
    _check(False, offload=True)


def test_sharded_checkpoint_roundtrip_and_consolidate(tmp_path):
    """chapter 04 under torchrun (gloo): DCP layout, resume, and the consolidation tool."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    script = root / "04-fully-sharded-data-parallel" / "train_llm.py"

    def run(max_steps):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
               "--nproc-per-node", "2", str(script), "-d", "synthetic", "-m", "debug-llama", "-s", "32", "-b", "2",
               "--num-samples", "32", "--log-freq", "1", "--device", "cpu", "--save-dir", str(tmp_path), "-e", "exp",
               "--ckpt-freq", "2", "--lr", "1e-3", "--max-steps", str(max_steps)]
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(script.parent), timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        return r.stderr

    run(2)
    exp = tmp_path / "exp"
    names = {p.name for p in exp.iterdir()}
    assert {"checkpoint", "state.json", "lr_scheduler.pt", "rank-0", "rank-1"} <= names, names
    ck = {p.name for p in (exp / "checkpoint").iterdir()}
    assert ".metadata" in ck and any(n.endswith(".distcp") for n in ck), ck
    assert json.loads((exp / "state.json").read_text())["global_step"] == 2
    log = run(4)
    assert "Resumed=True" in log
    from distributed_training_guide_b200.tools.consolidate import consolidate

    out = consolidate(str(exp), "debug-llama", world=2)
    sd = torch.load(out, weights_only=True)
    assert "model.layers.1.mlp.down_proj.weight" in sd and sd["lm_head.weight"].shape == (1024, 256)


def _load_pretrained(rank, world):
    """rank 0 holds a 'pretrained' state dict; load_into_fsdp distributes it group by group."""
    from distributed_training_guide_b200.engine import TrainEngine
    from distributed_training_guide_b200.models import build_model, get_config
    from distributed_training_guide_b200.tools.load_hf import load_into_fsdp

    eng = TrainEngine.create("debug-llama", parallelism="fsdp", batch_size=2, seq_length=32, device="cpu", lr=1e-3, seed=5)
    torch.manual_seed(123)
    src = build_model(get_config("debug-llama"), dtype=torch.bfloat16, device="cpu")
    src.init_weights(seed=999)  # different from the engine's own init
    sd = src.state_dict() if rank == 0 else None
    load_into_fsdp(eng.strategy.engine, (lambda name: sd[name]) if rank == 0 else None)
    full = eng.strategy.engine.full_state_dict()
    ref_sd = src.state_dict()
    return max(float((full[k].float() - ref_sd[k].float()).abs().max()) for k in ref_sd)


def test_pretrained_load_and_broadcast():
    res = run_distributed(_load_pretrained, world=2)
    assert all(r == 0.0 for r in res), res


def _fsdp_accum(rank, world, steps, K):
    from distributed_training_guide_b200.engine import TrainEngine

    torch.manual_seed(0)
    eng = TrainEngine.create("debug-llama", parallelism="fsdp", batch_size=1, seq_length=32, device="cpu", lr=1e-3)
    s = eng.strategy
    for i in range(steps):
        for m in range(K):
            g = torch.Generator().manual_seed(10_000 * i + 100 * m + rank)
            ids = torch.randint(0, eng.config.vocab_size, (1, 32), generator=g)
            s.pre_step(eng.model)
            out = eng.model(**s.prepare_batch({"input_ids": ids, "labels": ids.clone()}))
            with s.grad_sync(eng.model, enabled=(m == K - 1)):
                s.backward(eng.model, out.loss / K)
        eng.optimizer.step()
        eng.lr_scheduler.step()
        eng.optimizer.zero_grad()
    full = s.engine.full_state_dict()
    return {k: v.detach().float().clone() for k, v in full.items()}


def test_fsdp_gradient_accumulation_matches_big_batch():
    import numpy as np

    from distributed_training_guide_b200.engine import TrainEngine

    steps, world, K = 2, 2, 2
    sd0, sd1 = run_distributed(_fsdp_accum, world=world, args=(steps, K))
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


def _fsdp_gpt2(rank, world, steps):
    from distributed_training_guide_b200.engine import TrainEngine

    torch.manual_seed(0)
    eng = TrainEngine.create("debug-gpt2", parallelism="fsdp", batch_size=2, seq_length=32, device="cpu", lr=1e-3)
    eng.model.eval()   # dropout off: sharded and single-process runs draw different masks
    init = {k: v.detach().float().clone() for k, v in eng.strategy.engine.full_state_dict().items()}
    for i in range(steps):
        g = torch.Generator().manual_seed(1000 * i + rank)
        ids = torch.randint(0, eng.config.vocab_size, (2, 32), generator=g)
        eng.step({"input_ids": ids, "labels": ids.clone()})
    return init, {k: v.detach().float().clone() for k, v in eng.strategy.engine.full_state_dict().items()}


def test_gpt2_under_fsdp_matches_single_process():
    """GPT-2 (tied lm_head, biases, LayerNorm; gradients from autograd) through the sharded engine: the reference's
    smoke model for its FSDP chapter."""
    from distributed_training_guide_b200.engine import TrainEngine

    steps, world = 3, 2
    (init0, sd0), (init1, sd1) = run_distributed(_fsdp_gpt2, world=world, args=(steps,))
    torch.manual_seed(0)
    eng = TrainEngine.create("debug-gpt2", parallelism="single", batch_size=4, seq_length=32, device="cpu", lr=1e-3)
    eng.model.eval()
    ref_init = {k: v.detach().float().clone().numpy() for k, v in eng.model.state_dict().items()}
    for k in init0:   # same seed -> the sharded build materialises exactly the single-process weights
        assert np.array_equal(init0[k], ref_init[k]), k
    for i in range(steps):
        parts = [torch.randint(0, eng.config.vocab_size, (2, 32), generator=torch.Generator().manual_seed(1000 * i + r))
                 for r in range(world)]
        ids = torch.cat(parts)
        eng.step({"input_ids": ids, "labels": ids.clone()})
    ref_sd = {k: v.detach().float().numpy() for k, v in eng.model.state_dict().items() if k in sd0}
    for k in sd0:
        assert np.array_equal(sd0[k], sd1[k]), k
    err = update_rel_err({k: ref_init[k] for k in ref_sd}, sd0, ref_sd)
    assert err < 0.1, err


def _tied_llama(rank, world, model_dir, parallelism, steps):
    from distributed_training_guide_b200.engine import TrainEngine

    torch.manual_seed(0)
    eng = TrainEngine.create(model_dir, parallelism=parallelism, batch_size=2, seq_length=32, device="cpu", lr=1e-3)
    for i in range(steps):
        g = torch.Generator().manual_seed(1000 * i + rank)
        ids = torch.randint(0, eng.config.vocab_size, (2, 32), generator=g)
        eng.step({"input_ids": ids, "labels": ids.clone()})
    sd = eng.strategy.engine.full_state_dict() if parallelism == "fsdp" else eng.model.state_dict()
    return {k: v.detach().float().clone() for k, v in sd.items()}


@pytest.mark.parametrize("parallelism", ["ddp", "fsdp"])
def test_llama_with_tied_embeddings(tmp_path, parallelism):
    """Llama-3.2-style tie_word_embeddings: the lm_head and the embedding are one parameter living in the embed group."""
    import json

    from distributed_training_guide_b200.engine import TrainEngine
    from distributed_training_guide_b200.models import get_config
    from distributed_training_guide_b200.models.configs import to_hf_config_dict

    d = to_hf_config_dict(get_config("debug-llama"))
    d["tie_word_embeddings"] = True
    (tmp_path / "config.json").write_text(json.dumps(d))
    steps, world = 2, 2
    sd0, sd1 = run_distributed(_tied_llama, world=world, args=(str(tmp_path), parallelism, steps))
    torch.manual_seed(0)
    eng = TrainEngine.create(str(tmp_path), parallelism="single", batch_size=4, seq_length=32, device="cpu", lr=1e-3)
    assert eng.model.lm_head.weight is eng.model.model.embed_tokens.weight
    init = {k: v.detach().float().clone().numpy() for k, v in eng.model.state_dict().items()}
    for i in range(steps):
        parts = [torch.randint(0, eng.config.vocab_size, (2, 32), generator=torch.Generator().manual_seed(1000 * i + r))
                 for r in range(world)]
        ids = torch.cat(parts)
        eng.step({"input_ids": ids, "labels": ids.clone()})
    ref_sd = {k: v.detach().float().numpy() for k, v in eng.model.state_dict().items() if k in sd0}
    for k in sd0:
        assert np.array_equal(sd0[k], sd1[k]), k
    err = update_rel_err({k: init[k] for k in ref_sd}, sd0, ref_sd)
    assert err < 0.1, err
