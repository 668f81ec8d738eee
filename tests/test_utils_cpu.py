"""CLI surface, data pipeline, timers, LR rules, log-record schema, checkpoint layout + resume."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent

from distributed_training_guide_b200.utils import data as D
from distributed_training_guide_b200.utils.cli import CHAPTER_EXTRAS, get_parser
from distributed_training_guide_b200.utils.lr import scale_lr, warmup_cosine_schedule
from distributed_training_guide_b200.utils.timers import LocalTimer


def test_cli_flag_matrix():
    base = ["-d", "synthetic", "-m", "gpt2"]
    a = get_parser("01-single-gpu").parse_args(base)
    assert (a.save_dir, a.seed, a.num_epochs, a.lr, a.batch_size, a.log_freq, a.ckpt_freq, a.seq_length) == (
        "../outputs", 0, 100, 3e-5, 1, 10, 500, 1024)
    assert a.experiment_name is None
    with pytest.raises(SystemExit):
        get_parser("01-single-gpu").parse_args(["-m", "gpt2"])  # dataset is required
    a4 = get_parser("04-fully-sharded-data-parallel").parse_args(base + ["--cpu-offload"])
    assert a4.cpu_offload is True
    a5 = get_parser("05-training-llama-405b").parse_args(base + ["--checkpoint-activations", "--prefetch-layers"])
    assert a5.checkpoint_activations and a5.prefetch_layers and not a5.cpu_offload
    assert get_parser("07-2d-parallel").parse_args(base + ["-tp", "4"]).tensor_parallel == 4
    assert get_parser("07-2d-parallel").parse_args(base).tensor_parallel == 8
    with pytest.raises(SystemExit):
        get_parser("deepspeed", require_experiment=True).parse_args(base)  # -e required there
    ds = get_parser("deepspeed", require_experiment=True).parse_args(base + ["-e", "x", "--local_rank", "3"])
    assert ds.local_rank == 3
    assert set(CHAPTER_EXTRAS) >= {"01-single-gpu", "02-distributed-data-parallel", "04-fully-sharded-data-parallel",
                                   "05-training-llama-405b", "06-tensor-parallel", "07-2d-parallel", "deepspeed"}


def test_synthetic_and_text_datasets(tmp_path):
    from types import SimpleNamespace

    from distributed_training_guide_b200.models import get_config

    cfg = get_config("debug-llama")
    args = SimpleNamespace(dataset_name="synthetic", dataset_subset=None, model_name="debug-llama", seq_length=64, seed=1,
                           batch_size=2, num_samples=10)
    ds = D.load_and_preprocess_data(args, cfg)
    assert len(ds) == 10 and ds[0]["input_ids"].shape == (64,)
    assert torch.equal(ds[0]["input_ids"], ds[0]["labels"]) and ds[0]["attention_mask"].all()
    assert torch.equal(ds[3]["input_ids"], D.load_and_preprocess_data(args, cfg)[3]["input_ids"])  # deterministic
    # seq_length clamp rule
    args.seq_length = 99999
    assert D.clamp_seq_length(args.seq_length, cfg) == min(1024, cfg.max_position_embeddings)
    # text file -> byte tokens -> chunks, remainder dropped
    p = tmp_path / "corpus.txt"
    p.write_text("hello world\n" * 50)
    args.dataset_name, args.seq_length = str(p), 32
    chunks = D.load_and_preprocess_data(args, cfg)
    total = 50 * (len("hello world\n".encode()) + 1)
    assert len(chunks) == total // 32 and chunks[0]["input_ids"].max() < cfg.vocab_size


def test_distributed_sampler_partitions():
    ds = D.SyntheticTokens(32, 8, 100, seed=0)
    seen = []
    for r in range(4):
        dl = D.build_dataloader(ds, 2, dp_size=4, dp_rank=r, seed=0, distributed=True, num_workers=0, pin_memory=False)
        dl.sampler.set_epoch(0)
        assert len(dl) == 4
        seen += [tuple(row.tolist()) for b in dl for row in b["input_ids"]]
    assert len(set(seen)) == 32  # disjoint cover


def test_lr_rules_and_schedules():
    assert scale_lr(1e-3, 8, "linear") == pytest.approx(8e-3)
    assert scale_lr(1e-3, 4, "sqrt") == pytest.approx(2e-3)
    assert scale_lr(1e-3, 4, "none") == 1e-3
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sch = warmup_cosine_schedule(opt, total_num_steps=10, warmup_num_steps=2, warmup_min_ratio=0.0, cos_min_ratio=0.01)
    lrs = []
    for _ in range(12):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step(); sch.step()
    assert lrs[0] == 0.0 and lrs[2] == pytest.approx(1.0) and lrs[-1] == pytest.approx(0.01)


def test_local_timer_cpu():
    t = LocalTimer(torch.device("cpu"))
    for _ in range(3):
        with t:
            sum(range(1000))
    assert len(t.measurements) == 3 and t.avg_elapsed_ms() >= 0
    t.reset()
    assert t.measurements == []


SCHEMA = {"global_step", "lr", "running_loss", "epoch", "epoch_progress", "num_batches_remaining", "total_gb",
          "curr_alloc_gb", "peak_alloc_gb", "curr_resv_gb", "peak_resv_gb", "tokens_per_s", "time/total", "time/data",
          "time/forward", "time/backward", "time/update"}


def _run_ch01(tmp, extra):
    cmd = [sys.executable, str(ROOT / "01-single-gpu" / "train_llm.py"), "-d", "synthetic", "-m", "debug-llama", "-s", "32",
           "-b", "2", "--num-samples", "24", "--log-freq", "1", "--device", "cpu", "--save-dir", str(tmp), "-e", "exp",
           "--ckpt-freq", "3", "--lr", "1e-3"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(ROOT / "01-single-gpu"), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    recs = []
    for line in r.stderr.splitlines():
        if "INFO:{" in line:
            recs.append(eval(line.split("INFO:", 1)[1]))  # the record is a printed dict, as in the reference
    return recs, r.stderr


def test_chapter01_log_schema_checkpoint_layout_and_resume(tmp_path):
    full, _ = _run_ch01(tmp_path / "a", ["--max-steps", "6"])
    assert set(full[0]) == SCHEMA
    exp = tmp_path / "a" / "exp"
    assert {p.name for p in exp.iterdir()} >= {"state.json", "model.pt", "optimizer.pt", "lr_scheduler.pt"}
    st = json.loads((exp / "state.json").read_text())
    assert set(st) == {"epoch", "global_step", "epoch_step", "running_loss"} and st["global_step"] == 6
    sd = torch.load(exp / "model.pt", weights_only=True)
    assert "model.layers.0.self_attn.q_proj.weight" in sd and "lm_head.weight" in sd  # plain HF names
    # interrupted at 3 then resumed to 6 must reproduce the uninterrupted losses
    first, _ = _run_ch01(tmp_path / "b", ["--max-steps", "3"])
    second, log = _run_ch01(tmp_path / "b", ["--max-steps", "6"])
    assert "Resumed=True" in log
    a = [r["running_loss"] for r in full]
    b = [r["running_loss"] for r in first + second]
    assert len(b) == 6 and all(abs(x - y) < 2e-3 for x, y in zip(a, b)), (a, b)


def test_zero_config_selects_engine(tmp_path):
    import json
    from types import SimpleNamespace

    from distributed_training_guide_b200.parallel.strategies import DataParallelZero1, FullyShardedDataParallel
    from distributed_training_guide_b200.parallel.zero_config import ZeroConfigured, load_zero_config

    def args_for(cfg):
        p = tmp_path / "ds.json"
        p.write_text(json.dumps(cfg))
        return SimpleNamespace(deepspeed_config=str(p), batch_size=4, lr=1.0, wandb="off", grad_accum_steps=1)

    a = args_for({"train_micro_batch_size_per_gpu": 2, "optimizer": {"type": "AdamW", "params": {"lr": 5e-4}},
                  "zero_optimization": {"stage": 3, "offload_optimizer": {"device": "cpu"}}, "made_up_key": 1})
    s3 = ZeroConfigured(a)
    assert isinstance(s3.inner, FullyShardedDataParallel) and a.batch_size == 2 and a.lr == 5e-4 and a.cpu_offload
    a = args_for({"zero_optimization": {"stage": 1}})
    s1 = ZeroConfigured(a)
    assert isinstance(s1.inner, DataParallelZero1) and s1.inner.zero1 and not a.cpu_offload
    s0 = ZeroConfigured(args_for({"zero_optimization": {"stage": 0}}))
    assert isinstance(s0.inner, DataParallelZero1) and not s0.inner.zero1
    with pytest.raises(ValueError):
        load_zero_config(args_for({"bf16": {"enabled": False}}).deepspeed_config)
