"""Reference arm: run the UNMODIFIED LambdaLabsML/distributed-training-guide chapter script
from ``baseline/_ref`` through its own CLI / stock code path and time it.

``python -m pip install --no-index ... --target baseline/_ref /root/reference`` fails
("Neither 'setup.py' nor 'pyproject.toml' found": the reference is a set of scripts, not a
package), so :func:`ensure_installed` falls back to a verbatim copy of the tree into
``baseline/_ref`` (git-ignored, shipped to the GPU box by gpurun).  Nothing under ``_ref`` is
edited.  What this harness supplies around the script is only what the no-network box lacks:

  * a local "model" directory: the named architecture's ``config.json`` + a throw-away word-level
    tokenizer (weights are random-init in the reference anyway: ``AutoModelForCausalLM.from_config``);
  * a local text dataset whose lines are exactly ``seq_length`` tokens, so the script's own
    tokenize/group pipeline yields exactly ``(warmup+steps) * world * batch`` samples and the
    script stops by itself after ``warmup+steps`` optimizer steps (``--num-epochs 1``);
  * a logging handler that timestamps (after ``cuda.synchronize``) each per-step record the script
    logs with ``--log-freq 1``; the timed region is the span between record W and record W+K,
    max over ranks.

None of this repository's models, kernels or engine are imported on this path.
"""
from __future__ import annotations

import json
import logging
import os
import random
import runpy
import shutil
import subprocess
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
REF_SRC = Path("/root/reference")

CHAPTER_SCRIPTS = {
    "ddp": "02-distributed-data-parallel/train_llm.py",
    "fsdp": "04-fully-sharded-data-parallel/train_llm.py",
    "tp": "06-tensor-parallel/train_llm.py",
    "2d": "07-2d-parallel/train_llm.py",
    "single": "01-single-gpu/train_llm.py",
}

# HF config.json payloads for the benchmark models (kept here so this file does not import the
# b200 package).  Llama-2-7B is BASELINE.json's headline config.
MODEL_CONFIGS = {
    "meta-llama/Llama-2-7b-hf": dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008,
                                     num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=32,
                                     max_position_embeddings=4096, rope_theta=10000.0),
    "meta-llama/Meta-Llama-3-8B": dict(vocab_size=128256, hidden_size=4096, intermediate_size=14336,
                                       num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8,
                                       max_position_embeddings=8192, rope_theta=500000.0),
    "meta-llama/Meta-Llama-3-70B": dict(vocab_size=128256, hidden_size=8192, intermediate_size=28672,
                                        num_hidden_layers=80, num_attention_heads=64, num_key_value_heads=8,
                                        max_position_embeddings=8192, rope_theta=500000.0),
    "debug-llama": dict(vocab_size=1024, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                        num_attention_heads=2, num_key_value_heads=2, max_position_embeddings=2048, rope_theta=10000.0),
}


def ensure_installed() -> str:
    """Returns a one-line description of how ``baseline/_ref`` was populated (raises if impossible)."""
    marker = REF_DIR / ".installed"
    if marker.exists():
        return marker.read_text().strip()
    if not REF_SRC.exists():
        raise RuntimeError("baseline/_ref is empty and /root/reference is not available to copy from")
    REF_DIR.mkdir(parents=True, exist_ok=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
           "/opt/wheelhouse", "--target", str(REF_DIR), str(REF_SRC)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode == 0:
        how = "pip install --target baseline/_ref succeeded"
    else:
        for item in REF_SRC.iterdir():
            dst = REF_DIR / item.name
            if item.is_dir():
                shutil.copytree(item, dst, dirs_exist_ok=True)
            else:
                shutil.copy2(item, dst)
        how = ("pip install refused (reference has no setup.py/pyproject.toml: it is a tree of scripts); "
               "copied the unmodified tree to baseline/_ref")
    marker.write_text(how + "\n")
    return how


# ---------------------------------------------------------------------------------------------
# offline assets
# ---------------------------------------------------------------------------------------------
def _write_model_dir(path: Path, model_name: str, num_layers=None):
    cfg = dict(MODEL_CONFIGS[model_name])
    if num_layers:
        cfg["num_hidden_layers"] = int(num_layers)
    cfg.update(model_type="llama", architectures=["LlamaForCausalLM"], rms_norm_eps=1e-5, hidden_act="silu",
               tie_word_embeddings=False, attention_bias=False, mlp_bias=False, bos_token_id=1, eos_token_id=2,
               torch_dtype="bfloat16")
    path.mkdir(parents=True, exist_ok=True)
    (path / "config.json").write_text(json.dumps(cfg, indent=1))
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast

    n_words = min(2000, cfg["vocab_size"] - 8)
    vocab = {"<unk>": 0, "<s>": 1, "</s>": 2}
    for i in range(n_words):
        vocab[f"w{i}"] = 3 + i
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="<unk>", bos_token="<s>", eos_token="</s>")
    fast.save_pretrained(str(path))
    return n_words


def _write_dataset(path: Path, n_lines: int, seq_length: int, n_words: int, seed: int = 0):
    path.mkdir(parents=True, exist_ok=True)
    rng = random.Random(seed)
    with open(path / "train.txt", "w") as fp:
        for _ in range(n_lines):
            fp.write(" ".join(f"w{rng.randrange(n_words)}" for _ in range(seq_length)) + "\n")


def prepare_assets(root: Path, model_name: str, seq_length: int, n_samples: int, rank: int, num_layers=None):
    ready = root / ".ready"
    if rank == 0:
        if root.exists():
            shutil.rmtree(root)
        n_words = _write_model_dir(root / "model", model_name, num_layers)
        _write_dataset(root / "data", n_samples, seq_length, n_words)
        ready.write_text("ok")
    else:
        t0 = time.time()
        while not ready.exists():
            if time.time() - t0 > 600:
                raise TimeoutError("rank 0 never finished preparing the offline assets")
            time.sleep(0.2)
    return str(root / "model"), str(root / "data")


def _environment_shims():
    """Library-version glue OUTSIDE the reference tree (nothing under ``baseline/_ref`` is touched): the reference
    was written against transformers 4.4x; this image ships 5.5, which dropped the per-instance
    ``LlamaRotaryEmbedding.rope_init_fn`` attribute that chapter 04/05/07's ``reset_rope`` calls after
    ``to_empty()``.  Restore it as a class attribute with the same contract ((config, device) -> (inv_freq, scaling))
    so the unmodified scripts run; returns the list of shims applied (reported in the JSON line)."""
    applied = []
    try:
        from transformers.models.llama.modeling_llama import LlamaRotaryEmbedding

        if not hasattr(LlamaRotaryEmbedding, "rope_init_fn") and hasattr(LlamaRotaryEmbedding,
                                                                         "compute_default_rope_parameters"):
            LlamaRotaryEmbedding.rope_init_fn = staticmethod(LlamaRotaryEmbedding.compute_default_rope_parameters)
            applied.append("transformers>=5: LlamaRotaryEmbedding.rope_init_fn -> compute_default_rope_parameters")
    except Exception as e:  # pragma: no cover - depends on the installed transformers
        applied.append(f"rope shim failed: {type(e).__name__}")
    return applied


# ---------------------------------------------------------------------------------------------
# run + time
# ---------------------------------------------------------------------------------------------
class _StepRecorder(logging.Handler):
    """Collects the reference's per-step info dicts with a device-synchronised wall timestamp."""

    def __init__(self):
        super().__init__(level=logging.INFO)
        self.records = []

    def emit(self, record):
        msg = record.msg
        if isinstance(msg, dict) and "global_step" in msg and "tokens_per_s" in msg:
            import torch

            if torch.cuda.is_available():
                torch.cuda.synchronize()
            self.records.append((time.perf_counter(), dict(msg)))


def run_reference(parallelism: str, model_name: str, gpus: int, steps: int, warmup: int, seq_length: int,
                  batch_size: int, num_layers=None, extra_args=()):
    """Runs the chapter script in this process (this process is one rank).  Returns a dict with
    ms_per_step (max over ranks), tokens_per_s (whole job) and the script's own timer breakdown."""
    how = ensure_installed()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_RANK", str(rank))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    tag = f"{os.getppid() if world > 1 else os.getpid()}"
    root = Path(os.environ.get("TMPDIR", "/tmp")) / f"dtg_ref_assets_{tag}"
    os.environ["HF_HOME"] = str(root.parent / f"dtg_ref_hf_{tag}")
    os.environ["HF_HUB_OFFLINE"] = "1"
    os.environ["HF_DATASETS_OFFLINE"] = "1"
    os.environ["TOKENIZERS_PARALLELISM"] = "false"
    os.environ.setdefault("OMP_NUM_THREADS", "8")

    dp = world if parallelism in ("ddp", "fsdp", "single") else 1
    if parallelism == "2d":
        tp = int(extra_args[extra_args.index("-tp") + 1]) if "-tp" in extra_args else 4
        dp = max(1, world // tp)
    n_samples = (steps + warmup) * dp * batch_size
    extra_args = list(extra_args)
    model_dir, data_dir = prepare_assets(root, model_name, seq_length, n_samples, rank, num_layers)

    script = REF_DIR / CHAPTER_SCRIPTS[parallelism]
    argv = [str(script), "-d", data_dir, "-m", model_dir, "-s", str(seq_length), "-b", str(batch_size),
            "--num-epochs", "1", "--log-freq", "1", "--ckpt-freq", "1000000", *extra_args]
    rec = _StepRecorder()
    root_logger = logging.getLogger()
    root_logger.addHandler(rec)
    root_logger.setLevel(logging.INFO)
    if not any(isinstance(h, logging.StreamHandler) for h in root_logger.handlers):
        root_logger.addHandler(logging.StreamHandler(sys.stderr))

    shims = _environment_shims()
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = argv
    os.chdir(script.parent)
    try:
        runpy.run_path(str(script), run_name="__main__")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
        root_logger.removeHandler(rec)

    if len(rec.records) < warmup + steps:
        raise RuntimeError(f"reference logged {len(rec.records)} steps, expected {warmup + steps}")
    # record i is emitted at the END of step i+1; steps W+1..W+K lie between records W-1 and W+K-1
    if warmup < 1:
        raise ValueError("need at least one warm-up step to bracket the timed region")
    t_start = rec.records[warmup - 1][0]
    t_end = rec.records[warmup + steps - 1][0]
    elapsed_ms = 1000.0 * (t_end - t_start)
    timed = [r for _, r in rec.records[warmup:warmup + steps]]
    timers_ms = sum(r["time/total"] for r in timed)

    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and world > 1:
        t = torch.tensor([elapsed_ms, timers_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, timers_ms = t.tolist()
    tokens = steps * dp * batch_size * seq_length
    out = {
        "ms_per_step": elapsed_ms / steps,
        "tokens_per_s": 1000.0 * tokens / elapsed_ms,
        "ref_timers_ms_per_step": timers_ms / steps,
        "ref_breakdown_ms": {k: sum(r[k] for r in timed) / steps
                             for k in ("time/data", "time/forward", "time/backward", "time/update")},
        "peak_alloc_gb": max(r.get("peak_alloc_gb", 0.0) for r in timed),
        "last_loss": timed[-1].get("running_loss"),
        "dp_size": dp,
        "install": how,
        "environment_shims": shims,
        "script": CHAPTER_SCRIPTS[parallelism],
    }
    if dist.is_available() and dist.is_initialized():
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass
    shutil.rmtree(root, ignore_errors=True) if rank == 0 else None
    return out
