"""Chapter scripts end to end on a GPU: chapter 01 with the Llama kernels and with GPT-2 (PyTorch-op path),
checkpoint + resume; chapters 02/04/06/07 under torchrun when >= 2 GPUs are present."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run(script, args, nproc=0, timeout=600):
    if nproc:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
               "--nproc-per-node", str(nproc), str(script)] + args
    else:
        cmd = [sys.executable, str(script)] + args
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(Path(script).parent), timeout=timeout)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    recs = [eval(l.split("INFO:", 1)[1]) for l in r.stderr.splitlines() if "INFO:{" in l]
    return recs, r.stderr


@pytest.mark.parametrize("model,seq", [("debug-llama-gqa", "256"), ("debug-gpt2", "64")])
def test_chapter01_on_gpu(tmp_path, model, seq):
    script = ROOT / "01-single-gpu" / "train_llm.py"
    common = ["-d", "synthetic", "-m", model, "-s", seq, "-b", "2", "--num-samples", "32", "--log-freq", "1",
              "--save-dir", str(tmp_path), "-e", "exp", "--ckpt-freq", "3", "--lr", "1e-3"]
    recs, _ = _run(script, common + ["--max-steps", "3"])
    assert len(recs) == 3 and all(r["tokens_per_s"] > 0 for r in recs)
    assert recs[0]["peak_alloc_gb"] > 0
    assert json.loads((tmp_path / "exp" / "state.json").read_text())["global_step"] == 3
    recs2, log = _run(script, common + ["--max-steps", "6"])
    assert "Resumed=True" in log and recs2[-1]["global_step"] == 6
    assert all(0 < r["running_loss"] < 20 for r in recs + recs2)  # random tokens: finite, near ln(V)


@pytest.mark.multigpu
@pytest.mark.parametrize("chapter,extra", [
    ("02-distributed-data-parallel", []),
    ("04-fully-sharded-data-parallel", []),
    ("04-fully-sharded-data-parallel", ["--cpu-offload"]),
    ("05-training-llama-405b", ["--checkpoint-activations", "--prefetch-layers"]),
    ("06-tensor-parallel", []),
    ("07-2d-parallel", ["-tp", "2"]),
])
def test_distributed_chapters_on_gpus(tmp_path, chapter, extra):
    script = ROOT / chapter / "train_llm.py"
    model = "debug-llama-tp" if chapter.startswith(("06", "07")) else "debug-llama-gqa"
    args = ["-d", "synthetic", "-m", model, "-s", "256", "-b", "2", "--num-samples", "32", "--log-freq", "1",
            "--save-dir", str(tmp_path), "-e", "exp", "--ckpt-freq", "2", "--lr", "1e-3", "--max-steps", "4"] + extra
    recs, _ = _run(script, args, nproc=2)
    losses = [r["running_loss"] for r in recs if r["global_step"] in (1, 4)]
    assert len(recs) >= 4 and losses[-1] < losses[0] + 0.5
    assert (tmp_path / "exp" / "state.json").exists()


@pytest.mark.multigpu
def test_2d_dp2_tp2_on_four_gpus(tmp_path):
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    script = ROOT / "07-2d-parallel" / "train_llm.py"
    args = ["-d", "synthetic", "-m", "debug-llama-tp", "-s", "256", "-b", "2", "--num-samples", "64", "--log-freq", "1",
            "--save-dir", str(tmp_path), "-e", "exp", "--ckpt-freq", "2", "--lr", "1e-3", "--max-steps", "4", "-tp", "2"]
    recs, _ = _run(script, args, nproc=4)
    assert len(recs) >= 4 and all(0 < r["running_loss"] < 20 for r in recs)
    # resume from the sharded checkpoint
    recs2, log = _run(script, args[:-4] + ["--max-steps", "6", "-tp", "2"], nproc=4)
    assert "Resumed=True" in log and recs2[-1]["global_step"] == 6
