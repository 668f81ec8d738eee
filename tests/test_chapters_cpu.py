"""Every distributed chapter script end to end under torchrun on the CPU (gloo, 2 processes, debug model): launch,
train, checkpoint, relaunch, resume.  The GPU twin is tests/test_gpu_chapters.py."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _torchrun(script, args, nproc=2, timeout=420):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
           "--nproc-per-node", str(nproc), str(script)] + args
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(Path(script).parent), timeout=timeout)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    recs = [eval(l.split("INFO:", 1)[1]) for l in r.stderr.splitlines() if "INFO:{" in l and "rank=0" in l]
    return recs, r.stderr


@pytest.mark.parametrize("chapter,model,extra", [
    ("02-distributed-data-parallel", "debug-llama", []),
    ("02-distributed-data-parallel", "debug-gpt2", []),
    ("04-fully-sharded-data-parallel", "debug-llama", ["--cpu-offload"]),
    ("05-training-llama-405b", "debug-llama", ["--checkpoint-activations", "--prefetch-layers"]),
    ("06-tensor-parallel", "debug-llama-tp", []),
    ("07-2d-parallel", "debug-llama-tp", ["-tp", "2"]),
    ("alternative-frameworks/deepspeed", "debug-llama", ["--deepspeed_config", "ds_config.json", "--wandb", "off"]),
])
def test_chapter_trains_checkpoints_and_resumes(tmp_path, chapter, model, extra):
    script = ROOT / chapter / "train_llm.py"
    common = ["-d", "synthetic", "-m", model, "-s", "32", "-b", "2", "--num-samples", "32", "--log-freq", "1",
              "--save-dir", str(tmp_path), "-e", "exp", "--ckpt-freq", "2", "--lr", "1e-3", "--device", "cpu"] + extra
    recs, _ = _torchrun(script, common + ["--max-steps", "2"])
    assert [r["global_step"] for r in recs] == [1, 2] and all(0 < r["running_loss"] < 20 for r in recs)
    assert recs[-1]["tokens_per_s"] > 0
    assert json.loads((tmp_path / "exp" / "state.json").read_text())["global_step"] == 2
    recs2, log = _torchrun(script, common + ["--max-steps", "4"])
    assert "Resumed=True" in log and [r["global_step"] for r in recs2] == [3, 4]
