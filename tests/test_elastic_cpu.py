"""The fault-injection toy under torchrun --max-restarts: failures are injected, the gang restarts, the
state file carries the step count across restarts."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_toy_restarts_and_finishes(tmp_path):
    state = tmp_path / "toy-state.json"
    env = dict(os.environ, TOY_STATE_FILE=str(state), TORCHELASTIC_ERROR_FILE=str(tmp_path / "error.json"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
           "--nproc-per-node", "2", "--max-restarts", "20",
           str(ROOT / "related-topics" / "elastic-training" / "toy.py"), "--steps", "40", "--failure-prob", "0.03",
           "--step-time", "0.001"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert json.loads(state.read_text())["num_steps"] == 40
    assert "finished 40 steps" in out
    if "injected failure" in out:  # with p=0.03 x 2 ranks x 40 steps a failure is near-certain
        assert "restart count=1" in out or "restart count=2" in out
