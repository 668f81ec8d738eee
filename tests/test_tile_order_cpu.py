"""The GEMM's persistent tile order is a bijection onto the output tiles in every mode (host build of the same header)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host compilation only)")
def test_tile_order_visits_every_tile_once(tmp_path):
    exe = tmp_path / "tile_order_test"
    src = ROOT / "tests" / "native" / "tile_order_test.cu"
    inc = ROOT / "distributed_training_guide_b200" / "csrc"
    r = subprocess.run(["nvcc", "-std=c++17", "-O1", f"-I{inc}", "-o", str(exe), str(src)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout[-2000:]
