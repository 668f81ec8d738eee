"""Sustained (power-capped) throughput of the tcgen05 GEMM vs cuBLAS: each shape is run back to back for
``--seconds`` so the chip settles at its 1 kW operating point (short bursts run at 1965 MHz and hide
differences in energy per FLOP).  Samples SM clock and power while the loop runs.
Usage: python tools/gemm_sustained.py [--seconds 2.0] [--shapes square_8192 7b_qkv_fwd ...]"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_training_guide_b200 import _ext  # noqa: E402
from gemm_bench import SHAPES  # noqa: E402


class Sampler:
    def __init__(self):
        import pynvml

        pynvml.nvmlInit()
        self.nv = pynvml
        self.h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        self.clk, self.pw, self._stop = [], [], False

    def __enter__(self):
        self.clk, self.pw, self._stop = [], [], False
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def _run(self):
        while not self._stop:
            self.clk.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.pw.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            time.sleep(0.05)

    def __exit__(self, *a):
        self._stop = True
        self.t.join()

    def summary(self):
        half = len(self.clk) // 2  # second half of the window: settled
        return {"sm_mhz": statistics.median(self.clk[half:] or [0]), "power_w": statistics.median(self.pw[half:] or [0])}


def sustained(fn, seconds):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # calibrate the iteration count from a short burst, then time the second half of a long run
    s.record()
    for _ in range(10):
        fn()
    e.record()
    torch.cuda.synchronize()
    per = s.elapsed_time(e) / 10
    n = max(20, int(seconds * 1000 / per))
    with Sampler() as smp:
        for _ in range(n // 2):
            fn()
        s.record()
        for _ in range(n // 2):
            fn()
        e.record()
        torch.cuda.synchronize()
    return s.elapsed_time(e) / (n // 2), smp.summary()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--shapes", nargs="+", default=["square_8192", "7b_qkv_fwd", "7b_gateup_fwd", "7b_down_fwd",
                                                    "7b_gateup_dgrad", "7b_gateup_wgrad"])
    ap.add_argument("--out", default="gpurun_out/gemm_sustained.json")
    args = ap.parse_args()
    C = _ext.load(True)
    res = {}
    for name in args.shapes:
        M, N, K, ta, tb = SHAPES[name]
        a = torch.randn((K, M) if ta else (M, K), device="cuda", dtype=torch.bfloat16)
        b = torch.randn((N, K) if tb else (K, N), device="cuda", dtype=torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        A, B = (a.t() if ta else a), (b.t() if tb else b)
        flops = 2.0 * M * N * K
        row = {}
        for label, fn in (("cublas", lambda: torch.matmul(A, B, out=out)),
                          ("tcgen05", lambda: C.gemm(a, b, out, ta, tb, False, 3))):
            ms, clk = sustained(fn, args.seconds)
            row[label] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), **clk,
                          "flop_per_clk_per_sm": round(flops / (ms * 1e-3) / (clk["sm_mhz"] * 1e6) / 148, 0) if clk["sm_mhz"] else None}
            time.sleep(1.0)  # let the chip cool between arms
        res[name] = row
        print(name, json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fp:
        json.dump(res, fp, indent=1)


if __name__ == "__main__":
    main()
