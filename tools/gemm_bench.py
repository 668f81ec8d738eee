"""Device-timed throughput of the tcgen05 GEMM vs torch.matmul (cuBLAS) on the training shapes.
Writes gpurun_out/gemm_bench.json.  Usage: python tools/gemm_bench.py [--variants 1 2]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_training_guide_b200 import _ext  # noqa: E402
from distributed_training_guide_b200.utils.timers import device_time_ms  # noqa: E402

SHAPES = {  # name: (M, N, K, trans_a, trans_b)
    "7b_qkv_fwd": (4096, 12288, 4096, False, True),
    "7b_o_fwd": (4096, 4096, 4096, False, True),
    "7b_gateup_fwd": (4096, 22016, 4096, False, True),
    "7b_down_fwd": (4096, 4096, 11008, False, True),
    "7b_lmhead_fwd": (4096, 32000, 4096, False, True),
    "7b_gateup_dgrad": (4096, 4096, 22016, False, False),
    "7b_down_dgrad": (4096, 11008, 4096, False, False),
    "7b_gateup_wgrad": (22016, 4096, 4096, True, False),
    "7b_down_wgrad": (4096, 11008, 4096, True, False),
    "7b_qkv_wgrad": (12288, 4096, 4096, True, False),
    "7b_lmhead_wgrad": (32000, 4096, 4096, True, False),
    "square_8192": (8192, 8192, 8192, False, True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, nargs="+", default=[1, 2])
    ap.add_argument("--out", default="gpurun_out/gemm_bench.json")
    args = ap.parse_args()
    C = _ext.load(True)
    res = {}
    for name, (M, N, K, ta, tb) in SHAPES.items():
        a = torch.randn((K, M) if ta else (M, K), device="cuda", dtype=torch.bfloat16)
        b = torch.randn((N, K) if tb else (K, N), device="cuda", dtype=torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        flops = 2.0 * M * N * K
        row = {"M": M, "N": N, "K": K, "trans_a": ta, "trans_b": tb}
        A = a.t() if ta else a
        B = b.t() if tb else b
        ms, _ = device_time_ms(lambda: torch.matmul(A, B, out=out), warmup=3, iters=10)
        row["cublas_ms"] = ms
        row["cublas_tflops"] = flops / ms / 1e9
        for v in args.variants:
            try:
                ms, _ = device_time_ms(lambda: C.gemm(a, b, out, ta, tb, False, v), warmup=3, iters=10)
                row[f"v{v}_ms"] = ms
                row[f"v{v}_tflops"] = flops / ms / 1e9
            except Exception as e:  # noqa: BLE001
                row[f"v{v}_error"] = repr(e)
        res[name] = row
        print(name, {k: (round(x, 3) if isinstance(x, float) else x) for k, x in row.items()}, flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fp:
        json.dump(res, fp, indent=1)


if __name__ == "__main__":
    main()
