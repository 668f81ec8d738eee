"""Run the attention fwd/bwd kernels on the 7B shape (target for ncu) and time them."""
import sys, os, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_training_guide_b200 import _ext
from distributed_training_guide_b200.utils.timers import device_time_ms

C = _ext.load(True)
B, S, nh, nkv = 1, 4096, 32, 32
qkv = torch.randn(B, S, nh + 2 * nkv, 128, device="cuda", dtype=torch.bfloat16)
do = torch.randn(B, S, nh, 128, device="cuda", dtype=torch.bfloat16)
sc = 1 / math.sqrt(128)
o, lse = C.attn_fwd(qkv, nh, nkv, sc)
for ver in (1, 2):
    ms_v, _ = device_time_ms(lambda: C.attn_fwd(qkv, nh, nkv, sc, ver), warmup=2, iters=5)
    print(f"attn fwd v{ver} {ms_v:.3f} ms = {4 * B * nh * S * S * 128 / 2 / ms_v / 1e9:.1f} TFLOP/s")
ms_f, _ = device_time_ms(lambda: C.attn_fwd(qkv, nh, nkv, sc), warmup=2, iters=5)
for mode in (1, 2):
    ms_m, _ = device_time_ms(lambda: C.attn_bwd(do, qkv, o, lse, nh, nkv, sc, None, mode), warmup=2, iters=5)
    print(f"attn bwd mode {mode} ({'P/dS via smem' if mode == 1 else 'P/dS in TMEM'}) {ms_m:.3f} ms = "
          f"{2.5 * 4 * B * nh * S * S * 128 / 2 / ms_m / 1e9:.1f} TFLOP/s (5-GEMM count)")
ms_b, _ = device_time_ms(lambda: C.attn_bwd(do, qkv, o, lse, nh, nkv, sc), warmup=2, iters=5)
fl = 4 * B * nh * S * S * 128 / 2
print(f"attn fwd {ms_f:.3f} ms = {fl / ms_f / 1e9:.1f} TFLOP/s ; bwd {ms_b:.3f} ms = {2.5 * fl / ms_b / 1e9:.1f} TFLOP/s (5-GEMM count)")
q, k, v = qkv[:, :, :nh].transpose(1, 2), qkv[:, :, nh:nh + nkv].transpose(1, 2), qkv[:, :, nh + nkv:].transpose(1, 2)
ms_s, _ = device_time_ms(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True), warmup=2, iters=5)
print(f"torch sdpa fwd {ms_s:.3f} ms = {fl / ms_s / 1e9:.1f} TFLOP/s")
try:
    from flash_attn import flash_attn_func
    qq, kk, vv = qkv[:, :, :nh].contiguous(), qkv[:, :, nh:nh + nkv].contiguous(), qkv[:, :, nh + nkv:].contiguous()
    ms_fa, _ = device_time_ms(lambda: flash_attn_func(qq, kk, vv, causal=True), warmup=2, iters=5)
    print(f"flash_attn2 fwd {ms_fa:.3f} ms = {fl / ms_fa / 1e9:.1f} TFLOP/s")
except Exception as e:
    print("flash_attn unavailable", e)
