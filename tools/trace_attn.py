"""In-kernel timeline of the attention backward pipeline (clock64 at the hand-over points of CTA (0,0))."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_training_guide_b200 import _ext
C = _ext.load(True)
B, S, nh, nkv = 1, 4096, 32, 32
qkv = torch.randn(B, S, nh + 2 * nkv, 128, device="cuda", dtype=torch.bfloat16)
do = torch.randn(B, S, nh, 128, device="cuda", dtype=torch.bfloat16)
sc = 1 / math.sqrt(128)
o, lse = C.attn_fwd(qkv, nh, nkv, sc)
for _ in range(2):
    C.attn_bwd(do, qkv, o, lse, nh, nkv, sc)
trace = torch.zeros(1024, dtype=torch.int64, device="cuda")
C.attn_bwd(do, qkv, o, lse, nh, nkv, sc, trace)
torch.cuda.synchronize()
t = trace.cpu().view(2, 64, 8)
names = ["sm:sdp_full", "sm:computed", "sm:pds_empty", "sm:arrived", "mma:scores_issued", "mma:pds_full", "mma:grads_committed"]
for k, label in enumerate(("KV pass", "Q pass")):
    x = t[k]
    base = int(x[2, 0])
    print(f"== {label}: per-iteration timeline of CTA(0,0), cycles relative to iteration 2's sdp_full")
    for it in range(2, 14):
        row = [int(x[it, c]) - base for c in range(7)]
        print(f"it {it:2d}: " + "  ".join(f"{n}={v:6d}" for n, v in zip(names, row)))
    per = (int(x[40, 3]) - int(x[8, 3])) / 32.0
    print(f"steady-state period: {per:.0f} cycles / iteration")
