"""Run one tcgen05 GEMM shape a few times (target for `ncu -k regex:gemm_bf16`)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_training_guide_b200 import _ext

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M, N, K = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (4096, 4096, 4096)
C = _ext.load(True)
print("max active clusters cg1/cg2:", C.gemm_max_active_clusters(1), C.gemm_max_active_clusters(2))
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    C.gemm(a, b, out, False, True, False, variant)
torch.cuda.synchronize()
print("ok")
