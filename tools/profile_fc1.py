"""fc1-shaped GEMM with increasingly heavy epilogues (for ncu --set full): plain, bias+GELU, bias+GELU+gelu' aux, dgrad*gelu'."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avt_b200 import ops
M = 15760
dev = "cuda"
bf = torch.bfloat16
a = torch.randn(M, 768, device=dev).to(bf)
w1 = (torch.randn(3072, 768, device=dev) * 0.03).to(bf)
b1 = torch.randn(3072, device=dev)
h = torch.empty(M, 3072, device=dev, dtype=bf)
z = torch.empty(M, 3072, device=dev, dtype=bf)
g = torch.randn(M, 768, device=dev).to(bf)
w2 = (torch.randn(768, 3072, device=dev) * 0.03).to(bf)
dz = torch.empty(M, 3072, device=dev, dtype=bf)
for _ in range(3):
    ops.gemm(a, w1, h)
    ops.gemm(a, w1, h, bias=b1, act=1)
    ops.gemm(a, w1, h, bias=b1, act=1, aux_z=z, aux_grad=True)
    ops.gemm(g, w2, dz, b_mn=True, dact_z=z, dact=1, dact_is_grad=True)
torch.cuda.synchronize()
print("ok")
