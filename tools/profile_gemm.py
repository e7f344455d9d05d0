"""A few GEMM launches with the ViT-B cfg2 shapes (for ncu --set full): plain bf16 store, GELU + aux store, residual."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avt_b200 import ops
M = 15760
dev = "cuda"
a = torch.randn(M, 768, device=dev).to(torch.bfloat16)
w1 = torch.randn(3072, 768, device=dev).to(torch.bfloat16) * 0.03
wq = torch.randn(2304, 768, device=dev).to(torch.bfloat16) * 0.03
wp = torch.randn(768, 768, device=dev).to(torch.bfloat16) * 0.03
b1 = torch.randn(3072, device=dev)
h = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
z = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
qkv = torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
res = torch.randn(M, 768, device=dev)
xo = torch.empty(M, 768, device=dev)
dz = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
g = torch.randn(M, 768, device=dev).to(torch.bfloat16)
w2 = torch.randn(768, 3072, device=dev).to(torch.bfloat16) * 0.03
for _ in range(2):
    ops.gemm(a, wq, qkv)                                        # plain
    ops.gemm(a, w1, h, bias=b1, act=1, aux_z=z, aux_grad=True)  # GELU + gelu' aux
    ops.gemm(a, wp, xo, residual=res)                           # residual fp32
    ops.gemm(g, w2, dz, b_mn=True, dact_z=z, dact=1, dact_is_grad=True)  # dgrad * gelu'
    ops.gemm(h, g, torch.zeros(3072, 768, device=dev), a_mn=True, b_mn=True, split_k=4)  # wgrad
torch.cuda.synchronize()
print("ok")
