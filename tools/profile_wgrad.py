"""ViT weight-gradient GEMMs (stream-K, fp32 atomics epilogue) next to the forward GEMM of the same flops, for ncu --set full."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avt_b200 import ops
from avt_b200.engine import _split_k_for
M, D = 15760, 768
dev, bf = "cuda", torch.bfloat16
h = torch.randn(M, 4 * D, device=dev).to(bf)
dy = torch.randn(M, D, device=dev).to(bf)
w2 = (torch.randn(D, 4 * D, device=dev) * 0.03).to(bf)
y = torch.empty(M, D, device=dev, dtype=bf)
g2 = torch.zeros(D, 4 * D, device=dev)
att = torch.randn(M, D, device=dev).to(bf)
gp = torch.zeros(D, D, device=dev)
sk2 = _split_k_for(D, 4 * D, M, 256)
skp = _split_k_for(D, D, M, 256)
for _ in range(3):
    ops.gemm(h, w2, y)                                                        # fwd fc2 (plain store epilogue)
    ops.gemm(dy, h, g2, a_mn=True, b_mn=True, split_k=sk2, accumulate=True)   # wgrad fc2
    ops.gemm(dy, att, gp, a_mn=True, b_mn=True, split_k=skp, accumulate=True)  # wgrad proj
torch.cuda.synchronize()
print("ok", sk2, skp)
