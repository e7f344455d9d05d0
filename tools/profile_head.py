"""AVT-h GEMMs at the expts/01 shape (M = 80 rows) for ncu: c_fc forward (split-K 2 + finishing pass), its dgrad and its
weight gradient (contraction over 80 rows: the kernel is its fp32 TMA-store epilogue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avt_b200 import ops
from avt_b200.engine import small_m_split
dev, bf = "cuda", torch.bfloat16
M, K, N = 80, 2048, 8192
a = torch.randn(M, K, device=dev).to(bf)
w = (torch.randn(K, N, device=dev) * 0.02).to(bf)          # Conv1D layout [in, out]
bias = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=bf)
z = torch.empty(M, N, device=dev, dtype=bf)
dy = torch.randn(M, N, device=dev).to(bf)
dx = torch.empty(M, K, device=dev, dtype=bf)
dW = torch.empty(K, N, device=dev)
sk = small_m_split(M, N, K)
ws = torch.empty(max(sk, small_m_split(M, K, N)) * M * max(N, K), device=dev)
for _ in range(3):
    ops.gemm(a, w, out, b_mn=True, bias=bias, act=2, aux_z=z, aux_grad=True, split_k=sk, workspace=ws)
    ops.gemm(dy, w, dx, b_mn=False, split_k=small_m_split(M, K, N), workspace=ws)
    ops.gemm(a, dy, dW, a_mn=True, b_mn=True)
torch.cuda.synchronize()
print("ok", sk, small_m_split(M, K, N))
