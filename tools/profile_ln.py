"""LayerNorm fwd/bwd + colsum at the cfg2 shape for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avt_b200 import ops
rows, D = 15760, 768
x = torch.randn(rows, D, device="cuda")
add = torch.randn(rows, D, device="cuda").to(torch.bfloat16)
g, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
y = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
xo = torch.empty(rows, D, device="cuda")
mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
dy = torch.randn(rows, D, device="cuda").to(torch.bfloat16)
dx = torch.randn(rows, D, device="cuda")
dxb = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
dg, db = torch.empty(D, device="cuda"), torch.empty(D, device="cuda")
ws = torch.empty(ops.layernorm_bwd_workspace(rows, D), dtype=torch.uint8, device="cuda")
big = torch.randn(rows, 3072, device="cuda").to(torch.bfloat16)
acc = torch.zeros(3072, device="cuda")
for _ in range(3):
    ops.layernorm_fwd(x, g, b, 1e-6, y, mean, rstd, add=add, x_out=xo)
    ops.layernorm_bwd(dy, x, mean, rstd, g, dx, dg, db, ws, dx_in=dx, dx_bf16=dxb)
    ops.colsum(big, acc)
torch.cuda.synchronize()
print("ok")
