"""One launch of the attention forward (and optionally backward) at the cfg2 shape, for the AVT_ATTN_TRACE build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avt_b200 import ops
F, H, N = 80, 12, 197
D = H * 64
qkv = torch.randn(F * N, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.empty(F * N, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(F * H, N, device="cuda")
ops.attention_tc_fwd(qkv, out, lse, F, H, N, scale=0.125)
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "bwd":
    dout = torch.randn(F * N, D, device="cuda").to(torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    ops.attention_tc_bwd(qkv, out, dout, lse, dqkv, F, H, N, scale=0.125)
    torch.cuda.synchronize()
