"""CUDA-event timing of the tcgen05 attention kernels at the cfg2 shape (80 frames x 12 heads x 197 tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avt_b200 import ops
F, H, N = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (80, 12, 197)
D = H * 64
qkv = torch.randn(F * N, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.empty(F * N, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(F * H, N, device="cuda")
dout = torch.randn(F * N, D, device="cuda").to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, n=20, cold=True):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


fw = t(lambda: ops.attention_tc_fwd(qkv, out, lse, F, H, N, scale=0.125))
bw = t(lambda: ops.attention_tc_bwd(qkv, out, dout, lse, dqkv, F, H, N, scale=0.125))
fww = t(lambda: ops.attention_tc_fwd(qkv, out, lse, F, H, N, scale=0.125), cold=False)
bww = t(lambda: ops.attention_tc_bwd(qkv, out, dout, lse, dqkv, F, H, N, scale=0.125), cold=False)
print(f"L2-warm (inputs left in L2 by the previous launch, as after the qkv GEMM): fwd {fww:.1f} us  bwd {bww:.1f} us")
gf = 4.0 * N * N * 64 * F * H / 1e9
print(f"attention F={F} H={H} N={N}: fwd {fw:.1f} us ({gf / fw:.0f} TF/s)  bwd {bw:.1f} us ({2 * gf / bw:.0f} TF/s algorithmic 2x)")
