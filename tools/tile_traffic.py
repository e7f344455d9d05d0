import os, sys
sys.path.insert(0, "/root/repo")
import torch
sys.path.insert(0, "/root/repo/tools")
from sweep import timeit, rnd, dev, bf
from avt_b200 import ops
M = 15760
for N, K in [(3072, 768), (768, 3072)]:
    a = rnd(M, K); w = rnd(N, K, scale=0.03); out = torch.empty(M, N, device=dev, dtype=bf)
    for bn, cg in [(256, 2), (128, 2), (256, 1), (128, 1)]:
        t = timeit(lambda: ops.gemm(a, w, out, block_n=bn, cta_group=cg))
        tiles = -(-M // (128 * cg)) * (N // bn)
        kb = K // 64
        per_kb = (128 * cg + bn) * 64 * 2   # operand bytes per group per k-block
        traffic = tiles * kb * per_kb
        print(f"N{N} K{K} bn{bn} cg{cg}: {t:6.1f} us {2.0*M*N*K/t/1e6:7.1f} TF/s  operand traffic {traffic/1e6:6.0f} MB -> {traffic/t/1e6:5.2f} TB/s", flush=True)
