"""Micro-benchmarks with CUDA events (warm, back-to-back launches; rotating buffers defeat L2 residency for the big ones).

    python tools/sweep.py [wgrad] [head] [ln] [attn] [fc1]

Prints one line per configuration: used to set the split-K heuristics and to check kernel changes before a full bench.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avt_b200 import ops

dev = "cuda"
bf = torch.bfloat16


def timeit(fn, iters=20, warm=3):
    """Device time per call: `iters` calls captured in one CUDA graph (no host launch cost between them)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    except Exception:
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def rnd(*shape, dt=bf, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dt)


def wgrad():
    M = 15760
    for name, mw, nw in [("fc1/fc2 wgrad", 3072, 768), ("fc2 wgrad (768x3072)", 768, 3072), ("qkv wgrad", 2304, 768),
                         ("proj wgrad", 768, 768)]:
        x, dy = rnd(M, nw), rnd(M, mw)
        dW = torch.zeros(mw, nw, device=dev)
        for sk in (1, 2, 3, 4, 5, 6, 8, 12, 16):
            t = timeit(lambda: ops.gemm(dy, x, dW, a_mn=True, b_mn=True, split_k=sk, accumulate=sk > 1))
            print(f"{name:24s} sk{sk:<3d} {t:8.1f} us  {2.0 * M * mw * nw / t / 1e6:8.1f} TF/s", flush=True)


def head():
    M = 80
    shapes = [("c_attn", 6144, 2048), ("attn.c_proj", 2048, 2048), ("c_fc", 8192, 2048), ("mlp.c_proj", 2048, 8192),
              ("encoder", 2048, 768), ("decoder", 768, 2048)]
    for name, N, K in shapes:
        a = rnd(M, K)
        w = rnd(K, N, scale=0.02)          # Conv1D layout [in, out]
        out = torch.empty(M, N, device=dev, dtype=bf)
        bias = torch.randn(N, device=dev)
        for bn in (64, 128):
            for sk in (1, 2, 3, 4, 6, 8, 11, 16):
                ws = torch.empty(sk * M * N, device=dev) if sk > 1 else None
                try:
                    t = timeit(lambda: ops.gemm(a, w, out, b_mn=True, bias=bias, split_k=sk, block_n=bn, workspace=ws))
                except RuntimeError as e:
                    print(name, bn, sk, "ERR", str(e)[:80])
                    continue
                print(f"{name:12s} fwd N{N} K{K} bn{bn:<3d} sk{sk:<2d} {t:7.1f} us  {K * N * 2 / t / 1e3:7.1f} GB/s(weights)", flush=True)
        # wgrad: dW[K, N] = X^T dY, contraction over 80 rows, fp32 output
        dy = rnd(M, N)
        dW = torch.empty(K, N, device=dev)
        for bn, cg in ((256, 2), (128, 2), (256, 1), (128, 1)):
            t = timeit(lambda: ops.gemm(a, dy, dW, a_mn=True, b_mn=True, block_n=bn, cta_group=cg))
            print(f"{name:12s} wgrad [{K}x{N}] bn{bn} cg{cg} {t:7.1f} us  {K * N * 4 / t / 1e3:7.1f} GB/s(dW write)", flush=True)


def ln():
    for rows, D in [(15760, 768), (80, 2048), (80, 768), (23640, 1024)]:
        x = torch.randn(rows, D, device=dev)
        add = rnd(rows, D)
        g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
        y = torch.empty(rows, D, device=dev, dtype=bf)
        xo = torch.empty(rows, D, device=dev)
        mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
        dy = rnd(rows, D)
        dx = torch.randn(rows, D, device=dev)
        dxb = torch.empty(rows, D, device=dev, dtype=bf)
        dg, db, dc = (torch.empty(D, device=dev) for _ in range(3))
        ws = torch.empty(ops.layernorm_bwd_workspace(rows, D), dtype=torch.uint8, device=dev)
        ops.layernorm_fwd(x, g, b, 1e-6, y, mean, rstd)
        tf0 = timeit(lambda: ops.layernorm_fwd(x, g, b, 1e-6, y, mean, rstd))
        print(f"ln rows {rows} D {D}: fwd without residual add {tf0:6.1f} us {rows * D * 6 / tf0 / 1e3:7.1f} GB/s", flush=True)
        tf = timeit(lambda: ops.layernorm_fwd(x, g, b, 1e-6, y, mean, rstd, add=add, x_out=xo))
        tb = timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dx, dg, db, ws, dx_in=dx, dx_bf16=dxb, dx_colsum=dc))
        tba = timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dx, dg, db, ws, dx_in=dx, dx_bf16=dxb, dx_colsum=dc, accumulate=True))
        print(f"ln rows {rows} D {D}: bwd, accumulate mode (column sums by atomics where the pipelined kernel runs) {tba:6.1f} us", flush=True)
        fb, bb = rows * D * (4 + 2 + 4 + 2), rows * D * (2 + 4 + 4 + 4 + 2)
        print(f"ln rows {rows} D {D}: fwd {tf:6.1f} us {fb / tf / 1e3:7.1f} GB/s | bwd(+reduce) {tb:6.1f} us {bb / tb / 1e3:7.1f} GB/s",
              flush=True)
    big = rnd(15760, 3072)
    acc = torch.zeros(3072, device=dev)
    t = timeit(lambda: ops.colsum(big, acc))
    print(f"colsum 15760x3072: {t:6.1f} us {big.numel() * 2 / t / 1e3:7.1f} GB/s")


def attn():
    B, H, N, hd = 8, 4, 10, 512
    D = H * hd
    qkv = rnd(B * N, 3 * D, scale=0.5)
    out = torch.empty(B * N, D, device=dev, dtype=bf)
    lse = torch.empty(B * H, N, device=dev)
    dout = rnd(B * N, D)
    dqkv = torch.empty_like(qkv)
    for p in (0.0, 0.1):
        tf = timeit(lambda: ops.attention_simt_fwd(qkv, out, lse, B, H, N, hd, causal=True, scale=hd ** -0.5, drop_p=p, seed=1))
        tb = timeit(lambda: ops.attention_simt_bwd(qkv, out, dout, lse, dqkv, B, H, N, hd, causal=True, scale=hd ** -0.5,
                                                   drop_p=p, seed=1))
        print(f"avt-h attention (B8 H4 N10 hd512) p={p}: fwd {tf:6.1f} us  bwd {tb:6.1f} us", flush=True)
    F, H, N = 80, 12, 197
    D = H * 64
    qkv = rnd(F * N, 3 * D)
    out = torch.empty(F * N, D, device=dev, dtype=bf)
    lse = torch.empty(F * H, N, device=dev)
    dout = rnd(F * N, D)
    dqkv = torch.empty_like(qkv)
    tf = timeit(lambda: ops.attention_tc_fwd(qkv, out, lse, F, H, N, scale=0.125))
    tb = timeit(lambda: ops.attention_tc_bwd(qkv, out, dout, lse, dqkv, F, H, N, scale=0.125))
    fl = 4.0 * N * N * 64 * F * H
    print(f"vit attention (80 frames x 12 heads): fwd {tf:6.1f} us {fl / tf / 1e6:6.1f} TF/s | bwd {tb:6.1f} us {2.5 * fl / tb / 1e6:6.1f} TF/s",
          flush=True)


def fc1():
    M = 15760
    a = rnd(M, 768)
    w1 = rnd(3072, 768, scale=0.03)
    b1 = torch.randn(3072, device=dev)
    h = torch.empty(M, 3072, device=dev, dtype=bf)
    z = torch.empty(M, 3072, device=dev, dtype=bf)
    g = rnd(M, 768)
    w2 = rnd(768, 3072, scale=0.03)
    dz = torch.empty(M, 3072, device=dev, dtype=bf)
    fl = 2.0 * M * 768 * 3072
    for name, fn in [("fc1 plain", lambda: ops.gemm(a, w1, h)),
                     ("fc1 bias", lambda: ops.gemm(a, w1, h, bias=b1)),
                     ("fc1 bias+gelu", lambda: ops.gemm(a, w1, h, bias=b1, act=1)),
                     ("fc1 bias+gelu+gelu' aux", lambda: ops.gemm(a, w1, h, bias=b1, act=1, aux_z=z, aux_grad=True)),
                     ("fc2 dgrad * gelu'", lambda: ops.gemm(g, w2, dz, b_mn=True, dact_z=z, dact=1, dact_is_grad=True)),
                     ("fc2 dgrad plain", lambda: ops.gemm(g, w2, dz, b_mn=True))]:
        t = timeit(fn)
        print(f"{name:28s} {t:7.1f} us {fl / t / 1e6:7.1f} TF/s", flush=True)


def vit():
    """Every GEMM of one ViT-B layer (fwd, dgrad, wgrad) with the epilogue engine.py gives it, at M = 15760 rows."""
    from avt_b200.engine import _split_k_for
    M, D = 15760, 768
    x, att, ln2 = rnd(M, D), rnd(M, D), rnd(M, D)
    wq, wp, w1, w2 = rnd(3 * D, D, scale=0.03), rnd(D, D, scale=0.03), rnd(4 * D, D, scale=0.03), rnd(D, 4 * D, scale=0.03)
    bq, bp, b1, b2 = (torch.randn(n, device=dev) for n in (3 * D, D, 4 * D, D))
    qkv, y, h, z = (torch.empty(M, n, device=dev, dtype=bf) for n in (3 * D, D, 4 * D, 4 * D))
    dy, dz, dqkv, dln = rnd(M, D), rnd(M, 4 * D), rnd(M, 3 * D), torch.empty(M, D, device=dev, dtype=bf)
    dz_out = torch.empty(M, 4 * D, device=dev, dtype=bf)
    gq, gp, g1, g2 = (torch.zeros_like(w, dtype=torch.float32) for w in (wq, wp, w1, w2))
    gb = torch.zeros(4 * D, device=dev)
    x32, xo32 = torch.randn(M, D, device=dev), torch.empty(M, D, device=dev)
    ops.gemm(ln2, w1, h, bias=b1, act=1, aux_z=z, aux_grad=True)

    def wg(dyv, xv, dW, colsum):
        sk = _split_k_for(dyv.shape[1], xv.shape[1], M, 256)
        return lambda: ops.gemm(dyv, xv, dW, a_mn=True, b_mn=True, split_k=sk, accumulate=sk > 1, a_colsum=gb[:dyv.shape[1]] if colsum else None)

    cases = [("fwd qkv", 3 * D * D, lambda: ops.gemm(x, wq, qkv, bias=bq)),
             ("fwd proj", D * D, lambda: ops.gemm(att, wp, y, bias=bp)),
             ("fwd fc1 gelu+aux", 4 * D * D, lambda: ops.gemm(ln2, w1, h, bias=b1, act=1, aux_z=z, aux_grad=True)),
             ("fwd fc2", 4 * D * D, lambda: ops.gemm(h, w2, y, bias=b2)),
             ("fwd proj +residual f32", D * D, lambda: ops.gemm(att, wp, xo32, bias=bp, residual=x32)),
             ("fwd fc2 +residual f32", 4 * D * D, lambda: ops.gemm(h, w2, xo32, bias=b2, residual=x32)),
             ("dgrad fc2 *gelu'", 4 * D * D, lambda: ops.gemm(dy, w2, dz_out, b_mn=True, dact_z=z, dact=1, dact_is_grad=True)),
             ("dgrad fc1", 4 * D * D, lambda: ops.gemm(dz, w1, dln, b_mn=True)),
             ("dgrad proj", D * D, lambda: ops.gemm(dy, wp, dln, b_mn=True)),
             ("dgrad qkv", 3 * D * D, lambda: ops.gemm(dqkv, wq, dln, b_mn=True)),
             ("wgrad fc2", 4 * D * D, wg(dy, h, g2, False)),
             ("wgrad fc1 +bias", 4 * D * D, wg(dz, ln2, g1, True)),
             ("wgrad proj", D * D, wg(dy, att, gp, False)),
             ("wgrad qkv +bias", 3 * D * D, wg(dqkv, x, gq, True))]
    tot_t = tot_f = 0.0
    for name, nk, fn in cases:
        t = timeit(fn)
        fl = 2.0 * M * nk
        tot_t += t
        tot_f += fl
        print(f"{name:20s} {t:7.1f} us {fl / t / 1e6:7.1f} TF/s", flush=True)
    print(f"layer total {tot_t:7.1f} us {tot_f / tot_t / 1e6:7.1f} TF/s  (x12 = {12 * tot_t / 1e3:.2f} ms)")


if __name__ == "__main__":
    todo = sys.argv[1:] or ["wgrad", "head", "ln", "attn", "fc1", "vit"]
    print(torch.cuda.get_device_name(0))
    for name in todo:
        print(f"== {name}")
        globals()[name]()
