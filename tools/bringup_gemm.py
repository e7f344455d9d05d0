"""GPU bring-up / perf probe for avt_gemm_bf16 (run under gpurun). Not a test: prints diagnostics.

    python tools/bringup_gemm.py all          # every case in its own subprocess (a trap cannot poison the rest)
    python tools/bringup_gemm.py case A B BN  # one operand-major combination with layout probes
    python tools/bringup_gemm.py perf         # TFLOP/s on the ViT-B/16 cfg2 shapes
"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _probe(ops, torch, a_mn, b_mn, bn, M, N, K, kind):
    """kind: 'rand' | 'rowid' | 'kid' (B = identity-like so out reveals the operand mapping)."""
    dev = "cuda"
    if kind == "rand":
        g = torch.Generator(device=dev).manual_seed(1)
        A = torch.randn(M, K, generator=g, device=dev)
        B = torch.randn(N, K, generator=g, device=dev)
    else:
        B = torch.zeros(N, K, device=dev)
        idx = torch.arange(min(N, K), device=dev)
        B[idx, idx] = 1.0
        if kind == "rowid":
            A = (torch.arange(M, device=dev).float() % 128).view(M, 1).expand(M, K).contiguous()
        else:
            A = (torch.arange(K, device=dev).float() % 64).view(1, K).expand(M, K).contiguous()
    A = A.to(torch.bfloat16)
    B = B.to(torch.bfloat16)
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    out = torch.full((M, N), -777.0, device=dev)
    ops.gemm(a, b, out, a_mn=a_mn, b_mn=b_mn, block_n=bn)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = (out.double() - ref).abs()
    rel = (err.norm() / (ref.norm() + 1e-30)).item()
    nbad = (err > 1e-2 * (ref.abs().max().item() + 1e-9)).sum().item()
    print(f"  [{kind}] M{M} N{N} K{K} bn{bn} a_mn{int(a_mn)} b_mn{int(b_mn)}: rel {rel:.3e} bad {nbad}/{M*N}"
          f" untouched {(out == -777.0).sum().item()}", flush=True)
    if nbad and kind != "rand":
        torch.set_printoptions(linewidth=220, precision=0, sci_mode=False)
        print("   got[0:12, 0:24]:\n", out[:12, :24].cpu())
        print("   got[64:70, 0:24]:\n", out[64:70, :24].cpu())
    return nbad == 0


def run_case(a_mn, b_mn, bn):
    import torch
    from avt_b200 import ops
    ok = True
    shapes = [(128, bn, 64), (128, bn, 256), (256, 2 * bn, 128), (200, 768, 768), (80, 2048, 80)]
    for (M, N, K) in shapes:
        for kind in ("rand", "rowid", "kid"):
            if kind != "rand" and (M, N, K) != shapes[0] and ok:
                continue
            ok = _probe(ops, torch, a_mn, b_mn, bn, M, N, K, kind) and ok
    print("CASE", "OK" if ok else "FAIL", a_mn, b_mn, bn, flush=True)
    return ok


def run_perf():
    import torch
    from avt_b200 import ops
    dev = "cuda"
    M = 15760
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    print(torch.cuda.get_device_name(0))
    cases = [
        ("qkv fwd", M, 2304, 768, False, False, 1), ("proj fwd", M, 768, 768, False, False, 1),
        ("fc1 fwd", M, 3072, 768, False, False, 1), ("fc2 fwd", M, 768, 3072, False, False, 1),
        ("fc1 dgrad", M, 768, 3072, False, True, 1), ("fc2 dgrad", M, 3072, 768, False, True, 1),
        ("fc1 wgrad", 3072, 768, M, True, True, 1), ("fc1 wgrad sk4", 3072, 768, M, True, True, 4),
        ("fc2 wgrad sk4", 768, 3072, M, True, True, 4), ("qkv wgrad sk5", 2304, 768, M, True, True, 5),
        ("proj wgrad sk16", 768, 768, M, True, True, 16),
        ("avth c_fc fwd", 80, 8192, 2048, False, True, 1), ("avth c_proj fwd", 80, 2048, 8192, False, True, 1),
    ]
    for name, m, n, k, a_mn, b_mn, sk in cases:
        for bn, cg in ((256, 2), (256, 1), (128, 2)):
            a = torch.randn((k, m) if a_mn else (m, k), device=dev).to(torch.bfloat16)
            b = torch.randn((k, n) if b_mn else (n, k), device=dev).to(torch.bfloat16)
            out = torch.zeros(m, n, device=dev, dtype=torch.float32 if sk > 1 else torch.bfloat16)
            for _ in range(3):
                ops.gemm(a, b, out, a_mn=a_mn, b_mn=b_mn, split_k=sk, block_n=bn, cta_group=cg)
            ts = []
            for _ in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.gemm(a, b, out, a_mn=a_mn, b_mn=b_mn, split_k=sk, block_n=bn, cta_group=cg)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            t = ts[len(ts) // 2]
            # cuBLAS comparison on the same logical problem
            A = a.t() if a_mn else a
            Bt = b if b_mn else b.t()
            for _ in range(3):
                torch.matmul(A, Bt)
            tc = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(A, Bt)
                e1.record()
                torch.cuda.synchronize()
                tc.append(e0.elapsed_time(e1))
            tc.sort()
            fl = 2.0 * m * n * k
            print(f"{name:18s} bn{bn} cg{cg} M{m} N{n} K{k}: {t*1e3:8.1f} us {fl/t/1e9:8.1f} TF/s | cuBLAS {tc[2]*1e3:8.1f} us"
                  f" {fl/tc[2]/1e9:8.1f} TF/s", flush=True)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    if mode == "case":
        ok = run_case(bool(int(sys.argv[2])), bool(int(sys.argv[3])), int(sys.argv[4]))
        sys.exit(0 if ok else 1)
    if mode == "perf":
        run_perf()
        return
    os.makedirs("gpurun_out", exist_ok=True)
    results = []
    t0 = time.time()
    for a_mn, b_mn in ((0, 0), (0, 1), (1, 0), (1, 1)):
        for bn in (64, 256, 128):
            try:
                r = subprocess.run([sys.executable, __file__, "case", str(a_mn), str(b_mn), str(bn)], capture_output=True,
                                   text=True, timeout=240)
                out, rc = r.stdout + r.stderr[-3000:], r.returncode
            except subprocess.TimeoutExpired as e:
                out, rc = f"TIMEOUT {e}", -9
            results.append((a_mn, b_mn, bn, rc))
            print(f"##### a_mn={a_mn} b_mn={b_mn} bn={bn} rc={rc} t={time.time()-t0:.0f}s\n{out}", flush=True)
    print("SUMMARY", results)
    if all(rc == 0 for *_, rc in results[:3]):
        try:
            r = subprocess.run([sys.executable, __file__, "perf"], capture_output=True, text=True, timeout=600)
            print(r.stdout + r.stderr[-3000:])
        except subprocess.TimeoutExpired as e:
            print("perf TIMEOUT", e)


if __name__ == "__main__":
    main()
