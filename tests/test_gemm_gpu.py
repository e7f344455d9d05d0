"""Parity of the tcgen05 GEMM (avt_gemm_bf16) against fp64 matmul on the same bf16-rounded operands.

Tolerances (stated per the north star: bf16 path <= 1e-3 relative):
  * fp32 output: rel-L2 <= 1e-5 (only fp32 accumulation-order noise; operands are exactly representable).
  * bf16 output: every element within one bf16 rounding of the fp64 result: |err| <= 2^-8 |ref| + 1e-6 * scale.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from avt_b200 import ops
    return ops


def _mk(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda", dtype=torch.float32) * scale).to(torch.bfloat16)


def _ref(a, b, a_mn, b_mn):
    A = a.double().t() if a_mn else a.double()
    B = b.double().t() if b_mn else b.double()
    return A @ B.t()


def _check(out, ref, what="", fp32_tol=1e-5):
    ref = ref.double()
    o = out.double()
    scale = ref.abs().max().item() + 1e-30
    if out.dtype == torch.float32:
        rel = ((o - ref).norm() / (ref.norm() + 1e-30)).item()
        assert rel <= fp32_tol, f"{what}: rel-L2 {rel:.3e}"
    else:
        err = (o - ref).abs()
        tol = ref.abs() * 2.0**-8 + 1e-6 * scale
        bad = (err > tol).sum().item()
        assert bad == 0, f"{what}: {bad} elements off by more than one bf16 ulp; max err {err.max().item():.3e} scale {scale:.3e}"


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K,bn,cg", [
    (128, 256, 64, 256, 1), (128, 64, 64, 64, 1), (256, 512, 256, 256, 1), (384, 384, 192, 128, 1),
    (200, 768, 768, 256, 1), (80, 2048, 768, 64, 1), (1000, 1024, 80, 128, 1), (15760, 768, 768, 256, 1),
    # CTA pairs (tcgen05 cta_group::2): 256-row pair tiles, ragged M, both block widths
    (256, 256, 64, 256, 2), (512, 512, 256, 256, 2), (200, 768, 768, 256, 2), (1000, 1024, 80, 128, 2),
    (384, 384, 192, 128, 2), (15760, 768, 768, 256, 2), (15760, 2304, 768, 256, 2),
])
def test_gemm_layouts(a_mn, b_mn, M, N, K, bn, cg):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = _mk((K, M) if a_mn else (M, K), g)
    b = _mk((K, N) if b_mn else (N, K), g)
    if a_mn and M % 8:
        pytest.skip("transposed A needs M % 8 == 0 (16-byte row pitch)")
    for dt in (torch.float32, torch.bfloat16):
        out = torch.full((M, N), float("nan"), device="cuda", dtype=dt)
        ops.gemm(a, b, out, a_mn=a_mn, b_mn=b_mn, block_n=bn, cta_group=cg)
        torch.cuda.synchronize()
        _check(out, _ref(a, b, a_mn, b_mn), f"M{M} N{N} K{K} bn{bn} cg{cg} a_mn{a_mn} b_mn{b_mn} {dt}")


def _gelu_tanh(x):
    return 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * x**3)))


@pytest.mark.parametrize("act", [1, 2])
@pytest.mark.parametrize("cg,M", [(1, 300), (2, 1500)])
@pytest.mark.parametrize("out_dt", [torch.bfloat16, torch.float32])
def test_gemm_bias_act_aux(act, cg, M, out_dt):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    N, K = 512, 256
    a, b = _mk((M, K), g), _mk((N, K), g, 0.1)
    bias = torch.randn(N, generator=g, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=out_dt)
    z = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, out, bias=bias, act=act, aux_z=z, cta_group=cg)
    pre = (_ref(a, b, False, False) + bias.double()).requires_grad_(True)
    _check(z, pre.detach(), "aux_z")
    want = torch.nn.functional.gelu(pre) if act == 1 else _gelu_tanh(pre)
    err = (out.double() - want).abs()
    assert (err <= want.abs() * 2.0**-7 + 2e-3).all(), err.max().item()
    if out_dt == torch.float32:     # branch-free erf / tanh formulations: abs error ~1e-6
        assert err.max().item() < 2e-5 * max(1.0, want.abs().max().item())
    # aux_mode = 1: the epilogue saves act'(pre-activation) instead
    ops.gemm(a, b, out, bias=bias, act=act, aux_z=z, aux_grad=True, cta_group=cg)
    (dwant,) = torch.autograd.grad(want.sum(), pre)
    derr = (z.double() - dwant).abs()
    assert (derr <= dwant.abs() * 2.0**-8 + 1e-3).all(), derr.max().item()
    # ... and a backward GEMM multiplies by it directly (dact_mode = 1)
    g2 = _mk((M, K), g)
    dz = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(g2, b, dz, dact_z=z, dact_is_grad=True, cta_group=cg)
    ref = _ref(g2, b, False, False) * z.double()
    assert ((dz.double() - ref).norm() / ref.norm()).item() < 1e-5


@pytest.mark.parametrize("act", [1, 2])
def test_gemm_dact(act):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(6)
    M, N, K = 256, 256, 128
    a, b = _mk((M, K), g), _mk((N, K), g, 0.1)
    z = _mk((M, N), g)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(a, b, out, dact_z=z, dact=act)
    zz = z.double().requires_grad_(True)
    y = torch.nn.functional.gelu(zz) if act == 1 else _gelu_tanh(zz)
    (dz,) = torch.autograd.grad(y.sum(), zz)
    want = _ref(a, b, False, False) * dz
    rel = ((out.double() - want).norm() / want.norm()).item()
    assert rel < 1e-4, rel


def test_gemm_residual_accumulate_splitk():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(7)
    M, N, K = 384, 768, 4096
    a, b = _mk((K, M), g), _mk((K, N), g, 0.05)
    ref = _ref(a, b, True, True)
    res = torch.randn(M, N, generator=g, device="cuda")
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, b, out, a_mn=True, b_mn=True, residual=res)
    _check(out, ref + res.double(), "residual")
    out = res.clone()
    ops.gemm(a, b, out, a_mn=True, b_mn=True, accumulate=True)
    _check(out, ref + res.double(), "accumulate")
    out = res.clone()
    ops.gemm(a, b, out, a_mn=True, b_mn=True, split_k=7, accumulate=True)      # atomics onto the existing values
    _check(out, ref + res.double(), "split_k accumulate")
    out = res.clone()
    ops.gemm(a, b, out, a_mn=True, b_mn=True, split_k=7)                       # overwrite: zero-filled first
    _check(out, ref, "split_k overwrite")
    ws = torch.empty(7 * M * N, device="cuda")
    o1, o2 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    ops.gemm(a, b, o1, a_mn=True, b_mn=True, split_k=7, workspace=ws)          # slices + fixed-order sum
    ops.gemm(a, b, o2, a_mn=True, b_mn=True, split_k=7, workspace=ws)
    _check(o1, ref, "split_k slices")
    assert torch.equal(o1, o2)                                                 # bit-reproducible


@pytest.mark.parametrize("M,N,K,bn,cg,sk,b_mn", [
    (3072, 768, 15760, 256, 2, 2, True),      # timm Mlp.fc1 weight gradient at the BASELINE shape (dW = dz^T ln2)
    (768, 768, 15760, 256, 2, 8, True),       # proj: 9 pair tiles, split 8
    (2304, 768, 4000, 256, 2, 5, True),       # K tail (4000 = 62.5 k-blocks), uneven split
    (200, 512, 1000, 128, 1, 3, True),        # ragged M (200 rows in 2 tiles), single CTAs, N tile > 1
    (384, 256, 320, 256, 1, 1, False),        # no split, K-major B
    (640, 1024, 2048, 256, 2, 4, True),       # ragged pair tile (640 = 2.5 x 256)
])
def test_gemm_wgrad_fused_bias_colsum(M, N, K, bn, cg, sk, b_mn):
    """a_colsum[m] += sum_k A[m, k] while the weight gradient dW = A B^T is computed (A = dY^T, MN-major)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = _mk((K, M), g)                                  # dY [rows, out_features]
    b = _mk((K, N) if b_mn else (N, K), g, 0.05)
    out = torch.zeros(M, N, device="cuda")
    col = torch.full((M,), 3.0, device="cuda")          # accumulates onto the existing values
    ops.gemm(a, b, out, a_mn=True, b_mn=b_mn, split_k=sk, block_n=bn, cta_group=cg, accumulate=sk > 1, a_colsum=col)
    _check(out, _ref(a, b, True, b_mn), "wgrad with fused colsum")
    ref = a.double().sum(0) + 3.0
    assert ((col.double() - ref).norm() / ref.norm()).item() <= 1e-5
    # and the GEMM result is unchanged w.r.t. the plain path
    out2 = torch.zeros(M, N, device="cuda")
    ops.gemm(a, b, out2, a_mn=True, b_mn=b_mn, split_k=1, block_n=bn, cta_group=cg)
    # one unsplit fp32 accumulation chain over K = 15 760 products: rounding noise grows ~sqrt(K) * 2^-24
    _check(out2, _ref(a, b, True, b_mn), "plain wgrad", fp32_tol=1e-5 if K <= 4096 else 4e-5)


def test_gemm_pos_cls():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(8)
    frames, period, N, K = 3, 197, 256, 128
    M = frames * period
    a, b = _mk((M, K), g), _mk((N, K), g, 0.1)
    bias = torch.randn(N, generator=g, device="cuda")
    pos = torch.randn(period, N, generator=g, device="cuda")
    cls = torch.randn(N, generator=g, device="cuda")
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, b, out, bias=bias, pos=pos, cls=cls, pos_period=period)
    want = (_ref(a, b, False, False) + bias.double()).view(frames, period, N)
    want[:, 0] = cls.double()
    want = want + pos.double()
    _check(out, want.view(M, N), "pos/cls")


def test_gemm_dropout_statistics():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(9)
    M, N, K = 512, 1024, 64
    a, b = _mk((M, K), g), _mk((N, K), g)
    out0 = torch.empty(M, N, device="cuda")
    ops.gemm(a, b, out0)
    p = 0.1
    out1 = torch.empty(M, N, device="cuda")
    out2 = torch.empty(M, N, device="cuda")
    ops.gemm(a, b, out1, drop_p=p, drop_seed=123, drop_offset=1000)
    ops.gemm(a, b, out2, drop_p=p, drop_seed=123, drop_offset=1000)
    assert torch.equal(out1, out2)  # same (seed, offset) -> same mask
    kept = out1 != 0
    rate = 1.0 - kept.float().mean().item()
    assert abs(rate - p) < 0.005, rate
    assert torch.allclose(out1[kept], out0[kept] / (1 - p), rtol=1e-5, atol=1e-6)
    out3 = torch.empty(M, N, device="cuda")
    ops.gemm(a, b, out3, drop_p=p, drop_seed=124, drop_offset=1000)
    assert not torch.equal(out1 != 0, out3 != 0)


@pytest.mark.parametrize("b_mn", [False, True])
def test_gemm_small_m_split_k_two_pass(b_mn):
    """AVT-h weight-streaming GEMMs (M = 80): split-K partials in an fp32 workspace + finishing epilogue kernel
    must equal the single-pass fused epilogue."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(21)
    M, N, K = 80, 2048, 8192
    a = _mk((M, K), g)
    b = _mk((K, N) if b_mn else (N, K), g, 0.02)
    bias = torch.randn(N, generator=g, device="cuda")
    ws = torch.empty(8 * M * N, device="cuda")
    outs = []
    for sk in (1, 4):
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, b, out, b_mn=b_mn, bias=bias, act=2, aux_z=aux, aux_grad=True, drop_p=0.1, drop_seed=5, drop_offset=7 << 28,
                 split_k=sk, workspace=ws if sk > 1 else None)
        outs.append((out.float(), aux.float()))
    pre = _ref(a, b, False, b_mn) + bias.double()
    assert ((outs[0][1].double() - outs[1][1].double()).abs() <= 2.0**-7 * outs[0][1].double().abs() + 1e-3).all()
    kept = (outs[0][0] != 0) & (outs[1][0] != 0)
    assert torch.equal(outs[0][0] != 0, outs[1][0] != 0)                      # same dropout mask
    want = _gelu_tanh(pre) / 0.9
    for o, _ in outs:
        err = (o.double() - want).abs()[kept]
        assert (err <= want.abs()[kept] * 2.0**-7 + 2e-3).all()
    # fp32 output through the TMA-store path (no split) and through the two-pass path
    o1, o2 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    ops.gemm(a, b, o1, b_mn=b_mn, bias=bias)
    ops.gemm(a, b, o2, b_mn=b_mn, bias=bias, split_k=8, workspace=ws)
    _check(o1, pre, "fp32 tma store")
    _check(o2, pre, "fp32 two-pass")


@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("kind", ["store", "store_bias", "gelu_aux", "mul_z"])
def test_gemm_specialized_epilogues_match_generic(kind, b_mn):
    """The compile-time specialised ViT epilogues (kEpiStore / kEpiGeluAux / kEpiMulZ, 256-wide CTA-pair kernel) must be
    bit-identical to the generic run-time epilogue on the same inputs, and both must match the fp64 reference."""
    ops = _ops()
    from avt_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(41)
    M, N, K = 1000, 768, 320                     # ragged M (4 pair tiles, the last one 232 rows), 3 N tiles, 5 k-blocks
    a = _mk((M, K), g)
    b = _mk((K, N) if b_mn else (N, K), g, 0.1)
    bias = torch.randn(N, generator=g, device="cuda") if kind in ("store_bias", "gelu_aux") else None
    z = _mk((M, N), g) if kind == "mul_z" else None
    outs = []
    for special in (1, 0):
        _lib.lib().avt_set_gemm_specialized_epilogues(special)
        out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        aux = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16) if kind == "gelu_aux" else None
        ops.gemm(a, b, out, b_mn=b_mn, bias=bias, act=1 if kind == "gelu_aux" else 0, aux_z=aux, aux_grad=aux is not None,
                 dact_z=z, dact_is_grad=z is not None)
        outs.append((out, aux))
    _lib.lib().avt_set_gemm_specialized_epilogues(1)
    assert torch.equal(outs[0][0], outs[1][0])
    pre = _ref(a, b, False, b_mn) + (bias.double() if bias is not None else 0.0)
    if kind == "gelu_aux":
        assert torch.equal(outs[0][1], outs[1][1])
        want = torch.nn.functional.gelu(pre)
        dwant = 0.5 * (1 + torch.erf(pre / 2 ** 0.5)) + pre * torch.exp(-pre * pre / 2) / (2 * torch.pi) ** 0.5
        assert ((outs[0][1].double() - dwant).abs() <= 2.0**-7 * dwant.abs() + 1e-3).all()
    elif kind == "mul_z":
        want = pre * z.double()
    else:
        want = pre
    assert ((outs[0][0].double() - want).abs() <= 2.0**-7 * want.abs() + 1e-3).all()


@pytest.mark.gpu
@pytest.mark.parametrize("b_mn,with_bias", [(False, True), (True, False)])
def test_gemm_residual_epilogue_matches_generic(b_mn, with_bias):
    """kEpiResidual (branch-closing Linear: fp32 out = acc + bias + fp32 residual, both through 16-column TMA boxes) against the
    generic run-time epilogue (bit-identical: the same fp32 additions in the same order) and the fp64 reference."""
    ops = _ops()
    from avt_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(43)
    M, N, K = 1000, 768, 320                     # ragged M (the last pair tile has 232 rows), 3 N tiles, 5 k-blocks
    a = _mk((M, K), g)
    b = _mk((K, N) if b_mn else (N, K), g, 0.1)
    bias = torch.randn(N, generator=g, device="cuda") if with_bias else None
    res = torch.randn(M, N, generator=g, device="cuda") * 3
    outs = []
    for special in (1, 0):
        _lib.lib().avt_set_gemm_specialized_epilogues(special)
        out = torch.full((M, N), float("nan"), device="cuda")
        ops.gemm(a, b, out, b_mn=b_mn, bias=bias, residual=res)
        outs.append(out)
    _lib.lib().avt_set_gemm_specialized_epilogues(1)
    want = _ref(a, b, False, b_mn) + (bias.double() if bias is not None else 0.0) + res.double()
    assert not torch.isnan(outs[0]).any()
    rel = lambda x, y: ((x.double() - y).norm() / y.norm()).item()
    assert rel(outs[0], want) < 1e-5 and rel(outs[1], want) < 1e-5
    assert (outs[0] - outs[1]).abs().max() <= 1e-5 * want.abs().max()
    # in place: the residual buffer is also the output (x += branch)
    x = res.clone()
    ops.gemm(a, b, x, b_mn=b_mn, bias=bias, residual=x)
    assert torch.equal(x, outs[0])


@pytest.mark.parametrize("N,K,sk,b_mn", [(2048, 2048, 8, True), (8192, 2048, 2, True), (2048, 8192, 8, False), (768, 2048, 8, True),
                                         (2048, 768, 4, True), (6144, 2048, 4, False)])
def test_gemm_small_m_cluster_splitk(N, K, sk, b_mn):
    """AVT-h shapes (M = 80 rows): split factors 2 / 4 / 8 run as one thread-block cluster per output tile and reduce through
    distributed shared memory inside the kernel. Checked against fp64, against the unsplit kernel, for bit-reproducibility, and
    with every epilogue the head uses (bias + gelu_new + saved derivative, x saved derivative, fp32 out + residual, dropout)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(N + K + sk)
    M = 80
    a = _mk((M, K), g)
    b = _mk((K, N) if b_mn else (N, K), g, 0.05)
    bias = torch.randn(N, generator=g, device="cuda")
    ws = torch.empty(sk * M * N, device="cuda")
    ref = _ref(a, b, False, b_mn) + bias.double()
    o1 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    o2, o0 = torch.empty_like(o1), torch.empty_like(o1)
    ops.gemm(a, b, o1, b_mn=b_mn, bias=bias, split_k=sk, workspace=ws)
    ops.gemm(a, b, o2, b_mn=b_mn, bias=bias, split_k=sk, workspace=ws)
    ops.gemm(a, b, o0, b_mn=b_mn, bias=bias)
    _check(o1, ref, f"cluster split-K {sk}")
    assert torch.equal(o1, o2)                                   # fixed summation order
    assert (o1.float() - o0.float()).abs().max() <= 2.0**-7 * ref.abs().max()
    # gelu_new + saved derivative, then a backward GEMM that multiplies by it
    h = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    z = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, h, b_mn=b_mn, bias=bias, act=2, aux_z=z, aux_grad=True, split_k=sk, workspace=ws)
    pre = ref.clone().requires_grad_(True)
    want = _gelu_tanh(pre)
    (dwant,) = torch.autograd.grad(want.sum(), pre)
    assert ((h.double() - want.detach()).abs() <= want.detach().abs() * 2.0**-7 + 2e-3).all()
    assert ((z.double() - dwant).abs() <= dwant.abs() * 2.0**-7 + 2e-3).all()
    dz = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, dz, b_mn=b_mn, dact_z=z, dact_is_grad=True, split_k=sk, workspace=ws)
    want2 = _ref(a, b, False, b_mn) * z.double()
    assert ((dz.double() - want2).abs() <= want2.abs() * 2.0**-7 + 2e-3 * want2.abs().max()).all()
    # fp32 out + residual; dropout mask identical to the unsplit kernel's (Philox index = element index)
    res = torch.randn(M, N, generator=g, device="cuda")
    of = torch.empty(M, N, device="cuda")
    ops.gemm(a, b, of, b_mn=b_mn, bias=bias, residual=res, split_k=sk, workspace=ws)
    assert ((of.double() - (ref + res.double())).norm() / (ref + res.double()).norm()).item() < 1e-5
    d1, d0 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    ops.gemm(a, b, d1, b_mn=b_mn, bias=bias, drop_p=0.25, drop_seed=11, drop_offset=5, split_k=sk, workspace=ws)
    ops.gemm(a, b, d0, b_mn=b_mn, bias=bias, drop_p=0.25, drop_seed=11, drop_offset=5)
    assert torch.equal(d1 == 0, d0 == 0)
    keep = d1 != 0
    assert 0.70 < keep.float().mean().item() < 0.80
    assert ((d1.double() - ref / 0.75)[keep].abs().max() / ref.abs().max()).item() < 1e-5
