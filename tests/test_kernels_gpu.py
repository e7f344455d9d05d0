"""Parity of the non-GEMM kernels (through the C-ABI) against fp64 PyTorch restatements of the same op.

Tolerances: fp32 outputs rel-L2 <= 1e-5 (north star fp32 bound); bf16 outputs within one bf16 rounding
(2^-8 relative) of the fp64 result plus a small absolute floor for values that cancel to ~0.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from avt_b200 import ops
    return ops


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def assert_bf16_close(out, ref, what="", ulps=1.0, floor=1e-3):
    ref = ref.double()
    err = (out.double() - ref).abs()
    tol = ulps * ref.abs() * 2.0**-8 + floor * ref.abs().max().item() * 2.0**-8 * 4
    assert (err <= tol).all(), f"{what}: max err {err.max().item():.3e}, worst ratio {(err / tol).max().item():.2f}"


# three backward variants: bulk-async pipelined (many rows, 256 <= D <= 1024), one CTA per row (<= 296 rows),
# register-resident fallback (everything else)
@pytest.mark.parametrize("rows,D,eps", [(80 * 197, 768, 1e-6), (80, 2048, 1e-5), (33, 64, 1e-6), (10, 1024, 1e-6), (7, 640, 1e-5),
                                        (3001, 768, 1e-6), (120 * 197, 1024, 1e-6), (1500, 264, 1e-5), (1000, 64, 1e-6),
                                        (400, 2048, 1e-5), (297, 512, 1e-6)])
def test_layernorm_fwd_bwd(rows, D, eps):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(rows + D)
    x = torch.randn(rows, D, generator=g, device="cuda") * 2 + 0.5
    gamma = 1 + 0.2 * torch.randn(D, generator=g, device="cuda")
    beta = 0.1 * torch.randn(D, generator=g, device="cuda")
    y16 = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    y32 = torch.empty(rows, D, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, eps, y16, mean, rstd)
    ops.layernorm_fwd(x, gamma, beta, eps, y32)
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd, (D,), gd, bd, eps)
    assert rel(y32, ref) < 1e-5
    assert_bf16_close(y16, ref, "ln fwd bf16")
    assert rel(mean, xd.mean(-1)) < 1e-5
    assert rel(rstd, 1 / torch.sqrt(xd.var(-1, unbiased=False) + eps)) < 1e-5
    # backward: dy in bf16 (as produced by a dgrad GEMM), residual gradient added
    dy = (torch.randn(rows, D, generator=g, device="cuda")).to(torch.bfloat16)
    dres = torch.randn(rows, D, generator=g, device="cuda")
    ref.backward(dy.double())
    dx = torch.empty(rows, D, device="cuda")
    dxb = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    dg, db = torch.empty(D, device="cuda"), torch.empty(D, device="cuda")
    ws = torch.empty(ops.layernorm_bwd_workspace(rows, D), dtype=torch.uint8, device="cuda")
    dcol = torch.full((D,), 7.0, device="cuda")
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dg, db, ws, dx_in=dres, dx_bf16=dxb, dx_colsum=dcol)
    assert rel(dx, xd.grad + dres.double()) < 1e-5
    assert rel(dg, gd.grad) < 1e-5 and rel(db, bd.grad) < 1e-5
    assert rel(dcol, (xd.grad + dres.double()).sum(0)) < 1e-5       # bias gradient of the branch-closing Linear
    assert_bf16_close(dxb, xd.grad + dres.double(), "ln bwd bf16 copy")
    # accumulate mode + fp32 dy + in-place dx
    dg2, db2 = dg.clone(), db.clone()
    dx2 = dres.clone()
    ops.layernorm_bwd(dy.float(), x, mean, rstd, gamma, dx2, dg2, db2, ws, dx_in=dx2, accumulate=True)
    assert rel(dx2, dx) < 1e-6 and rel(dg2, 2 * dg) < 1e-6 and rel(db2, 2 * db) < 1e-6
    # accumulate mode as the block stack uses it (bf16 dy, column sums of dx too): the pipelined kernel adds its per-CTA
    # column sums with atomics into the running gradient vectors, no partials pass
    dg4, db4, dcol4 = dg.clone(), db.clone(), dcol.clone()
    dx4 = torch.empty(rows, D, device="cuda")
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx4, dg4, db4, ws, dx_in=dres, dx_colsum=dcol4, accumulate=True)
    assert rel(dx4, dx) < 1e-6 and rel(dg4, 2 * dg) < 2e-6 and rel(db4, 2 * db) < 2e-6 and rel(dcol4, 2 * dcol) < 2e-6
    # no residual gradient (first LayerNorm backward of a chain)
    dx3 = torch.empty(rows, D, device="cuda")
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx3, dg2, db2, ws)
    assert rel(dx3, xd.grad) < 1e-5 and rel(dg2, dg) < 1e-6


def test_layernorm_fused_residual_add():
    """x' = x + bf16 branch; LN(x'); x' written back (timm Block / GPT2Block residual adds)."""
    ops = _ops()
    rows, D = 1000, 768
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn(rows, D, generator=g, device="cuda")
    add = torch.randn(rows, D, generator=g, device="cuda").to(torch.bfloat16)
    gamma, beta = torch.randn(D, generator=g, device="cuda"), torch.randn(D, generator=g, device="cuda")
    y = torch.empty(rows, D, device="cuda")
    xo = torch.empty(rows, D, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, 1e-6, y, mean, rstd, add=add, x_out=xo)
    xn = x.double() + add.double()
    assert rel(xo, xn) < 1e-7
    assert rel(y, torch.nn.functional.layer_norm(xn, (D,), gamma.double(), beta.double(), 1e-6)) < 1e-5
    # strided variant used for the ViT CLS rows: one row per frame, x_out compact
    F, ntok = 5, 197
    xb = torch.randn(F * ntok, D, generator=g, device="cuda")
    ab = torch.randn(F * ntok, D, generator=g, device="cuda").to(torch.bfloat16)
    yc, xc = torch.empty(F, D, device="cuda"), torch.empty(F, D, device="cuda")
    ops.layernorm_fwd(xb, gamma, beta, 1e-6, yc, rows=F, x_stride=ntok * D, add=ab, add_stride=ntok * D, x_out=xc)
    xn = (xb.double() + ab.double()).view(F, ntok, D)[:, 0]
    assert rel(xc, xn) < 1e-7
    assert rel(yc, torch.nn.functional.layer_norm(xn, (D,), gamma.double(), beta.double(), 1e-6)) < 1e-5


def test_layernorm_strided_cls_rows():
    ops = _ops()
    F, ntok, D = 5, 197, 768
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(F * ntok, D, generator=g, device="cuda")
    gamma, beta = torch.randn(D, generator=g, device="cuda"), torch.randn(D, generator=g, device="cuda")
    y = torch.empty(F, D, device="cuda")
    mean, rstd = torch.empty(F, device="cuda"), torch.empty(F, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, 1e-6, y, mean, rstd, rows=F, x_stride=ntok * D)
    ref = torch.nn.functional.layer_norm(x.view(F, ntok, D)[:, 0].double(), (D,), gamma.double(), beta.double(), 1e-6)
    assert rel(y, ref) < 1e-5


def test_cast_patchify_colsum_framesum():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(4)
    src = torch.randn(1000003, generator=g, device="cuda")
    dst = torch.empty(1000003, device="cuda", dtype=torch.bfloat16)
    ops.cast_bf16(src, dst)
    assert torch.equal(dst, src.to(torch.bfloat16))
    F, C, H, W, ps = 3, 3, 64, 48, 16
    video = torch.randn(F, C, H, W, generator=g, device="cuda")
    P = (H // ps) * (W // ps)
    out = torch.empty(F * (P + 1), C * ps * ps, device="cuda", dtype=torch.bfloat16)
    ops.patchify(video, out, ps)
    ref = torch.nn.functional.unfold(video, ps, stride=ps).transpose(1, 2)  # [F, P, C*ps*ps] in (c, kh, kw) order
    ref = torch.cat([torch.zeros(F, 1, C * ps * ps, device="cuda"), ref], 1).reshape(F * (P + 1), -1)
    assert torch.equal(out, ref.to(torch.bfloat16))
    x = torch.randn(15760, 768, generator=g, device="cuda").to(torch.bfloat16)
    acc = torch.ones(768, device="cuda")
    ops.colsum(x, acc)
    assert rel(acc, x.double().sum(0) + 1) < 1e-5
    Fr, period, D = 7, 197, 256
    dx = torch.randn(Fr * period, D, generator=g, device="cuda")
    ws = torch.empty(period * D, device="cuda")
    dpos, dcls, dbias = torch.empty(period, D, device="cuda"), torch.empty(D, device="cuda"), torch.empty(D, device="cuda")
    ops.frame_sum_grads(dx, Fr, period, D, ws, dpos=dpos, dcls=dcls, dbias=dbias)
    s = dx.double().view(Fr, period, D).sum(0)
    assert rel(dpos, s) < 1e-6 and rel(dcls, s[0]) < 1e-6 and rel(dbias, s[1:].sum(0)) < 1e-5


def test_dropout_apply_matches_gemm_epilogue_mask():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 80, 2048, 64
    a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, generator=g, device="cuda").to(torch.bfloat16)
    plain, dropped = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    ops.gemm(a, b, plain)
    ops.gemm(a, b, dropped, drop_p=0.1, drop_seed=77, drop_offset=5 << 28)
    again = torch.empty(M, N, device="cuda")
    again16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.dropout_apply(plain, 0.1, 77, 5 << 28, y_f32=again, y_bf16=again16)
    assert torch.equal(again, dropped)
    assert torch.equal(again16, dropped.to(torch.bfloat16))


def _attn_ref(qkv, B, H, N, hd, causal, scale, mask=None, keep_scale=1.0):
    D = H * hd
    q, k, v = qkv.double().view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        s = s.masked_fill(~torch.tril(torch.ones(N, N, dtype=torch.bool, device=s.device)), float("-inf"))
    p = s.softmax(-1)
    lse = torch.logsumexp(s, -1)
    if mask is not None:
        p = p * mask * keep_scale
    o = (p @ v).transpose(1, 2).reshape(B * N, D)
    return o, lse.reshape(B * H, N)


@pytest.mark.parametrize("B,H,N,hd,causal", [(8, 4, 10, 512, True), (2, 2, 15, 1024, True), (3, 12, 197, 64, False),
                                              (2, 2, 5, 16, True), (2, 2, 5, 32, False), (2, 8, 10, 256, True)])
def test_attention_simt(B, H, N, hd, causal):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(B * N + hd)
    D = H * hd
    qkv = (torch.randn(B * N, 3 * D, generator=g, device="cuda") * 0.7).to(torch.bfloat16)
    scale = hd ** -0.5
    out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B * H, N, device="cuda")
    ops.attention_simt_fwd(qkv, out, lse, B, H, N, hd, causal=causal, scale=scale)
    qd = qkv.double().requires_grad_(True)
    o_ref, lse_ref = _attn_ref(qd, B, H, N, hd, causal, scale)
    assert rel(lse, lse_ref) < 1e-5
    assert_bf16_close(out, o_ref, "attn out")
    dout = torch.randn(B * N, D, generator=g, device="cuda").to(torch.bfloat16)
    o_ref.backward(dout.double())
    dqkv = torch.empty_like(qkv)
    ops.attention_simt_bwd(qkv, out, dout, lse, dqkv, B, H, N, hd, causal=causal, scale=scale)
    assert rel(dqkv, qd.grad) < 4e-3, rel(dqkv, qd.grad)   # bf16 storage of dq/dk/dv: ~2^-9 rms per element


def test_attention_simt_dropout_consistency():
    """fwd and bwd regenerate the same Philox mask: check against an explicit-mask fp64 reference."""
    ops = _ops()
    B, H, N, hd, p = 4, 4, 10, 64, 0.25
    D = H * hd
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn(B * N, 3 * D, generator=g, device="cuda").to(torch.bfloat16)
    scale = hd ** -0.5
    out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B * H, N, device="cuda")
    ops.attention_simt_fwd(qkv, out, lse, B, H, N, hd, causal=True, scale=scale, drop_p=p, seed=9, offset=1 << 30)
    # recover the mask through the dense dropout kernel (same (seed, offset, linear index) convention)
    ones = torch.ones(B * H * N * N, device="cuda")
    m = torch.empty_like(ones)
    ops.dropout_apply(ones, p, 9, 1 << 30, y_f32=m)
    mask = (m != 0).double().view(B, H, N, N)
    qd = qkv.double().requires_grad_(True)
    o_ref, _ = _attn_ref(qd, B, H, N, hd, True, scale, mask=mask, keep_scale=1 / (1 - p))
    assert_bf16_close(out, o_ref, "attn dropout out")
    dout = torch.randn(B * N, D, generator=g, device="cuda").to(torch.bfloat16)
    o_ref.backward(dout.double())
    dqkv = torch.empty_like(qkv)
    ops.attention_simt_bwd(qkv, out, dout, lse, dqkv, B, H, N, hd, causal=True, scale=scale, drop_p=p, seed=9, offset=1 << 30)
    assert rel(dqkv, qd.grad) < 4e-3
    keep_rate = mask[torch.tril(torch.ones(N, N, dtype=torch.bool, device="cuda")).expand(B, H, N, N)].mean().item()
    assert abs(keep_rate - (1 - p)) < 0.03


@pytest.mark.gpu
def test_preprocess_u8_matches_reference_transform_chain():
    """avt_preprocess_u8 vs the reference's transform tail restated with torch ops (common/transforms.py: to_tensor :124-141,
    hflip :162-170, normalize :144-159, crop :38-44) in the reference's order: ToTensor, flip, * scale, normalize, crop."""
    from avt_b200 import ops
    g = torch.Generator().manual_seed(0)
    F, Hin, Win, h, w = 6, 40, 56, 32, 32
    frames = torch.randint(0, 256, (F, Hin, Win, 3), dtype=torch.uint8, generator=g)
    flip = torch.tensor([0, 1, 0, 1, 1, 0], dtype=torch.uint8)
    mean, std, cy, cx = (0.45, 0.5, 0.4), (0.225, 0.5, 0.25), 5, 17
    clip = frames.permute(0, 3, 1, 2).float() / 255.0                       # (T, C, H, W)
    clip = torch.where(flip.bool()[:, None, None, None], clip.flip(-1), clip)
    clip = (clip * 1.0 - torch.tensor(mean)[None, :, None, None]) / torch.tensor(std)[None, :, None, None]
    ref = clip[..., cy:cy + h, cx:cx + w]
    out = torch.empty(F, 3, h, w, device="cuda")
    ops.preprocess_u8(frames.cuda(), out, cy, cx, flip=flip.cuda(), mean=mean, std=std)
    assert torch.allclose(out.cpu(), ref, rtol=1e-6, atol=1e-6)
    out2 = torch.empty(F, 3, Hin, Win, device="cuda")
    ops.preprocess_u8(frames.cuda(), out2, 0, 0)                             # expts/01: mean = std = 0.5, no flip, no crop
    assert torch.allclose(out2.cpu(), (frames.permute(0, 3, 1, 2).float() / 255.0 - 0.5) / 0.5, rtol=1e-6, atol=1e-6)


def test_zero_and_launch_counter():
    """avt_zero (memset node behind the gradient-buffer zeroing) and avt_kernel_launch_count (bench.py's gpu_launches)."""
    ops = _ops()
    from avt_b200 import _lib
    t = torch.randn(1000, 77, device="cuda")
    ops.zero_(t[:500])
    assert t[:500].abs().max().item() == 0.0 and t[500:].abs().min().item() > 0.0
    n0 = _lib.lib().avt_kernel_launch_count()
    a = torch.randn(64, 128, device="cuda")
    b = torch.empty(64, 128, device="cuda", dtype=torch.bfloat16)
    ops.cast_bf16(a, b)
    ops.cast_bf16(a, b)
    assert _lib.lib().avt_kernel_launch_count() - n0 == 2       # (one kernel per cast at this size; the memset above is not a kernel)
    assert _lib.launch_count == _lib.lib().avt_kernel_launch_count()
