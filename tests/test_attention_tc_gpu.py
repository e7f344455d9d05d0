"""tcgen05 ViT attention vs an fp64 restatement of timm Attention's core (same bf16-rounded qkv)."""
import pytest
import torch

from test_kernels_gpu import _attn_ref, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("F,H,N", [(1, 1, 197), (3, 12, 197), (2, 16, 197), (2, 2, 5), (2, 3, 128), (1, 2, 208), (80, 12, 197)])
def test_attention_tc_fwd(F, H, N):
    from avt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(F * 1000 + N)
    D = H * 64
    qkv = (torch.randn(F * N, 3 * D, generator=g, device="cuda") * 1.5).to(torch.bfloat16)
    out = torch.full((F * N, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((F * H, N), float("nan"), device="cuda")
    ops.attention_tc_fwd(qkv, out, lse, F, H, N, scale=0.125)
    torch.cuda.synchronize()
    o_ref, lse_ref = _attn_ref(qkv, F, H, N, 64, False, 0.125)
    assert rel(lse, lse_ref) < 1e-5, rel(lse, lse_ref)
    # P is rounded to bf16 before the PV MMA (as in every flash-attention kernel) and O is stored as bf16:
    # two roundings of 2^-9 relative each -> rel-L2 <= 3e-3, and no element off by more than 2^-6 of the scale.
    assert rel(out, o_ref) < 3e-3, rel(out, o_ref)
    assert (out.double() - o_ref).abs().max().item() <= 2.0**-6 * o_ref.abs().max().item()
    assert not torch.isnan(out.float()).any()


@pytest.mark.parametrize("F,H,N", [(1, 1, 197), (3, 12, 197), (2, 2, 5), (2, 3, 128), (2, 2, 130), (1, 2, 208), (80, 12, 197)])
def test_attention_tc_bwd(F, H, N):
    from avt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(F * 1000 + N + 1)
    D = H * 64
    qkv = (torch.randn(F * N, 3 * D, generator=g, device="cuda") * 1.2).to(torch.bfloat16)
    out = torch.empty(F * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(F * H, N, device="cuda")
    ops.attention_tc_fwd(qkv, out, lse, F, H, N, scale=0.125)
    dout = torch.randn(F * N, D, generator=g, device="cuda").to(torch.bfloat16)
    dqkv = torch.full((F * N, 3 * D), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attention_tc_bwd(qkv, out, dout, lse, dqkv, F, H, N, scale=0.125)
    torch.cuda.synchronize()
    assert not torch.isnan(dqkv.float()).any()
    if F * H <= 64:   # fp64 autograd reference
        qd = qkv.double().requires_grad_(True)
        o_ref, _ = _attn_ref(qd, F, H, N, 64, False, 0.125)
        o_ref.backward(dout.double())
        ref = qd.grad
    else:             # full size: the CUDA-core kernel (itself checked against fp64 above) is the reference
        ref = torch.empty_like(dqkv)
        ops.attention_simt_bwd(qkv, out, dout, lse, ref, F, H, N, 64, causal=False, scale=0.125)
    for s, name in enumerate(("dq", "dk", "dv")):
        r = rel(dqkv[:, s * D:(s + 1) * D], ref[:, s * D:(s + 1) * D])
        # P, dS rounded to bf16 for the MMAs + bf16 storage of the result (+ bf16 reference at full size)
        assert r < 6e-3, (name, r)
