"""Module-level parity of the drop-in backbone / head (CUDA, bf16 tensor-core math) against the CPU oracle.

Tolerance policy (stated once, used below). The north star asks for 1e-3 relative on the bf16 path; that bound is
met at kernel level on identically-rounded operands (tests/test_gemm_gpu.py, test_kernels_gpu.py). At module level
every GEMM operand is *stored* in bf16 (2^-9 relative rounding each), so an L-layer stack cannot be closer than a few
1e-3 to an fp64 oracle - torch's own bf16 autocast of the oracle lands at 5e-3..1e-2 on these configs. The module tests
therefore require, per configuration, (a) rel-L2 <= 1.5x the error this path was MEASURED to have on B200 (table below),
for outputs and for the worst parameter gradient, AND (b) that neither exceeds 1.25x the error of torch-autocast-bf16 of
the oracle on the same inputs (we are consistently below it: the residual stream, LayerNorm statistics and softmax stay
in fp32). A kernel regression that raises an error by 50 % fails.
"""
import copy
import json
import os

import pytest
import torch

from oracle import avth as o_avth
from oracle import base_model as o_base
from oracle import vit as o_vit

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# Loose bounds for the few comparisons without a recorded error (2x the worst recorded one) ...
OUT_TOL, GRAD_TOL = 2e-2, 4e-2
# ... and per configuration (output, worst parameter gradient) rel-L2 bounds = 1.5x the errors measured on B200
# (tests record them in gpurun_out/parity_errors.json; round 2: outputs 5.0e-3..9.2e-3, gradients 7.0e-3..1.5e-2, always
# below torch's own bf16 autocast of the oracle on the same inputs, which every test also requires).
GOLDEN_AVTH_TOL, GOLDEN_BASE_TOL = (9e-3, 1.3e-2), (1.5e-2, 2.9e-2)
AVTH_TOL = {(64, 32, 5): (9.1e-3, 2.3e-2), (64, 128, 10): (8.7e-3, 1.55e-2), (2048, 768, 10): (8.9e-3, 1.55e-2),
            (768, 2048, 10): (9.9e-3, 1.5e-2), (768, 2048, 15): (7.8e-3, 1.35e-2)}
BACKBONE_TOL = {("vit_test_patch16_32", "stress"): (1.05e-2, 1.9e-2), ("vit_test_patch16_64", "stress"): (1.4e-2, 2.2e-2),
                ("vit_test_patch16_64", "default"): (7.5e-3, 1.05e-2), ("vit_base_patch16_224", "stress"): (9.6e-3, 2.1e-2),
                ("vit_large_patch16_224", "stress"): (9.2e-3, 2.0e-2)}


def record(name, value):
    """Measured errors -> gpurun_out/parity_errors.json (when the directory exists): the tolerances are kept at <= 1.5x these."""
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if not os.path.isdir(d):
        return
    path = os.path.join(d, "parity_errors.json")
    try:
        data = json.load(open(path))
    except Exception:
        data = {}
    data[name] = value
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def stress_init(m, seed=0):
    """O(1) attention logits / activations (the reference init N(0, 0.01) would hide attention bugs)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() >= 2 and not any(k in n for k in ("pos_embed", "cls_token", "wpe")):
                conv1d = any(k in n for k in ("c_attn", "c_fc", "c_proj"))
                fan_in = p.shape[0] if conv1d else p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)
            elif "norm" in n or "ln_" in n:
                p.copy_((1.0 if n.endswith("weight") else 0.0) + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))


def _vit_pair(model_type, init):
    from avt_b200 import backbone
    torch.manual_seed(0)
    ref = o_vit.create_model(model_type)
    if init == "stress":
        stress_init(ref)
    ours = backbone.create_model(model_type)
    ours.load_state_dict(ref.state_dict())
    return ours.cuda(), ref


@pytest.mark.parametrize("model_type,F,init,dtype", [
    ("vit_test_patch16_32", 3, "stress", torch.float64), ("vit_test_patch16_64", 2, "stress", torch.float64),
    ("vit_test_patch16_64", 5, "default", torch.float64), ("vit_base_patch16_224", 2, "stress", torch.float32),
    ("vit_large_patch16_224", 1, "stress", torch.float32),       # BASELINE cfg5 backbone: D 1024, 16 heads, 24 layers
])
def test_backbone_forward_backward_vs_oracle(model_type, F, init, dtype):
    ours, ref = _vit_pair(model_type, init)
    img = o_vit.CONFIGS[model_type][0]
    x = torch.randn(F, 3, img, img, generator=torch.Generator().manual_seed(1))
    gy = torch.randn(F, ref.embed_dim, generator=torch.Generator().manual_seed(2))
    # what stock PyTorch gives for the same module under bf16 autocast (forward AND gradients): the yardstick
    auto = copy.deepcopy(ref).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ya = auto(x.cuda()).float()
    ya.backward(gy.cuda())
    ref = ref.cpu().to(dtype)
    yr = ref(x.to(dtype))
    yr.backward(gy.to(dtype))
    yo = ours(x.cuda())
    yo.backward(gy.cuda())
    e_ours, e_autocast = rel(yo, yr), rel(ya, yr)
    gr, ga = dict(ref.named_parameters()), dict(auto.named_parameters())
    g_ours = {n: rel(p.grad, gr[n].grad) for n, p in ours.named_parameters()}
    g_auto = {n: rel(ga[n].grad, gr[n].grad) for n in g_ours}
    worst = max(g_ours, key=g_ours.get)
    record(f"backbone/{model_type}/{init}", dict(fwd=e_ours, fwd_autocast=e_autocast, grad=g_ours[worst], grad_name=worst,
                                                grad_autocast=max(g_auto.values())))
    out_tol, grad_tol = BACKBONE_TOL[(model_type, init)]
    assert e_ours <= out_tol and e_ours <= 1.25 * e_autocast + 1e-4, (e_ours, e_autocast)
    assert all(p.grad is not None for p in ours.parameters())
    assert g_ours[worst] <= grad_tol and g_ours[worst] <= 1.25 * max(g_auto.values()) + 1e-4, (worst, g_ours[worst], max(g_auto.values()))


def test_backbone_timm_wrapper_contract_and_no_grad_path():
    """(B, C, T, H, W) -> (B, C', T, 1, 1) as models/video_classification.py:213-227; eval under no_grad
    (func/train.py:357) uses the activation-free workspace and must agree with the training-mode forward."""
    from avt_b200 import backbone
    m = backbone.TIMMModel(1, "vit_test_patch16_32").cuda()
    video = torch.randn(4, 3, 2, 32, 32, device="cuda")
    y = m(video)
    assert y.shape == (4, 64, 2, 1, 1) and y.requires_grad
    with torch.no_grad():
        y2 = m(video)
    assert torch.equal(y, y2)
    # frame (b, t) of the clip layout equals the same image pushed through the ViT alone
    single = m.model(video[1, :, 1].unsqueeze(0).contiguous())
    assert torch.equal(single[0], y[1, :, 1, 0, 0])


def test_backbone_frames_are_independent_at_full_size():
    """Size-independent property at the BASELINE shape: 80 frames of ViT-B/16 — every frame's feature is bit-identical
    to running that frame in a 2-frame batch (no cross-frame leakage through tiles / padding rows)."""
    from avt_b200 import backbone
    torch.manual_seed(3)
    m = backbone.create_model("vit_base_patch16_224").cuda()
    stress_init(m, 4)
    x = torch.randn(80, 3, 224, 224, device="cuda")
    with torch.no_grad():
        full = m(x)
        sub = m(x[[0, 79]].contiguous())
    assert torch.isfinite(full).all()
    assert torch.equal(full[[0, 79]], sub)


def _avth_pair(C, Dh, nh, nl, **extra):
    from avt_b200 import future_prediction as fp
    torch.manual_seed(0)
    kw = dict(output_len=1, inter_dim=Dh, n_head=nh, n_layer=nl, return_past_too=True, avg_last_n=1, **extra)
    ref = o_avth.AVTh(C, future_pred_loss="mse", **kw)
    stress_init(ref)
    ours = fp.AVTh(C, future_pred_loss={"_target_": "torch.nn.MSELoss"}, future_pred_loss_wt=1.0, **kw)
    ours.load_state_dict(ref.state_dict())
    return ours.cuda(), ref


@pytest.mark.parametrize("C,Dh,nh,nl,B,T,dtype", [
    (64, 32, 2, 2, 2, 5, torch.float64), (64, 128, 2, 3, 3, 10, torch.float64), (2048, 768, 4, 6, 2, 10, torch.float32),
    (768, 2048, 4, 6, 8, 10, torch.float32), (768, 2048, 8, 2, 2, 15, torch.float32),
])
def test_avth_forward_backward_vs_oracle(C, Dh, nh, nl, B, T, dtype):
    ours, ref = _avth_pair(C, Dh, nh, nl)
    ref = ref.to(dtype).eval()
    ours.eval()   # dropout off: RNG streams cannot be matched; see test_avth_dropout_* for the train path
    x = torch.randn(B, T, C, generator=torch.Generator().manual_seed(1))
    g1 = torch.randn(B, T, C, generator=torch.Generator().manual_seed(2))
    g2 = torch.randn(B, C, generator=torch.Generator().manual_seed(3))
    xr = x.clone().to(dtype).requires_grad_(True)
    pr, fr, lr, _ = ref(xr, (B,))
    ((pr * g1.to(dtype)).sum() + (fr * g2.to(dtype)).sum() + lr["feat"].mean()).backward()
    xo = x.clone().cuda().requires_grad_(True)
    po, fo, lo, _ = ours(xo, (B,))
    ((po * g1.cuda()).sum() + (fo * g2.cuda()).sum() + lo["feat"].mean()).backward()
    assert po.shape == (B, T, C) and fo.shape == (B, C) and lo["feat"].shape == (B, T - 1, C)
    # stock PyTorch bf16 autocast of the same module on the same inputs: the yardstick
    auto = copy.deepcopy(ref).float().cuda()
    xa = x.clone().cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        pa, fa, la, _ = auto(xa, (B,))
        ((pa.float() * g1.cuda()).sum() + (fa.float() * g2.cuda()).sum() + la["feat"].float().mean()).backward()
    e_out = max(rel(po, pr), rel(fo, fr), rel(lo["feat"], lr["feat"]))
    a_out = max(rel(pa, pr), rel(fa, fr), rel(la["feat"], lr["feat"]))
    gr, ga = dict(ref.named_parameters()), dict(auto.named_parameters())
    g_ours, g_auto = {"x": rel(xo.grad, xr.grad)}, {"x": rel(xa.grad, xr.grad)}
    for n, p in ours.named_parameters():
        if n == "gpt_model.wpe.weight":
            assert torch.count_nonzero(p.grad[T:]) == 0      # only positions 0..T-1 are used
            g_ours[n], g_auto[n] = rel(p.grad[:T], gr[n].grad[:T]), rel(ga[n].grad[:T], gr[n].grad[:T])
        else:
            g_ours[n], g_auto[n] = rel(p.grad, gr[n].grad), rel(ga[n].grad, gr[n].grad)
    worst = max(g_ours, key=g_ours.get)
    record(f"avth/{C}-{Dh}-{nh}-{nl}-{B}-{T}", dict(out=e_out, out_autocast=a_out, grad=g_ours[worst], grad_name=worst,
                                                      grad_autocast=max(g_auto.values())))
    out_tol, grad_tol = AVTH_TOL[(C, Dh, T)]
    assert e_out <= out_tol and e_out <= 1.25 * a_out + 1e-4, (e_out, a_out)
    assert g_ours[worst] <= grad_tol and g_ours[worst] <= 1.25 * max(g_auto.values()) + 1e-4, (worst, g_ours[worst], max(g_auto.values()))


def test_avth_matches_reference_golden():
    """Golden vectors produced by the UNMODIFIED reference AVTh + HF GPT2Model (oracle/gen_golden.py)."""
    from avt_b200 import future_prediction as fp
    g = torch.load(os.path.join(GOLDEN, "avth_ref_small.pt"))
    m = fp.AVTh(g["in_features"], future_pred_loss={"_target_": "torch.nn.MSELoss"}, **g["cfg"])
    m.load_state_dict(g["state"])
    m.cuda().eval()
    x = g["x"].clone().cuda().requires_grad_(True)
    past, fut, losses, _ = m(x, (x.shape[0],))
    e_out = max(rel(past, g["past"]), rel(fut, g["future"]), rel(losses["feat"], g["feat"]))
    ((past * g["g_past"].cuda()).sum() + (fut * g["g_future"].cuda()).sum() + losses["feat"].mean()).backward()
    e_grad = max([rel(x.grad, g["dx"])] + [rel(p.grad, g["grads"][n]) for n, p in m.named_parameters()
                                            if n in g["grads"] and n != "gpt_model.wpe.weight"])
    record("golden/avth", dict(out=e_out, grad=e_grad))
    assert e_out <= GOLDEN_AVTH_TOL[0] and e_grad <= GOLDEN_AVTH_TOL[1], (e_out, e_grad)


def test_avth_rollout_matches_reference_golden():
    """Evaluation rollout vs golden vectors of the UNMODIFIED reference AVTh (HF GPT-2 with its KV cache)."""
    from avt_b200 import future_prediction as fp
    g = torch.load(os.path.join(GOLDEN, "avth_rollout_ref_small.pt"))
    m = fp.AVTh(g["in_features"], future_pred_loss={"_target_": "torch.nn.MSELoss"}, **g["cfg"])
    m.load_state_dict(g["state"])
    m.cuda().eval()
    x = g["x"].cuda()
    with torch.no_grad():
        past, fut, losses, _ = m(x, (x.shape[0],))
        _, fut2, _, _ = m(x, (x.shape[0], 2, g["in_features"]))
    assert fut.shape == g["future"].shape and fut2.shape == g["future_len2"].shape
    assert rel(past, g["past"]) <= OUT_TOL and rel(losses["feat"], g["feat"]) <= OUT_TOL
    assert rel(fut, g["future"]) <= 2 * OUT_TOL and rel(fut2, g["future_len2"]) <= 2 * OUT_TOL


def test_full_model_matches_reference_golden():
    """BaseModel-level parity through the glue: golden from the unmodified reference BaseModel/TIMMModel/AVTh."""
    from avt_b200.model import AVTModel
    g = torch.load(os.path.join(GOLDEN, "basemodel_ref_small.pt"))
    m = AVTModel("vit_test_patch16_32", 64, 32, head_kwargs=g["head"])
    m.load_state_dict(g["state"])
    m.cuda().eval()
    out, aux = m(g["video"].cuda(), target_shape=(g["video"].shape[0],))
    e_out = max([rel(out[k], g["outputs"][k]) for k in ("logits/action", "past_logits/action", "future", "past")] +
                [rel(aux["feat"], g["feat"])])
    loss = out["logits/action"].square().mean() + out["past_logits/action"].square().mean() + aux["feat"].mean()
    e_loss = abs(loss.item() - g["loss"].item()) / abs(g["loss"].item())
    loss.backward()
    e_grad = max(rel(p.grad, g["grads"][n]) for n, p in m.named_parameters() if n in g["grads"] and "wpe" not in n)
    record("golden/basemodel", dict(out=e_out, loss=e_loss, grad=e_grad))
    assert e_out <= GOLDEN_BASE_TOL[0] and e_loss <= GOLDEN_BASE_TOL[0] and e_grad <= GOLDEN_BASE_TOL[1], (e_out, e_loss, e_grad)


def test_avth_causality_at_full_size():
    """expts/01 head (768 -> 2048, 6 layers, 4 heads), B=8, T=10: perturbing frame t leaves every prediction made
    from frames < t bit-identical (causal mask + row-independent GEMMs)."""
    ours, _ = _avth_pair(768, 2048, 4, 6)
    ours.eval()
    x = torch.randn(8, 10, 768, device="cuda")
    x2 = x.clone()
    x2[:, 6:] += 0.5
    with torch.no_grad():
        p1, f1, _, _ = ours(x, (8,))
        p2, f2, _, _ = ours(x2, (8,))
    assert torch.equal(p1[:, :7], p2[:, :7])        # past[:, t] = decoded[:, t-1] depends on frames <= t-1
    assert not torch.equal(p1[:, 7:], p2[:, 7:]) and not torch.equal(f1, f2)


def test_avth_dropout_train_mode():
    from avt_b200 import future_prediction as fp
    ours, ref = _avth_pair(64, 128, 2, 2)
    ours.train()
    x = torch.randn(4, 10, 64, device="cuda", requires_grad=True)
    p1, f1, l1, _ = ours(x, (4,))
    (p1.sum() + f1.sum() + l1["feat"].mean()).backward()
    assert torch.isfinite(x.grad).all() and all(torch.isfinite(p.grad).all() for p in ours.parameters())
    p2, _, _, _ = ours(x, (4,))
    assert not torch.equal(p1, p2)                   # a fresh Philox offset every forward
    # all pdrop = 0 through the GPT2Config kwargs (models/future_prediction.py:69,89-93): train() == eval()
    nodrop, _ = _avth_pair(64, 128, 2, 2, embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    nodrop.train()
    a, _, _, _ = nodrop(x.detach(), (4,))
    nodrop.eval()
    b, _, _, _ = nodrop(x.detach(), (4,))
    assert torch.equal(a, b)
    # dropout is unbiased: the mean over many masks approaches the eval output
    ours.eval()
    with torch.no_grad():
        e = ours(x.detach(), (4,))[0]
        ours.train()
        acc = torch.zeros_like(e)
        for _ in range(64):
            acc += ours(x.detach(), (4,))[0]
    assert rel(acc / 64, e) < 0.25


@pytest.mark.parametrize("bf16_head_grads", [False, True])
def test_direct_grad_mode_and_fused_sgd_match_torch_sgd(bf16_head_grads):
    """FlatDataParallel/FlatSGD path (what bench.py runs) == autograd grads + torch.optim.SGD on the same model.
    bf16_head_grads: the AVT-h weight-gradient GEMMs store bf16 (single-GPU option; always on under data parallelism)."""
    from avt_b200.model import AVTModel
    from avt_b200.optim import FlatSGD
    from avt_b200.parallel import FlatDataParallel
    hk = dict(n_head=2, n_layer=2, inter_dim=64, n_positions=32, embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    torch.manual_seed(0)
    a = AVTModel("vit_test_patch16_32", 64, 32, dropout=0.0, head_kwargs=hk)
    b = AVTModel("vit_test_patch16_32", 64, 32, dropout=0.0, head_kwargs=hk)
    stress_init(a)
    b.load_state_dict(a.state_dict())
    a, b = a.cuda().train(), b.cuda().train()
    video = torch.randn(2, 4, 3, 1, 32, 32, device="cuda")

    def loss_of(m):
        out, aux = m(video, target_shape=(2,))
        return out["logits/action"].square().mean() + out["past_logits/action"].square().mean() + aux["feat"].mean()

    dp = FlatDataParallel(a, bf16_head_grads=bf16_head_grads)
    opt_a = None
    opt_b = torch.optim.SGD(b.parameters(), lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-3)
    for _ in range(3):
        la = loss_of(a)
        if opt_a is None:
            opt_a = FlatSGD([dp.vit, dp.head], dp.other, lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-3)
        opt_a.zero_grad()        # (also clears the fused torch-owned parameters: func/train.py:221 semantics)
        la.backward()
        dp.finish_backward()
        opt_a.step()
        lb = loss_of(b)
        opt_b.zero_grad()
        lb.backward()
        opt_b.step()
        assert abs(la.item() - lb.item()) <= (3e-3 if bf16_head_grads else 1e-4) * abs(lb.item()) + 1e-6
    pb = dict(b.named_parameters())
    for n, p in a.named_parameters():
        # bf16-stored AVT-h matrix gradients: every update carries a 2^-9 relative rounding (the data-parallel payload)
        tol = 2e-3 if bf16_head_grads else 1e-4
        assert rel(p, pb[n]) < tol, (n, rel(p, pb[n]))
    if bf16_head_grads:
        pk = dp.head._pack
        assert pk.gb is not None and pk.matrix_grads_bf16
        assert all(p.grad is None for n, p in dp.head.named_parameters() if p.dim() >= 2)


def test_state_dict_load_after_first_forward_updates_kernels():
    from avt_b200 import backbone
    m = backbone.create_model("vit_test_patch16_32").cuda()
    x = torch.randn(2, 3, 32, 32, device="cuda")
    with torch.no_grad():
        y0 = m(x)
        sd = {k: v + 0.05 * torch.randn_like(v) for k, v in m.state_dict().items()}
        m.load_state_dict(sd)          # in-place copy into the flat buffer views; bf16 shadow must be refreshed
        y1 = m(x)
    ref = o_vit.create_model("vit_test_patch16_32")
    ref.load_state_dict({k: v.cpu() for k, v in sd.items()})
    assert not torch.equal(y0, y1) and rel(y1, ref(x.cpu())) <= OUT_TOL


@pytest.mark.parametrize("C,Dh,nh,nl,B,T,olen,avg", [
    (64, 128, 2, 3, 3, 6, 3, -1),          # small: 3-step rollout, every past + future feature returned
    (768, 2048, 4, 6, 8, 10, 4, 1),        # expts/01 head, 4-step rollout at evaluation, last future feature only
])
def test_avth_eval_rollout_vs_oracle(C, Dh, nh, nl, B, T, olen, avg):
    """autoregressive rollout at evaluation (reference future_prediction.py:168-202, KV-cached GPT-2
    calls fed their own last hidden state). Ours runs it as growing causal passes; the oracle uses the KV cache."""
    from avt_b200 import future_prediction as fp
    torch.manual_seed(0)
    kw = dict(output_len=1, output_len_eval=olen, inter_dim=Dh, n_head=nh, n_layer=nl, return_past_too=True, avg_last_n=avg)
    ref = o_avth.AVTh(C, future_pred_loss="mse", **kw)
    stress_init(ref)
    ours = fp.AVTh(C, future_pred_loss={"_target_": "torch.nn.MSELoss"}, **kw)
    ours.load_state_dict(ref.state_dict())
    ours, ref = ours.cuda().eval(), ref.double().eval()
    x = torch.randn(B, T, C, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        pr, fr, lr, _ = ref(x.double(), (B,))
        po, fo, lo, _ = ours(x.cuda(), (B,))
        # target_shape with 3 dims selects the rollout length explicitly (reference :123-124)
        _, f2r, _, _ = ref(x.double(), (B, 2, C))
        _, f2o, _, _ = ours(x.cuda(), (B, 2, C))
    assert po.shape == pr.shape and fo.shape == fr.shape and lo["feat"].shape == lr["feat"].shape
    assert fo.shape == ((B, C) if avg > 0 else (B, T + olen, C))
    assert rel(po, pr) <= OUT_TOL and rel(fo, fr) <= 2 * OUT_TOL and rel(lo["feat"], lr["feat"]) <= OUT_TOL
    assert rel(f2o, f2r) <= 2 * OUT_TOL
    ours.train()
    with pytest.raises(NotImplementedError):
        ours(x.cuda(), (B, 2, C))
