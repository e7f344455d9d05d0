"""Model-level parity beyond the per-module tests: the BASELINE cfg2 shape end to end, test-time augmentation (multi-crop),
and forward/backward orderings that exercise the activation-workspace leases. Oracle = CPU fp32 restatement (oracle/).

Every comparison records its measured error in gpurun_out/parity_errors.json (when that directory exists), so the
tolerances below can be kept at <= 1.5x what the path actually achieves."""
import json
import os

import pytest
import torch

from oracle import base_model as o_base
from test_modules_gpu import rel, stress_init

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name, value):
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    path = os.path.join(d, "parity_errors.json")
    try:
        data = json.load(open(path))
    except Exception:
        data = {}
    data[name] = value
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)


def _pair(model_type, dim, classes, head_kwargs, init, dropout=0.2):
    from avt_b200.model import AVTModel
    torch.manual_seed(0)
    ref = o_base.BaseModel(model_type, dim, classes, head_kwargs=head_kwargs, dropout=dropout)
    if init == "stress":
        stress_init(ref, 7)
    ours = AVTModel(model_type, dim, classes, head_kwargs=head_kwargs, dropout=dropout)
    ours.load_state_dict(ref.state_dict())
    return ours.cuda(), ref


def _loss(out, aux, target, sub):
    return o_base.training_loss(out, aux, target, sub)


@pytest.mark.parametrize("init", ["default", "stress"])
def test_cfg2_full_size_forward_backward_vs_oracle(init):
    """BASELINE configs[1] at full size: ViT-B/16 + AVT-h (expts/01), 8 clips x 10 frames x 224^2, the training loss of
    func/train.py:207-217 and its gradients, eval mode (dropout cannot be RNG-matched). `default` = the reference's init."""
    ours, ref = _pair("vit_base_patch16_224", 768, 3806, None, init)
    ours.eval()
    ref.eval()
    g = torch.Generator().manual_seed(11)
    B, T = 8, 10
    video = torch.rand(B, T, 3, 1, 224, 224, generator=g) * 2 - 1
    target = torch.randint(0, 3806, (B,), generator=g)
    sub = torch.randint(0, 3806, (B, T, 1), generator=g)
    sub[torch.rand(B, T, 1, generator=g) < 0.1] = -1
    out_r, aux_r = ref(video, target_shape=(B,))
    loss_r = _loss(out_r, aux_r, target, sub)
    loss_r.backward()
    out_o, aux_o = ours(video.cuda(), target_shape=(B,))
    loss_o = _loss(out_o, aux_o, target.cuda(), sub.cuda())
    loss_o.backward()
    errs = {k: rel(out_o[k], out_r[k]) for k in ("future", "past", "logits/action", "past_logits/action")}
    errs["feat"] = rel(aux_o["feat"], aux_r["feat"])
    errs["loss"] = abs(loss_o.item() - loss_r.item()) / abs(loss_r.item())
    gr = dict(ref.named_parameters())
    gerr = {n: rel(p.grad, gr[n].grad) for n, p in ours.named_parameters() if "wpe" not in n}
    worst = max(gerr, key=gerr.get)
    errs["worst_grad"] = gerr[worst]
    record(f"cfg2_full_{init}", dict(errs, worst_grad_name=worst))
    tol_out, tol_grad = 1.4e-2, 2.1e-2        # 1.5x measured (outputs 9.1e-3, worst gradient 1.4e-2: patch_embed.proj.weight)
    assert all(v <= tol_out for k, v in errs.items() if k != "worst_grad"), errs
    assert errs["worst_grad"] <= tol_grad, (worst, errs["worst_grad"])


def test_multicrop_tta_eval_and_train_vs_oracle():
    """video.ndim == 7 (B, #clips, #crops, C, T', H, W): 3 crops averaged (models/base_model.py:239-273). In train mode the
    three forwards of the same shape each hold their own activation workspace until the single backward."""
    hk = dict(n_head=2, n_layer=2, inter_dim=64, n_positions=32, embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    ours, ref = _pair("vit_test_patch16_32", 64, 32, hk, "stress", dropout=0.0)
    g = torch.Generator().manual_seed(5)
    video = torch.randn(2, 4, 3, 3, 1, 32, 32, generator=g)
    ours.eval()
    ref.double().eval()
    with torch.no_grad():
        oo, lo = ours(video.cuda(), target_shape=(2,))
        orr, lr = ref(video.double(), target_shape=(2,))
    e_eval = max(rel(oo[k], orr[k]) for k in orr)
    ours.train()
    ref.train()
    oo, lo = ours(video.cuda(), target_shape=(2,))
    (oo["logits/action"].square().mean() + oo["past_logits/action"].square().mean() + lo["feat"].mean()).backward()
    orr, lr = ref(video.double(), target_shape=(2,))
    (orr["logits/action"].square().mean() + orr["past_logits/action"].square().mean() + lr["feat"].mean()).backward()
    gr = dict(ref.named_parameters())
    e_grad = max(rel(p.grad, gr[n].grad) for n, p in ours.named_parameters() if "wpe" not in n)
    record("multicrop", dict(eval=e_eval, grad=e_grad))
    assert e_eval <= 1.35e-2 and e_grad <= 2.35e-2, (e_eval, e_grad)      # 1.5x measured (8.9e-3, 1.56e-2)


def test_forward_forward_backward_backward_keeps_each_forwards_activations():
    """Two forwards of the same shape before their backwards (gradient accumulation, the advisor's round-1 finding): the
    gradients equal those of the two losses computed one after the other."""
    from avt_b200 import backbone
    torch.manual_seed(0)
    m = backbone.create_model("vit_test_patch16_64").cuda()
    stress_init(m, 3)
    x1 = torch.randn(3, 3, 64, 64, device="cuda")
    x2 = torch.randn(3, 3, 64, 64, device="cuda")
    g1, g2 = torch.randn(3, 128, device="cuda"), torch.randn(3, 128, device="cuda")
    m(x1).backward(g1)
    ga = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad()
    m(x2).backward(g2)
    gb = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad()
    y1, y2 = m(x1), m(x2)          # second forward must not overwrite what the first backward reads
    y1.backward(g1)
    y2.backward(g2)
    for n, p in m.named_parameters():
        assert torch.allclose(p.grad, ga[n] + gb[n], rtol=1e-5, atol=1e-6), n
    # an eval forward between a train forward and its backward is harmless as well
    m.zero_grad()
    y1 = m(x1)
    with torch.no_grad():
        m(x2)
    y1.backward(g1)
    for n, p in m.named_parameters():
        assert torch.equal(p.grad, ga[n]), n
    # retain_graph after another forward of the same shape took the workspace over: loud error, not wrong gradients
    y1 = m(x1)
    y1.backward(g1, retain_graph=True)
    m(x2).backward(g2)
    with pytest.raises(RuntimeError, match="overwritten"):
        y1.backward(g1)


@pytest.mark.parametrize("B,T,K,C,p", [(8, 10, 768, 3806, 0.0), (3, 4, 64, 37, 0.0), (8, 15, 768, 3806, 0.0), (8, 10, 768, 3806, 0.2)])
def test_fused_classifier_loss_vs_torch(B, T, K, C, p):
    """avt_b200.loss_head (dropout -> classifier GEMM -> softmax-CE fwd+bwd -> gradient GEMMs) against the reference's
    arithmetic in fp64: F.linear + CrossEntropyLoss(ignore_index=-1, reduction='none') means + top-k accuracy
    (models/base_model.py:203-216, func/train_eval_ops.py:57-85, common/utils.py:17-44)."""
    import torch.nn.functional as F
    from avt_b200.loss_head import FusedClassifierLoss
    from avt_b200.model import accuracy
    g = torch.Generator().manual_seed(B * 100 + T)
    lin = torch.nn.Linear(K, C)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(C, K, generator=g) * 0.05)
        lin.bias.copy_(torch.randn(C, generator=g) * 0.1)
    past = torch.randn(B, T, K, generator=g)
    fut = torch.randn(B, K, generator=g)
    tgt = torch.randint(0, C, (B,), generator=g)
    ptgt = torch.randint(0, C, (B, T), generator=g)
    ptgt[torch.rand(B, T, generator=g) < 0.15] = -1
    tgt[0] = -1
    lin_c = torch.nn.Linear(K, C).cuda()
    lin_c.load_state_dict(lin.state_dict())
    head = FusedClassifierLoss(lin_c)
    pc, fc = past.cuda().requires_grad_(True), fut.cuda().requires_grad_(True)
    lf, lp, a1, a5 = head(pc, fc, ptgt.cuda(), tgt.cuda(), p)
    (lf + lp).backward()
    if p > 0.0:      # masks cannot be matched: finite, a fresh mask per call, and unbiased on average
        assert torch.isfinite(lf) and torch.isfinite(pc.grad).all() and torch.isfinite(lin_c.weight.grad).all()
        frac = (pc.grad == 0).float().mean().item()
        assert 0.1 < frac < 0.45                     # ~p of the input gradient is masked (+ rows whose label is ignored)
        lf2, _, _, _ = head(pc.detach(), fc.detach(), ptgt.cuda(), tgt.cuda(), p)
        assert lf2.item() != lf.item()
        return
    lin64 = lin.double()
    p64, f64 = past.double().requires_grad_(True), fut.double().requires_grad_(True)
    lo_f, lo_p = lin64(f64), lin64(p64)
    rf = F.cross_entropy(lo_f, tgt, ignore_index=-1, reduction="none").mean()
    rp = F.cross_entropy(lo_p.flatten(0, 1), ptgt.flatten(), ignore_index=-1, reduction="none").mean()
    (rf + rp).backward()
    r1, r5 = accuracy(lo_f, tgt, topk=(1, min(5, C)))
    errs = dict(loss_f=abs(lf.item() - rf.item()) / abs(rf.item()), loss_p=abs(lp.item() - rp.item()) / abs(rp.item()),
                dW=rel(lin_c.weight.grad, lin64.weight.grad), db=rel(lin_c.bias.grad, lin64.bias.grad),
                dpast=rel(pc.grad, p64.grad), dfut=rel(fc.grad, f64.grad))
    record(f"loss_head/{B}-{T}-{K}-{C}", errs)
    # bf16 operands for the classifier GEMMs (like every other GEMM of the path): logits to ~3e-3, gradients to ~5e-3
    assert errs["loss_f"] < 4e-4 and errs["loss_p"] < 4e-4, errs                    # 1.5x measured (2.3e-4)
    assert max(errs["dW"], errs["db"], errs["dpast"], errs["dfut"]) < 6e-3, errs   # 1.5x measured (3.9e-3)
    assert abs(a1.item() - r1.item()) < 1e-3 and abs(a5.item() - r5.item()) < 1e-3
    # ignored rows contribute nothing
    assert torch.count_nonzero(fc.grad[0]) == 0


def test_training_losses_fused_head_equals_forward_plus_loss():
    """AVTModel.training_losses (fused classifier / cross-entropy head) == forward() + training_loss() + accuracy() of the
    same model, and both match the oracle; gradients included (eval mode: no dropout)."""
    from avt_b200.model import accuracy, past_targets, training_loss
    hk = dict(n_head=2, n_layer=2, inter_dim=64, n_positions=32)
    ours, ref = _pair("vit_test_patch16_32", 64, 37, hk, "stress")
    ours.eval()
    ref.double().eval()
    g = torch.Generator().manual_seed(9)
    B, T = 4, 6
    video = torch.randn(B, T, 3, 1, 32, 32, generator=g)
    target = torch.randint(0, 37, (B,), generator=g)
    sub = torch.randint(0, 37, (B, T, 1), generator=g)
    sub[torch.rand(B, T, 1, generator=g) < 0.2] = -1
    losses, accs = ours.training_losses(video.cuda(), target.cuda(), past_targets(sub.cuda()))
    fused = sum(losses.values())
    fused.backward()
    g_fused = {n: p.grad.clone() for n, p in ours.named_parameters()}
    ours.zero_grad()
    out, aux = ours(video.cuda(), target_shape=(B,))
    plain = training_loss(out, aux, target.cuda(), sub.cuda())
    plain.backward()
    a1, a5 = accuracy(out["logits/action"], target.cuda(), topk=(1, 5))
    assert abs(fused.item() - plain.item()) <= 2e-3 * abs(plain.item())
    assert abs(accs["acc1/action"].item() - a1.item()) < 1e-3 and abs(accs["acc5/action"].item() - a5.item()) < 1e-3
    out_r, aux_r = ref(video.double(), target_shape=(B,))
    loss_r = o_base.training_loss(out_r, aux_r, target, sub)
    loss_r.backward()
    gr = dict(ref.named_parameters())
    e_loss = abs(fused.item() - loss_r.item()) / abs(loss_r.item())
    e_grad = max(rel(g_fused[n], gr[n].grad) for n in g_fused if "wpe" not in n)
    e_grad_plain = max(rel(p.grad, gr[n].grad) for n, p in ours.named_parameters() if "wpe" not in n)
    record("training_losses_fused", dict(loss=e_loss, grad=e_grad, grad_unfused=e_grad_plain))
    assert e_loss <= 5e-3 and e_grad <= max(2.9e-2, 1.3 * e_grad_plain), (e_loss, e_grad, e_grad_plain)


def test_fp32_mode_matches_oracle_to_1e5():
    """BASELINE.json north star: 'within 1e-3 rel (bf16) / 1e-5 (fp32)'. precision='fp32' runs the same restated operators in
    fp32 on the CUDA cores (csrc/fp32_path.cu): backbone, head and the whole model against the fp64 oracle at 1e-5, next to
    the bf16 product path's few 1e-3 on the same inputs."""
    from avt_b200 import backbone
    from oracle import vit as o_vit
    res = {}
    for model_type, F in (("vit_test_patch16_64", 3), ("vit_base_patch16_224", 2)):
        torch.manual_seed(0)
        ref = o_vit.create_model(model_type)
        stress_init(ref, 5)
        ours = backbone.create_model(model_type)
        ours.load_state_dict(ref.state_dict())
        ours.cuda().eval()
        img = o_vit.CONFIGS[model_type][0]
        x = torch.randn(F, 3, img, img, generator=torch.Generator().manual_seed(1))
        with torch.no_grad():
            yr = ref.double()(x.double())
            y16 = ours(x.cuda())
            ours.precision = "fp32"
            y32 = ours(x.cuda())
        res[model_type] = dict(fp32=rel(y32, yr), bf16=rel(y16, yr))
        assert res[model_type]["fp32"] <= 1e-5, res
        with pytest.raises(NotImplementedError):
            ours(x.cuda())            # grad mode: the fp32 mode is inference-only
    # head (expts/01 geometry) and the whole model through the glue
    hk = dict(n_head=2, n_layer=2, inter_dim=64, n_positions=32)
    ours, ref = _pair("vit_test_patch16_32", 64, 37, hk, "stress")
    ours.eval()
    ref.double().eval()
    video = torch.randn(2, 4, 3, 1, 32, 32, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        out_r, aux_r = ref(video.double(), target_shape=(2,))
        ours.set_precision("fp32")
        out_o, aux_o = ours(video.cuda(), target_shape=(2,))
    res["model"] = {k: rel(out_o[k], out_r[k]) for k in out_r}
    res["model"]["feat"] = rel(aux_o["feat"], aux_r["feat"])
    assert max(res["model"].values()) <= 1e-5, res
    from avt_b200 import future_prediction as fp
    from oracle import avth as o_avth
    torch.manual_seed(0)
    kw = dict(output_len=1, inter_dim=2048, n_head=4, n_layer=6, return_past_too=True, avg_last_n=1)
    href = o_avth.AVTh(768, future_pred_loss="mse", **kw)
    stress_init(href, 3)
    hours = fp.AVTh(768, future_pred_loss={"_target_": "torch.nn.MSELoss"}, **kw)
    hours.load_state_dict(href.state_dict())
    hours.cuda().eval()
    hours.precision = "fp32"
    xh = torch.randn(8, 10, 768, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        pr, fr, lr, _ = href.double().eval()(xh.double(), (8,))
        po, fo, lo, _ = hours(xh.cuda(), (8,))
    res["avth_expts01"] = dict(past=rel(po, pr), future=rel(fo, fr), feat=rel(lo["feat"], lr["feat"]))
    record("fp32_mode", res)
    assert max(res["avth_expts01"].values()) <= 1e-5, res


def test_fp32_mode_matches_the_unmodified_reference_golden_to_1e5():
    """The golden fixtures were produced by the UNMODIFIED reference (BaseModel + TIMMModel + AVTh + HF GPT2Model, fp32 CPU,
    oracle/gen_golden.py): the fp32 mode reproduces its outputs to 1e-5 (the reference's own fp32 rounding is ~1e-6)."""
    import os
    from avt_b200 import future_prediction as fp
    from avt_b200.model import AVTModel
    golden = os.path.join(ROOT, "tests", "golden")
    g = torch.load(os.path.join(golden, "basemodel_ref_small.pt"))
    m = AVTModel("vit_test_patch16_32", 64, 32, head_kwargs=g["head"])
    m.load_state_dict(g["state"])
    m.cuda().eval().set_precision("fp32")
    with torch.no_grad():
        out, aux = m(g["video"].cuda(), target_shape=(g["video"].shape[0],))
    errs = {k: rel(out[k], g["outputs"][k]) for k in ("logits/action", "past_logits/action", "future", "past")}
    errs["feat"] = rel(aux["feat"], g["feat"])
    g2 = torch.load(os.path.join(golden, "avth_ref_small.pt"))
    h = fp.AVTh(g2["in_features"], future_pred_loss={"_target_": "torch.nn.MSELoss"}, **g2["cfg"])
    h.load_state_dict(g2["state"])
    h.cuda().eval()
    h.precision = "fp32"
    with torch.no_grad():
        past, fut, losses, _ = h(g2["x"].cuda(), (g2["x"].shape[0],))
    errs.update(avth_past=rel(past, g2["past"]), avth_future=rel(fut, g2["future"]), avth_feat=rel(losses["feat"], g2["feat"]))
    record("fp32_mode_golden", errs)
    assert max(errs.values()) <= 1e-5, errs
