"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference: models.base_model.BaseModel, models.video_classification.TIMMModel,
models.future_prediction.AVTh + installed transformers GPT2Model) on CPU under oracle/ref_host.py stubs.

    python -m oracle.gen_golden          (authoring container only; the fixtures are committed)

The reference has no tests of its own, so these vectors are what pins the oracle and the CUDA
path to the reference's behaviour. TEST INFRASTRUCTURE.
"""
import os

import torch

from . import ref_host

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _stress(model, seed):
    """Non-trivial weights: the reference init (all nn.Linear ~ N(0, 0.01)) makes attention near-uniform and
    would hide attention bugs (SURVEY.md §7 'hard parts')."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() >= 2 and not any(k in n for k in ("pos_embed", "cls_token", "wpe")):
                conv1d = any(k in n for k in ("c_attn", "c_fc", "c_proj"))
                fan_in = p.shape[0] if conv1d else p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)
            elif "norm" in n or "ln_" in n:
                p.copy_((1.0 if n.endswith("weight") else 0.0) + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))


def gen_avth():
    torch.manual_seed(0)
    cfg = dict(output_len=1, inter_dim=64, n_head=2, n_layer=2, n_positions=32, return_past_too=True, avg_last_n=1,
               future_pred_loss={"_target_": "torch.nn.MSELoss"}, future_pred_loss_wt=1.0)
    m = ref_host.build_reference_avth(64, **cfg)
    _stress(m, 1)
    m.eval()  # dropout off: RNG streams cannot be matched across implementations
    B, T = 3, 6
    x = torch.randn(B, T, 64, generator=torch.Generator().manual_seed(2)).requires_grad_(True)
    past, fut, losses, _ = m(x, (B,))
    g1 = torch.randn(past.shape, generator=torch.Generator().manual_seed(3))
    g2 = torch.randn(fut.shape, generator=torch.Generator().manual_seed(4))
    ((past * g1).sum() + (fut * g2).sum() + losses["feat"].mean()).backward()
    sd = {k: v.clone() for k, v in m.state_dict().items() if not k.endswith((".attn.bias", ".attn.masked_bias"))}
    torch.save(dict(cfg={k: v for k, v in cfg.items() if k != "future_pred_loss"}, in_features=64, state=sd,
                    x=x.detach(), past=past.detach(), future=fut.detach(), feat=losses["feat"].detach(), g_past=g1,
                    g_future=g2, dx=x.grad.clone(),
                    grads={n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}),
               os.path.join(OUT, "avth_ref_small.pt"))
    print("avth:", sum(v.numel() for v in sd.values()), "params")


def gen_avth_rollout():
    """Evaluation-time autoregressive rollout (reference future_prediction.py:168-202, KV-cached HF GPT-2 calls): 3 steps,
    every past + predicted feature returned (avg_last_n = -1)."""
    torch.manual_seed(0)
    cfg = dict(output_len=1, output_len_eval=3, inter_dim=64, n_head=2, n_layer=2, n_positions=32, return_past_too=True,
               avg_last_n=-1, future_pred_loss={"_target_": "torch.nn.MSELoss"}, future_pred_loss_wt=1.0)
    m = ref_host.build_reference_avth(64, **cfg)
    _stress(m, 11)
    m.eval()
    B, T = 3, 6
    x = torch.randn(B, T, 64, generator=torch.Generator().manual_seed(12))
    with torch.no_grad():
        past, fut, losses, _ = m(x, (B,))
        _, fut2, _, _ = m(x, (B, 2, 64))          # 3-d target_shape selects the rollout length (:123-124)
    sd = {k: v.clone() for k, v in m.state_dict().items() if not k.endswith((".attn.bias", ".attn.masked_bias"))}
    torch.save(dict(cfg={k: v for k, v in cfg.items() if k != "future_pred_loss"}, in_features=64, state=sd, x=x, past=past,
                    future=fut, feat=losses["feat"], future_len2=fut2), os.path.join(OUT, "avth_rollout_ref_small.pt"))
    print("avth rollout:", tuple(fut.shape), tuple(fut2.shape))


def gen_basemodel():
    torch.manual_seed(0)
    head = dict(n_head=2, n_layer=2, inter_dim=64, n_positions=32)
    m = ref_host.build_reference_model(num_classes=32, model_type="vit_test_patch16_32", backbone_dim=64, head=head)
    _stress(m, 5)
    m.eval()
    B, T = 2, 4
    video = torch.randn(B, T, 3, 1, 32, 32, generator=torch.Generator().manual_seed(6))
    out, aux = m(video, target_shape=(B,))
    loss = out["logits/action"].square().mean() + out["past_logits/action"].square().mean() + aux["feat"].mean()
    loss.backward()
    sd = {k: v.clone() for k, v in m.state_dict().items() if not k.endswith((".attn.bias", ".attn.masked_bias"))}
    torch.save(dict(head=head, state=sd, video=video,
                    outputs={k: out[k].detach() for k in ("logits/action", "past_logits/action", "future", "past", "backbone")},
                    feat=aux["feat"].detach(), loss=loss.detach(),
                    grads={n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}),
               os.path.join(OUT, "basemodel_ref_small.pt"))
    print("basemodel:", sum(v.numel() for v in sd.values()), "params; loss", loss.item())


if __name__ == "__main__":
    assert ref_host.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    gen_avth()
    gen_avth_rollout()
    gen_basemodel()
