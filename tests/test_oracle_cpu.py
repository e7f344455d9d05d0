"""CPU tests (run with -m "not gpu"): the oracle is pinned against
  (1) golden vectors produced by the UNMODIFIED reference (oracle/gen_golden.py -> tests/golden/*.pt),
  (2) torchvision's VisionTransformer (independent implementation of the same ViT math),
  (3) the installed HF transformers GPT2Model (the library the reference calls).
"""
import os

import pytest
import torch

from oracle import avth as o_avth
from oracle import base_model as o_base
from oracle import gpt2 as o_gpt2
from oracle import vit as o_vit

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_oracle_avth_matches_reference_golden():
    g = torch.load(os.path.join(GOLDEN, "avth_ref_small.pt"))
    m = o_avth.AVTh(g["in_features"], future_pred_loss="mse", **g["cfg"])
    m.load_state_dict(g["state"])
    m.eval()
    x = g["x"].clone().requires_grad_(True)
    past, fut, losses, _ = m(x, (x.shape[0],))
    assert rel(past, g["past"]) < 1e-5 and rel(fut, g["future"]) < 1e-5 and rel(losses["feat"], g["feat"]) < 1e-5
    ((past * g["g_past"]).sum() + (fut * g["g_future"]).sum() + losses["feat"].mean()).backward()
    assert rel(x.grad, g["dx"]) < 1e-5
    for n, p in m.named_parameters():
        if n in g["grads"]:
            assert rel(p.grad, g["grads"][n]) < 2e-5, n


def test_oracle_avth_rollout_matches_reference_golden():
    """Evaluation rollout (output_len_eval = 3 and a 3-d target_shape) of the oracle vs the UNMODIFIED reference AVTh with
    the installed HF GPT2Model and its KV cache (oracle/gen_golden.py: gen_avth_rollout)."""
    g = torch.load(os.path.join(GOLDEN, "avth_rollout_ref_small.pt"))
    m = o_avth.AVTh(g["in_features"], future_pred_loss="mse", **g["cfg"])
    m.load_state_dict(g["state"])
    m.eval()
    with torch.no_grad():
        past, fut, losses, _ = m(g["x"], (g["x"].shape[0],))
        _, fut2, _, _ = m(g["x"], (g["x"].shape[0], 2, g["in_features"]))
    assert fut.shape == g["future"].shape and fut2.shape == g["future_len2"].shape
    assert rel(past, g["past"]) < 1e-5 and rel(fut, g["future"]) < 1e-5 and rel(losses["feat"], g["feat"]) < 1e-5
    assert rel(fut2, g["future_len2"]) < 1e-5


def test_oracle_basemodel_matches_reference_golden():
    g = torch.load(os.path.join(GOLDEN, "basemodel_ref_small.pt"))
    m = o_base.BaseModel("vit_test_patch16_32", 64, 32, head_kwargs=g["head"])
    m.load_state_dict(g["state"])
    m.eval()
    out, aux = m(g["video"], target_shape=(g["video"].shape[0],))
    for k in ("logits/action", "past_logits/action", "future", "past"):
        assert rel(out[k], g["outputs"][k]) < 1e-5, k
    assert rel(aux["feat"], g["feat"]) < 1e-5
    loss = out["logits/action"].square().mean() + out["past_logits/action"].square().mean() + aux["feat"].mean()
    assert abs(loss.item() - g["loss"].item()) < 1e-5 * abs(g["loss"].item())
    loss.backward()
    for n, p in m.named_parameters():
        if n in g["grads"]:
            assert rel(p.grad, g["grads"][n]) < 5e-5, n


def test_oracle_vit_matches_torchvision():
    from torchvision.models.vision_transformer import VisionTransformer as TV
    torch.manual_seed(0)
    tv = TV(image_size=32, patch_size=16, num_layers=2, num_heads=2, hidden_dim=64, mlp_dim=256, num_classes=10).double()
    tv.heads = torch.nn.Identity()
    for p in tv.parameters():
        torch.nn.init.normal_(p, std=0.2)
    m = o_vit.create_model("vit_test_patch16_32").double()
    sd = tv.state_dict()
    mp = {"cls_token": "class_token", "pos_embed": "encoder.pos_embedding", "patch_embed.proj.weight": "conv_proj.weight",
          "patch_embed.proj.bias": "conv_proj.bias", "norm.weight": "encoder.ln.weight", "norm.bias": "encoder.ln.bias"}
    for i in range(2):
        e = f"encoder.layers.encoder_layer_{i}."
        b = f"blocks.{i}."
        mp.update({b + "norm1.weight": e + "ln_1.weight", b + "norm1.bias": e + "ln_1.bias",
                   b + "attn.qkv.weight": e + "self_attention.in_proj_weight", b + "attn.qkv.bias": e + "self_attention.in_proj_bias",
                   b + "attn.proj.weight": e + "self_attention.out_proj.weight", b + "attn.proj.bias": e + "self_attention.out_proj.bias",
                   b + "norm2.weight": e + "ln_2.weight", b + "norm2.bias": e + "ln_2.bias",
                   b + "mlp.fc1.weight": e + "mlp.0.weight", b + "mlp.fc1.bias": e + "mlp.0.bias",
                   b + "mlp.fc2.weight": e + "mlp.3.weight", b + "mlp.fc2.bias": e + "mlp.3.bias"})
    m.load_state_dict({k: sd[v] for k, v in mp.items()})
    x = torch.randn(3, 3, 32, 32, dtype=torch.float64)
    assert rel(m(x), tv(x)) < 1e-10


def test_oracle_gpt2_matches_hf():
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(0)
    cfg = transformers.GPT2Config(n_embd=64, n_layer=2, n_head=2, vocab_size=8, n_positions=32)
    hf = transformers.GPT2Model(cfg).eval()
    m = o_gpt2.GPT2Model(n_embd=64, n_layer=2, n_head=2, n_positions=32).eval()
    sd = {k: v for k, v in hf.state_dict().items() if not k.startswith("wte") and not k.endswith((".attn.bias", ".attn.masked_bias"))}
    m.load_state_dict(sd)
    x = torch.randn(2, 7, 64)
    pos = torch.arange(7)
    with torch.no_grad():
        ref = hf(inputs_embeds=x, position_ids=pos).last_hidden_state
        out, presents = m(x, None, pos)
        assert rel(out, ref) < 1e-5
        # KV-cache path == full causal recompute (reference rollout, future_prediction.py:168-202)
        nxt = out[:, -1:, :]
        step, _ = m(nxt, presents, torch.arange(7, 8))
        full, _ = m(torch.cat([x, nxt], 1), None, torch.arange(8))
        assert rel(step[:, -1], full[:, -1]) < 1e-5


def test_oracle_avth_rollout_shapes_and_causality():
    m = o_avth.AVTh(16, output_len=3, inter_dim=32, n_head=2, n_layer=1, n_positions=32, return_past_too=True,
                    avg_last_n=-1).eval()
    x = torch.randn(2, 5, 16)
    past, final, losses, _ = m(x, (2,))
    assert past.shape == (2, 5, 16) and final.shape == (2, 5 + 3, 16) and losses == {}
    x2 = x.clone()
    x2[:, 3:] += 1.0
    past2, _, _, _ = m(x2, (2,))
    assert torch.allclose(past[:, :4], past2[:, :4], atol=1e-6)   # position t only sees frames <= t-1 (shifted by 1)
    assert not torch.allclose(past[:, 4], past2[:, 4], atol=1e-4)


def test_oracle_multicrop_matches_reference():
    """Test-time augmentation (video.ndim == 7: crops folded by BaseModel.forward, models/base_model.py:239-273): the oracle's
    restatement equals the unmodified reference."""
    from oracle import base_model as ob
    from oracle import ref_host
    if not ref_host.available():
        pytest.skip("reference checkout not present")
    torch.manual_seed(0)
    hk = dict(n_head=2, n_layer=2, inter_dim=64, n_positions=32)
    ref = ref_host.build_reference_model(num_classes=32, model_type="vit_test_patch16_32", backbone_dim=64, head=hk).eval()
    o = ob.BaseModel("vit_test_patch16_32", 64, 32, head_kwargs=hk).eval()
    o.load_state_dict({k: v for k, v in ref.state_dict().items() if not k.endswith((".attn.bias", ".attn.masked_bias"))})
    v = torch.randn(2, 4, 3, 3, 1, 32, 32)
    with torch.no_grad():
        a, la = ref(v, target_shape=(2,))
        b, lb = o(v, target_shape=(2,))
    for k in b:
        assert (a[k] - b[k]).abs().max().item() < 1e-5, k
    assert (la["feat"] - lb["feat"]).abs().max().item() < 1e-5
