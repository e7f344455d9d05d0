"""GPU bring-up: module-level parity of the drop-in backbone / head against the CPU oracle + step timing.
Prints relative errors; a manual checker (lives under tests/ because it uses the oracle).   python tests/bringup_model.py [quick|full|time]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avt_b200 import backbone, future_prediction
from oracle import avth as o_avth
from oracle import vit as o_vit


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def stress_init(m, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() >= 2 and "pos_embed" not in n and "cls_token" not in n and "wpe" not in n:
                fan_in = p.shape[0] if ("c_attn" in n or "c_fc" in n or "c_proj" in n) else p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)
            elif "norm" in n or "ln_" in n:
                p.copy_((1.0 if n.endswith("weight") else 0.0) + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g) if p.dim() < 2 else 0.02 * torch.randn(p.shape, generator=g))


def vit_parity(model_type, F, dtype=torch.float64, init="stress"):
    torch.manual_seed(0)
    ref = o_vit.create_model(model_type)
    if init == "stress":
        stress_init(ref)
    ours = backbone.create_model(model_type)
    ours.load_state_dict(ref.state_dict())
    ours.cuda()
    ref = ref.to(dtype)
    img = o_vit.CONFIGS[model_type][0]
    x = torch.randn(F, 3, img, img)
    t0 = time.time()
    yr = ref(x.to(dtype))
    gy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(1))
    yr.backward(gy.to(dtype))
    t_ref = time.time() - t0
    yo = ours(x.cuda())
    yo.backward(gy.cuda())
    torch.cuda.synchronize()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ya = ref.float().cuda()(x.cuda())
    ref = ref.cpu().to(dtype)
    print(f"[vit {model_type} F={F} init={init}] fwd rel {rel(yo, yr):.3e}  torch-autocast-bf16 rel {rel(ya, yr):.3e}"
          f"  (oracle {t_ref:.1f}s)")
    gr = dict(ref.named_parameters())
    worst = []
    for n, p in ours.named_parameters():
        worst.append((rel(p.grad, gr[n].grad), n))
    worst.sort(reverse=True)
    print("   worst grads:", [(f"{e:.2e}", n) for e, n in worst[:6]])
    print("   median grad rel:", f"{sorted(e for e, _ in worst)[len(worst)//2]:.2e}")
    return ours


def avth_parity(C, Dh, nh, nl, B, T, dtype=torch.float64):
    torch.manual_seed(0)
    kw = dict(output_len=1, inter_dim=Dh, n_head=nh, n_layer=nl, return_past_too=True, avg_last_n=1)
    ref = o_avth.AVTh(C, future_pred_loss="mse", **kw)
    stress_init(ref)
    ours = future_prediction.AVTh(C, future_pred_loss={"_target_": "torch.nn.MSELoss"}, future_pred_loss_wt=1.0, **kw)
    ours.load_state_dict(ref.state_dict())
    ours.cuda()
    ref = ref.to(dtype).eval()
    ours.eval()
    x = torch.randn(B, T, C)
    xr = x.detach().clone().to(dtype).requires_grad_(True)
    pr, fr, lr, _ = ref(xr, (B,))
    g1, g2 = torch.randn(pr.shape), torch.randn(fr.shape)
    (pr * g1.to(dtype)).sum().add((fr * g2.to(dtype)).sum()).add(lr["feat"].mean()).backward()
    xo = x.detach().clone().cuda().requires_grad_(True)
    po, fo, lo, _ = ours(xo, (B,))
    (po * g1.cuda()).sum().add((fo * g2.cuda()).sum()).add(lo["feat"].mean()).backward()
    torch.cuda.synchronize()
    print(f"[avth C{C} Dh{Dh} h{nh} L{nl} B{B} T{T}] past {rel(po, pr):.3e} future {rel(fo, fr):.3e} "
          f"feat {rel(lo['feat'], lr['feat']):.3e} dfeats {rel(xo.grad, xr.grad):.3e}")
    gr = dict(ref.named_parameters())
    worst = sorted(((rel(p.grad, gr[n].grad), n) for n, p in ours.named_parameters()), reverse=True)
    print("   worst grads:", [(f"{e:.2e}", n) for e, n in worst[:6]])
    print("   median grad rel:", f"{worst[len(worst)//2][0]:.2e}")


def time_step(B=8, T=10, steps=5):
    torch.manual_seed(0)
    bb = backbone.TIMMModel(1, "vit_base_patch16_224").cuda()
    head = future_prediction.AVTh(768, output_len=1, inter_dim=2048, n_head=4, n_layer=6, return_past_too=True,
                                  avg_last_n=1, future_pred_loss={"_target_": "torch.nn.MSELoss"}).cuda()
    bb.model.direct_grads = True
    head.direct_grads = True
    video = torch.randn(B * T, 3, 1, 224, 224, device="cuda")
    for it in range(steps + 2):
        if it == 2:
            torch.cuda.synchronize()
            t0 = time.time()
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        f = bb(video)                                   # (B*T, 768, 1, 1, 1)
        feats = f.mean([-1, -2]).permute(0, 2, 1).reshape(B, T, 768)
        past, fut, losses, _ = head(feats, (B,))
        loss = past.square().mean() + fut.square().mean() + losses["feat"].mean()
        loss.backward()
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"[time] B={B} T={T}: {ms:.2f} ms/step (wall {(time.time()-t0)/steps*1e3:.2f}) -> {B/ms*1e3:.1f} clips/s; "
          f"mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    vit_parity("vit_test_patch16_32", 3)
    vit_parity("vit_test_patch16_64", 2)
    vit_parity("vit_test_patch16_64", 2, init="default")
    avth_parity(64, 32, 2, 2, 2, 5)
    avth_parity(64, 128, 2, 3, 3, 10)
    if mode in ("full", "time"):
        avth_parity(768, 2048, 4, 6, 2, 10, dtype=torch.float32)
        vit_parity("vit_base_patch16_224", 2, dtype=torch.float32)
        time_step()
