"""Device time of the parts of one training step, each captured in its own CUDA graph and replayed (no host gaps):
ViT backbone fwd+bwd, AVT-h head fwd+bwd, loss head (classifier + CE), fused SGD.   python tools/time_parts.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from avt_b200.model import AVTModel, past_targets, training_loss
from avt_b200.optim import FlatSGD
from avt_b200.parallel import FlatDataParallel

dev = torch.device("cuda", 0)
torch.manual_seed(42)
B, T = 8, 10
model = AVTModel().to(dev).train()
dp = FlatDataParallel(model, bf16_head_grads=True)     # as bench.py runs it
video, target, sub = (t.to(dev) for t in bench.synth_batch(torch, B, T, 0, dev))
ptgt = past_targets(sub)


def graph_time(fn, iters=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


frames = video.flatten(0, 1).transpose(1, 2).flatten(0, 1)   # what TIMMModel feeds the ViT: (80, 3, 224, 224)
vit = model.backbone.model
dfeat = torch.randn(B * T, 768, device=dev)


def vit_step():
    f = vit(frames)
    f.backward(dfeat)


head = model.future_predictor
feats = torch.randn(B, T, 768, device=dev, requires_grad=True)


def head_step():
    past, fut, losses, _ = head(feats, (B,))
    (past.sum() * 1e-3 + fut.sum() * 1e-3 + losses["feat"].mean()).backward()


pf = torch.randn(B, T, 768, device=dev, requires_grad=True)
ff = torch.randn(B, 768, device=dev, requires_grad=True)


def loss_step():
    out = {"past_logits/action": model.classifiers["action"](model.dropout(pf)),
           "logits/action": model.classifiers["action"](model.dropout(ff))}
    training_loss(out, {}, target, past_tgt=ptgt).backward()


vit_step(); head_step(); loss_step()
opt = FlatSGD([dp.vit, dp.head], dp.other, lr=1e-4, momentum=0.9, nesterov=True, weight_decay=1e-6)
t_vit = graph_time(vit_step)
t_head = graph_time(head_step)
t_loss = graph_time(loss_step)
t_sgd = graph_time(lambda: opt.step())
print(f"ViT fwd+bwd {t_vit:.3f} ms | AVT-h fwd+bwd {t_head:.3f} ms | classifier+loss fwd+bwd {t_loss:.3f} ms | SGD {t_sgd:.3f} ms | "
      f"sum {t_vit + t_head + t_loss + t_sgd:.3f} ms")
