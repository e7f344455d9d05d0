"""Training steps of the bench workload, exactly as bench.py's `core` runs them (fused loss head, bf16 AVT-h gradients,
FlatSGD inside finish_backward), eager, for ncu launch lists: python tools/profile_step.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from avt_b200.model import AVTModel, past_targets
from avt_b200.optim import FlatSGD
from avt_b200.parallel import FlatDataParallel

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(42)
dev = torch.device("cuda", 0)
model = AVTModel("vit_base_patch16_224", 768, bench.NUM_CLASSES).to(dev).train()
dp = FlatDataParallel(model, bf16_head_grads=True)
video, target, sub = (t.to(dev) for t in bench.synth_batch(torch, 8, 10, 0, dev))
ptgt = past_targets(sub)
opt = None
for i in range(steps):
    torch.cuda.nvtx.range_push(f"step{i}")
    dp.begin_step()
    losses, acc = model.training_losses(video, target, ptgt)
    loss = sum(losses.values())
    if opt is None:
        opt = FlatSGD([dp.vit, dp.head], dp.other, lr=1e-4, momentum=0.9, nesterov=True, weight_decay=1e-6)
        opt.use_device_lr(dev)
        model.attach_loss_head_to(opt)
    opt.zero_grad()
    loss.backward()
    dp.finish_backward(opt)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("loss", loss.item())
