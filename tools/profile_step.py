"""Two training steps of the bench workload (for ncu launch lists): python tools/profile_step.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from avt_b200.model import AVTModel, training_loss
from avt_b200.optim import FlatSGD
from avt_b200.parallel import FlatDataParallel

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(42)
dev = torch.device("cuda", 0)
model = AVTModel().to(dev).train()
dp = FlatDataParallel(model)
video, target, sub = (t.to(dev) for t in bench.synth_batch(torch, 8, 10, 0, dev))
opt = None
for i in range(steps):
    torch.cuda.nvtx.range_push(f"step{i}")
    out, aux = model(video, target_shape=(8,))
    loss = training_loss(out, aux, target, sub)
    if opt is None:
        opt = FlatSGD([dp.vit, dp.head], dp.other, lr=1e-4, momentum=0.9, nesterov=True, weight_decay=1e-6)
    for p in dp.other:
        p.grad = None
    loss.backward()
    dp.finish_backward()
    opt.step()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("loss", loss.item())
