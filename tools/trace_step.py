"""In-stream (warm, unserialised) timing of every C-ABI call of one training step, plus the host enqueue time.

    python tools/trace_step.py [batch] [frames] [model]

Every avt_* call of step 3 is bracketed by CUDA events on the launching stream (the events add ~2 us of gaps, so the
sum is a little above the untraced step); torch's own kernels show up as the residual "other". Not a bench number.
"""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from avt_b200 import _lib
from avt_b200.model import AVTModel, training_loss
from avt_b200.optim import FlatSGD
from avt_b200.parallel import FlatDataParallel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10
model_type = sys.argv[3] if len(sys.argv) > 3 else "vit_base_patch16_224"
torch.manual_seed(42)
dev = torch.device("cuda", 0)
model = AVTModel(model_type, 1024 if "large" in model_type else 768, bench.NUM_CLASSES).to(dev).train()
dp = FlatDataParallel(model, bf16_head_grads=True)
video, target, sub = (t.to(dev) for t in bench.synth_batch(torch, B, T, 0, dev))
state = {"opt": None}


def step():
    out, aux = model(video, target_shape=(B,))
    loss = training_loss(out, aux, target, sub)
    if state["opt"] is None:
        state["opt"] = FlatSGD([dp.vit, dp.head], dp.other, lr=1e-4, momentum=0.9, nesterov=True, weight_decay=1e-6)
    for p in dp.other:
        p.grad = None
    loss.backward()
    dp.finish_backward()
    state["opt"].step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()

# host enqueue time vs device time of an untraced step
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
loss = step()
e1.record()
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"untraced step: device {e0.elapsed_time(e1):.2f} ms, host enqueue {t_enq * 1e3:.2f} ms")

events = []
orig_call = _lib.call


def traced(name, *args):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig_call(name, *args)
    b.record()
    key = name
    if name.startswith("avt_gemm_bf16"):
        M, N, K = (int(getattr(v, "value", v)) for v in args[6:9])
        key = f"gemm M{M} N{N} K{K} a{args[2]}b{args[5]} sk{args[10]}"
    elif name.startswith("avt_layernorm"):
        key = f"{name} rows{int(getattr(args[9] if name.endswith('fwd') else args[8], 'value', 0))}"
    elif name == "avt_colsum_bf16":
        key = f"colsum {args[1]}x{args[2]}"
    events.append((key, a, b))


_lib.call = traced
import avt_b200.ops as _ops   # ops holds a reference to the module, not the function: patching _lib.call is enough
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
step()
s1.record()
torch.cuda.synchronize()
_lib.call = orig_call
tot = collections.OrderedDict()
for key, a, b in events:
    t = a.elapsed_time(b) * 1e3
    n, s = tot.get(key, (0, 0.0))
    tot[key] = (n + 1, s + t)
total = s0.elapsed_time(s1) * 1e3
ours = sum(s for _, s in tot.values())
print(f"traced step {total:.0f} us; avt_* calls {ours:.0f} us; other (torch ops, gaps) {total - ours:.0f} us")
for key, (n, s) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{s:9.1f} us  {n:4d} x {s / n:8.1f}  {key}")
