"""Fused SGD for the flat AVT-b / AVT-h parameter buffers (reference: torch.optim.SGD built at func/train.py:743-747
from conf/opt/optimizer/sgd.yaml; momentum 0.9, nesterov, expts/01:26-28). One kernel per flat buffer updates the fp32
master weights, the momentum buffer and the bf16 shadow the GEMMs read; everything else (classifier) stays on a stock
torch.optim.SGD with the same hyper-parameters."""
import torch

from . import ops


class FlatSGD:
    def __init__(self, flat_modules, other_params, lr, momentum=0.9, weight_decay=0.0, nesterov=True):
        self.mods = list(flat_modules)
        self.lr, self.momentum, self.wd, self.nesterov = lr, momentum, weight_decay, nesterov
        self.m = [None] * len(self.mods)
        other_params = list(other_params)
        # torch-owned parameters (classifier): big contiguous fp32 tensors go through the same fused kernel (one pass,
        # no bf16 shadow) instead of torch's foreach SGD (4-5 passes); whatever is left (odd-sized biases) stays on a stock
        # torch.optim.SGD with the same hyper-parameters, which also serves lr schedulers through `param_groups`.
        self.fused_other = [p for p in other_params if p.numel() % 4 == 0 and p.numel() >= 1024]
        self.fused_m = [None] * len(self.fused_other)
        rest = [p for p in other_params if not any(p is q for q in self.fused_other)]
        self.other = torch.optim.SGD(rest, lr=lr, momentum=momentum, weight_decay=weight_decay,
                                     nesterov=nesterov) if rest else None

    @property
    def param_groups(self):  # lr schedulers poke at this
        return self.other.param_groups if self.other is not None else [{"lr": self.lr}]

    def step(self):
        self.sync_lr()
        for i in range(len(self.mods)):
            self.step_flat(i)
        self.step_other()

    def sync_lr(self):
        if self.other is not None:
            self.lr = self.other.param_groups[0]["lr"]

    def step_flat(self, i):
        """Update flat module i (its gradients must be final, i.e. all-reduced). FlatDataParallel.finish_backward calls
        the pieces one by one so that the update of the AVT-h buffer (78 % of the bytes, all-reduced long ago) runs
        while the last gradient slices of the backbone are still on the wire."""
        pack = self.mods[i]._pack
        first = self.m[i] is None
        if first:
            self.m[i] = torch.empty_like(pack.w)
        ops.sgd_step(pack.w, pack.g, self.m[i], pack.b, self.lr, self.momentum, self.wd, self.nesterov, first)
        pack.shadow_is_current()

    def step_other(self):
        for i, p in enumerate(self.fused_other):
            if p.grad is None:
                continue
            first = self.fused_m[i] is None
            if first:
                self.fused_m[i] = torch.empty_like(p.data)
            assert p.data.is_contiguous() and p.grad.is_contiguous() and p.dtype == torch.float32
            ops.sgd_step(p.data, p.grad, self.fused_m[i], None, self.lr, self.momentum, self.wd, self.nesterov, first)
        if self.other is not None:
            self.other.step()

    def zero_grad(self, set_to_none=True):
        """optimizer.zero_grad() of the reference loop (func/train.py:221). The flat modules overwrite their gradient
        buffers every backward; the torch-owned parameters accumulate like any autograd leaf, so ALL of them are cleared
        here - also the big ones that step through the fused kernel and are not in the inner torch optimizer."""
        for p in self.fused_other:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()
        if self.other is not None:
            self.other.zero_grad(set_to_none=set_to_none)
