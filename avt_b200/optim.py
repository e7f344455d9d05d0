"""Fused SGD for the flat AVT-b / AVT-h parameter buffers (reference: torch.optim.SGD built at func/train.py:743-747
from conf/opt/optimizer/sgd.yaml; momentum 0.9, nesterov, expts/01:26-28). One kernel per flat buffer updates the fp32
master weights, the momentum buffer and the bf16 shadow the GEMMs read; the big torch-owned tensors (classifier weight)
go through the same kernel, whatever is left stays on a stock torch.optim.SGD with the same hyper-parameters.

Parameter groups as the reference builds them (func/train.py:696-742): within a module list every parameter whose name
ends in 'bias' (or contains '.bn') decays with `weight_decay * bias_bn_wd_scale`, the rest with `weight_decay`. The flat
buffers keep their biases in front (engine.ParamPack.bias_end), so the two groups are the two regions of one kernel
launch. One lr / weight decay per FlatSGD (expts/01: `opt.lr_wd=[[__all__, 1e-4, 1e-6]]`); parameters with
requires_grad=False (lr 0 groups, freeze_encoder_decoder) are not representable in a flat update and are rejected.

Data parallel (avt_b200.parallel.FlatDataParallel): gradients arrive in bf16; `step_flat_shard` updates one rank's
contiguous 1/N of a buffer (momentum exists only for that shard)."""
import torch

from . import ops


class FlatSGD:
    def __init__(self, flat_modules, other_params, lr, momentum=0.9, weight_decay=0.0, nesterov=True, bias_bn_wd_scale=1.0):
        self.mods = list(flat_modules)
        self.lr, self.momentum, self.wd, self.nesterov = lr, momentum, weight_decay, nesterov
        self.wd_bias = weight_decay * bias_bn_wd_scale
        self.m = [None] * len(self.mods)
        self.lr_dev = None       # fp32 device scalar: set_lr() writes it, the kernels read it (CUDA-graph-safe schedules)
        for mod in self.mods:
            frozen = [n for n, p in mod.named_parameters() if not p.requires_grad]
            if frozen:
                raise NotImplementedError(f"FlatSGD updates whole flat buffers; frozen parameters {frozen[:3]}... need a stock "
                                          "optimizer over flat_parameter_groups() / named parameters")
        other_params = list(other_params)
        # torch-owned parameters (classifier): big contiguous fp32 tensors go through the same fused kernel (one pass,
        # no bf16 shadow) instead of torch's foreach SGD (4-5 passes); whatever is left (odd-sized biases) stays on a stock
        # torch.optim.SGD with the same hyper-parameters, which also serves lr schedulers through `param_groups`.
        self.fused_other = [p for p in other_params if p.numel() % 4 == 0 and p.numel() >= 1024]
        self.fused_m = [None] * len(self.fused_other)
        self.shadows = {}      # id(param) -> (bf16 flat view of the same length, callback after the update)
        rest = [p for p in other_params if not any(p is q for q in self.fused_other)]
        groups = [dict(params=[p for p in rest if p.dim() >= 2], weight_decay=weight_decay),
                  dict(params=[p for p in rest if p.dim() < 2], weight_decay=self.wd_bias)]       # 1-D = biases
        groups = [g for g in groups if g["params"]]
        self.other = torch.optim.SGD(groups, lr=lr, momentum=momentum, nesterov=nesterov) if groups else None

    def attach_shadow(self, param, shadow_flat, on_update):
        """A bf16 copy of a fused torch-owned parameter that the update kernel keeps current (same element order)."""
        assert shadow_flat.dtype == torch.bfloat16 and shadow_flat.numel() == param.numel() and shadow_flat.is_contiguous()
        self.shadows[id(param)] = (shadow_flat, on_update)

    @property
    def param_groups(self):  # lr schedulers poke at this
        return self.other.param_groups if self.other is not None else [{"lr": self.lr}]

    def set_lr(self, lr):
        """Per-iteration schedules (the reference steps its warm-up / cosine scheduler every iteration, func/train.py:233-234):
        the value also goes to a device scalar, so a step captured in a CUDA graph follows it without re-capturing."""
        self.lr = float(lr)
        if self.other is not None:
            for g in self.other.param_groups:
                g["lr"] = self.lr
        if self.lr_dev is not None:
            self.lr_dev.fill_(self.lr)

    def use_device_lr(self, device):
        if self.lr_dev is None:
            self.lr_dev = torch.full((1,), self.lr, dtype=torch.float32, device=device)
        return self.lr_dev

    def step(self):
        self.sync_lr()
        for i in range(len(self.mods)):
            self.step_flat(i)
        self.step_other()

    def sync_lr(self):
        if self.other is not None and self.other.param_groups[0]["lr"] != self.lr:
            self.set_lr(self.other.param_groups[0]["lr"])

    def step_flat(self, i):
        """Update flat module i (its gradients must be final, i.e. all-reduced). FlatDataParallel.finish_backward calls
        the pieces one by one so that updates run while the last gradient slices of the backbone are still on the wire."""
        pack = self.mods[i]._pack
        first = self.m[i] is None
        if first:
            self.m[i] = torch.empty_like(pack.w)
        if pack.gb is None:
            ops.sgd_step(pack.w, pack.g, self.m[i], pack.b, self.lr, self.momentum, self.wd, self.nesterov, first,
                         weight_decay_lo=self.wd_bias, lo_elems=pack.bias_end, lr_dev=self.lr_dev)
        else:   # data parallel: vector gradients in fp32 (pack.g), matrix gradients in bf16 (pack.gb)
            n = pack.small_end
            if n:
                ops.sgd_step(pack.w[:n], pack.g[:n], self.m[i][:n], pack.b[:n], self.lr, self.momentum, self.wd, self.nesterov,
                             first, weight_decay_lo=self.wd_bias, lo_elems=pack.bias_end, lr_dev=self.lr_dev)
            ops.sgd_step(pack.w[n:], pack.gb[n:], self.m[i][n:], pack.b[n:], self.lr, self.momentum, self.wd, self.nesterov, first,
                         lr_dev=self.lr_dev)
        pack.shadow_is_current()

    def step_flat_vectors(self, i):
        """Sharded mode, the replicated part: the vectors [0, small_end) of flat module i (fp32 gradients in pack.g,
        already averaged over ranks) are updated on every rank - the kernels read them from the fp32 master."""
        pack = self.mods[i]._pack
        if not hasattr(self, "m_vec"):
            self.m_vec = {}
        first = i not in self.m_vec
        if first:
            self.m_vec[i] = torch.empty(pack.small_end, dtype=torch.float32, device=pack.w.device)
        n = pack.small_end
        if n:
            ops.sgd_step(pack.w[:n], pack.g[:n], self.m_vec[i], pack.b[:n], self.lr, self.momentum, self.wd, self.nesterov, first,
                         weight_decay_lo=self.wd_bias, lo_elems=pack.bias_end, lr_dev=self.lr_dev)

    def step_flat_shard(self, i, grad_shard, rng):
        """Sharded optimizer: update elements [lo, hi) of flat module i from `grad_shard` (bf16 or fp32, hi - lo elements,
        already averaged over ranks). Momentum is kept for the shard only. The other ranks' shards of the bf16 shadow are
        stale until they are all-gathered (FlatDataParallel.begin_step)."""
        pack = self.mods[i]._pack
        lo, hi = rng
        first = self.m[i] is None
        if first:
            self.m[i] = torch.empty(hi - lo, dtype=torch.float32, device=pack.w.device)
        assert self.m[i].numel() == hi - lo and grad_shard.numel() == hi - lo
        ops.sgd_step(pack.w[lo:hi], grad_shard, self.m[i], pack.b[lo:hi], self.lr, self.momentum, self.wd, self.nesterov, first,
                     weight_decay_lo=self.wd_bias, lo_elems=min(max(pack.bias_end - lo, 0), hi - lo), lr_dev=self.lr_dev)
        pack.shadow_is_current()

    def step_other(self):
        for i, p in enumerate(self.fused_other):
            if p.grad is None:
                continue
            first = self.fused_m[i] is None
            if first:
                self.fused_m[i] = torch.empty_like(p.data)
            assert p.data.is_contiguous() and p.dtype == torch.float32
            wd = self.wd if p.dim() >= 2 else self.wd_bias
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            shadow = self.shadows.get(id(p))      # e.g. the classifier's bf16 copy (avt_b200.loss_head): refreshed in the same pass
            ops.sgd_step(p.data.view(-1), g.view(-1), self.fused_m[i].view(-1), shadow[0] if shadow else None, self.lr,
                         self.momentum, wd, self.nesterov, first, lr_dev=self.lr_dev)
            if shadow:
                shadow[1]()
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
