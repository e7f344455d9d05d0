"""Data parallelism for the hot path: one process per GPU, clips sharded over the batch, ONE gradient
all-reduce (mean) per flat buffer per step over NCCL / NVLink — the B200-native restatement of the reference's
DistributedDataParallel wrap (func/train.py:771-778; mean over ranks because lr is scaled by world size, :718).

The AVT-h gradients (78 % of the bytes) are complete before the ViT backward starts, so their all-reduce is
launched from a hook inside the head's backward and overlaps the whole backbone backward.
"""
import torch
import torch.distributed as dist

from . import _lib


def _sm_count():
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


def allreduce_mean_(tensors, group=None, async_op=False):
    """In-place mean over ranks of each tensor; returns work handles when async_op."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return []
    ws = dist.get_world_size(group)
    handles = []
    for t in tensors:
        if t.is_cuda:
            h = dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        else:  # gloo has no AVG
            h = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
            if async_op:
                h.wait()
                h = None
            t.div_(ws)
        if async_op and h is not None:
            handles.append(h)
    return handles


class FlatDataParallel:
    """Wraps an avt_b200.model.AVTModel-like module whose backbone.model / future_predictor own flat buffers."""

    def __init__(self, model, group=None, comm_sms=16):
        self.model, self.group = model, group
        # SMs left to NCCL while gradient all-reduces overlap the backbone backward (persistent GEMM grids shrink)
        self.comm_sms = comm_sms if (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1) else 0
        self.vit, self.head = model.backbone.model, model.future_predictor
        self.vit.direct_grads = True
        self.head.direct_grads = True
        self._handles = []          # backbone slices + torch-owned rest
        self._head_handles = []     # AVT-h flat buffer (issued first, finished first)
        self._layer_ranges = []
        self.head._grads_ready_hook = self._head_ready
        self.vit._grads_ready_hook = self._vit_ready
        self.other = [p for n, p in model.named_parameters()
                      if not n.startswith("backbone.model.") and not n.startswith("future_predictor.")]
        # The torch-owned parameters (classifier) get their gradients at the very start of the backward: reduce them
        # right there, under the whole AVT-h + backbone backward, instead of after it.
        self._other_handles, self._other_early = [], set()
        if self.comm_sms:
            for p in self.other:
                p.register_post_accumulate_grad_hook(self._other_ready)

    def broadcast_parameters(self):
        """DDP-constructor semantics: every rank starts from rank 0's weights (valid after the first forward)."""
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            for t in [self.vit.flat_buffers()[0], self.head.flat_buffers()[0]] + [p.data for p in self.other]:
                dist.broadcast(t, 0, group=self.group)

    def _other_ready(self, p):
        if p.grad is not None:
            self._other_early.add(id(p))
            self._other_handles += allreduce_mean_([p.grad], self.group, async_op=True)

    def _head_ready(self):
        if self.comm_sms:
            _lib.lib().avt_set_sm_limit(_sm_count() - self.comm_sms)
        self._head_handles += allreduce_mean_([self.head.flat_buffers()[1]], self.group, async_op=True)

    def _vit_layer_done(self, i):
        """Layer i's weight gradients are final: reduce that slice now, overlapped with the rest of the backward."""
        lo, hi = self.vit._stack.layer_grad_range(i)
        self._layer_ranges.append((lo, hi))
        self._handles += allreduce_mean_([self.vit.flat_buffers()[1][lo:hi]], self.group, async_op=True)

    def _vit_ready(self):
        g = self.vit.flat_buffers()[1]
        if self.vit._stack.layer_done_hook is None:      # first backward: per-layer overlap is armed from the next step on
            self.vit._stack.layer_done_hook = self._vit_layer_done
            self._handles += allreduce_mean_([g], self.group, async_op=True)
            return
        # everything not covered by the per-layer slices (biases, LayerNorm, cls/pos, patch embedding, final norm)
        rest, pos = [], 0
        for lo, hi in sorted(self._layer_ranges):
            if lo > pos:
                rest.append(g[pos:lo])
            pos = hi
        if pos < g.numel():
            rest.append(g[pos:])
        self._layer_ranges = []
        self._handles += allreduce_mean_(rest, self.group, async_op=True)

    def finish_backward(self, optimizer=None):
        """Call after loss.backward(): reduces the remaining (torch-owned) gradients and waits for all handles.
        With `optimizer` (an avt_b200.optim.FlatSGD over [backbone.model, future_predictor] + the other parameters) the
        update is interleaved with the waits: the AVT-h buffer - reduced while the backbone backward ran - is updated
        first (1.1 ms of HBM time), which hides the tail of the collective (last backbone slices, small tensors) that
        would otherwise sit between the backward and the optimizer."""
        grads = [p.grad for p in self.other if p.grad is not None and id(p) not in self._other_early]
        other_handles = self._other_handles + allreduce_mean_(grads, self.group, async_op=True)
        self._other_handles, self._other_early = [], set()
        if optimizer is not None:
            assert [m for m in optimizer.mods] == [self.vit, self.head], "FlatSGD([dp.vit, dp.head], dp.other, ...) expected"
            optimizer.sync_lr()
        for h in self._head_handles:
            h.wait()
        if optimizer is not None:
            optimizer.step_flat(1)
        for h in self._handles:
            h.wait()
        if optimizer is not None:
            optimizer.step_flat(0)
        for h in other_handles:
            h.wait()
        if optimizer is not None:
            optimizer.step_other()
        self._handles, self._head_handles = [], []
        if self.comm_sms:
            _lib.lib().avt_set_sm_limit(0)

    def flat_parameter_groups(self):
        """Three 'parameters' for a stock torch.optim optimizer: the two flat buffers (as leaf tensors whose .grad is
        the flat gradient buffer) + the torch-owned rest. Valid after the first forward."""
        out = []
        for m in (self.vit, self.head):
            w, g = m.flat_buffers()
            w.grad = g
            out.append(w)
        return out, self.other
