"""Data parallelism for the hot path: one process per GPU, clips sharded over the batch - the B200-native restatement of
the reference's DistributedDataParallel wrap (func/train.py:771-778; gradients are AVERAGED over ranks because the
learning rate is scaled by the world size, :718).

Per step and rank (N ranks), all gradient traffic in bf16 (792 MB instead of 1.585 GB of fp32):
  * AVT-h (78 % of the parameters, weight-bandwidth bound): ZeRO-1 style for its matrices (the vectors - biases, LayerNorm,
    0.03 % of the elements, which the kernels read from the fp32 master - are all-reduced in fp32 and updated on every
    rank). The gradients are complete before the backbone
    backward starts, so ONE reduce-scatter of the flat bf16 gradient buffer overlaps the whole ViT backward; the fused SGD
    then updates only this rank's 1/N shard of the fp32 master weights / momentum (HBM time of the update 1.2 ms -> 1.2/N)
    and the refreshed bf16 weights are all-gathered at the START of the next step, under the ViT forward, on a second
    communicator that is held to a few CTAs. The other ranks' fp32 master shards are brought up to date only when
    somebody asks for them (`sync_master_weights`, called from the head's state_dict hook: every rank calls
    model.state_dict() in the reference's store_checkpoint, func/train.py:61-69).
  * AVT-b: per-layer all-reduce of the bf16 copy of that layer's weight-gradient slice as soon as the layer's backward
    is done (its weight gradients are produced by split-K fp32 atomics, so they are down-cast slice by slice), the small
    rest at the end (vectors in fp32); full (replicated) fused SGD.
  * torch-owned parameters (classifier): all-reduced in fp32 from their gradient hooks at the start of the backward.
While collectives overlap compute, the persistent GEMM / attention grids are limited to the SMs NCCL leaves free.
"""
import torch
import torch.distributed as dist

from . import _lib, ops


def _sm_count():
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _gloo(group=None):
    """gloo (CPU tests, and the 2-process run that shares ONE GPU where NCCL refuses duplicate devices) has no AVG, no
    reduce-scatter and no bf16: the helpers below fall back to fp32 all-reduce / list all-gather, synchronously."""
    return dist.get_backend(group) == "gloo"


def allreduce_mean_(tensors, group=None, async_op=False):
    """In-place mean over ranks of each tensor; returns work handles when async_op."""
    if _world(group) == 1:
        return []
    ws = dist.get_world_size(group)
    handles = []
    for t in tensors:
        if not _gloo(group):
            h = dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
            if async_op:
                handles.append(h)
        else:
            t32 = t.float() if t.dtype != torch.float32 else t
            dist.all_reduce(t32, op=dist.ReduceOp.SUM, group=group)
            t32.div_(ws)
            if t32 is not t:
                t.copy_(t32)
    return handles


def shard_range(total, group=None, start=0):
    """[lo, hi) of this rank's contiguous shard of elements [start, total) of a flat buffer ((total - start) % world == 0)."""
    ws = _world(group)
    assert (total - start) % ws == 0, (total, start, ws)
    n = (total - start) // ws
    r = dist.get_rank(group) if ws > 1 else 0
    return start + r * n, start + (r + 1) * n


def reduce_scatter_mean(flat, group=None, async_op=False):
    """Mean over ranks of `flat`, this rank's shard only: returns (shard tensor, work handle or None). NCCL: one
    reduce-scatter (AVG); gloo (CPU tests) has neither reduce-scatter nor AVG: all-reduce + slice."""
    ws = _world(group)
    lo, hi = shard_range(flat.numel(), group)
    if ws == 1:
        return flat[lo:hi], None
    assert flat.is_contiguous()
    if not _gloo(group):
        out = torch.empty(hi - lo, dtype=flat.dtype, device=flat.device)
        h = dist.reduce_scatter_tensor(out, flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        return out, (h if async_op else None)
    tmp = flat.float()
    if tmp is flat:
        tmp = flat.clone()
    dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
    return tmp[lo:hi].div_(ws).to(flat.dtype), None


def all_gather_shards_(flat, group=None, async_op=False):
    """Every rank contributes its shard_range() slice of `flat`; afterwards all of `flat` is current everywhere."""
    ws = _world(group)
    if ws == 1:
        return None
    lo, hi = shard_range(flat.numel(), group)
    if not _gloo(group):
        return dist.all_gather_into_tensor(flat, flat[lo:hi], group=group, async_op=async_op)   # in place
    mine = flat[lo:hi].float()
    parts = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(parts, mine.clone(), group=group)
    flat.copy_(torch.cat(parts))
    return None


class FlatDataParallel:
    """Wraps an avt_b200.model.AVTModel-like module whose backbone.model / future_predictor own flat buffers."""

    def __init__(self, model, group=None, comm_sms=12, gather_ctas=16, bf16_head_grads=None):
        """bf16_head_grads: the AVT-h weight-gradient GEMMs store bf16 straight into the payload buffer (what torch autocast
        produces for a bf16 matmul's weight gradient). Always on for world > 1 (it IS the payload); on a single GPU it is an
        option (default off) that takes 4 B/parameter out of the weight-gradient stores and the fused SGD's reads."""
        self.model, self.group = model, group
        self.world = _world(group)
        self.bf16_head_grads = self.world > 1 if bf16_head_grads is None else (bool(bf16_head_grads) or self.world > 1)
        # SMs left to NCCL while collectives overlap the backbone backward / forward (persistent grids shrink)
        self.comm_sms = comm_sms if self.world > 1 else 0
        self.gather_ctas = gather_ctas
        self.vit, self.head = model.backbone.model, model.future_predictor
        self.vit.direct_grads = True
        self.head.direct_grads = True
        if self.bf16_head_grads:
            self.head.bf16_matrix_grads = True       # honoured when the head builds its flat buffers (first forward)
            if self.head._pack is not None and self.head._pack.gb is None:
                self.head._pack.enable_bf16_grads(matrices_direct=True)
                self.head._pack.attach_grads()
        self._handles = []          # backbone slices + torch-owned rest
        self._head_rs = None        # (gradient shard, handle) of the AVT-h reduce-scatter (matrix region)
        self._head_small = []       # handles of the all-reduce of the AVT-h vector gradients
        self._head_ag = None        # pending all-gather of the AVT-h bf16 weights
        self._shadow_shards_stale = False   # this rank updated its shard of the AVT-h weights; the others' copies are old
        self._master_stale = False
        self._layer_ranges = []
        self._gather_group = None
        self.head._grads_ready_hook = self._head_ready
        self.vit._grads_ready_hook = self._vit_ready
        self.other = [p for n, p in model.named_parameters()
                      if not n.startswith("backbone.model.") and not n.startswith("future_predictor.")]
        # The torch-owned parameters (classifier) get their gradients at the very start of the backward: reduce them
        # right there, under the whole AVT-h + backbone backward, instead of after it.
        self._other_handles, self._other_early = [], set()
        if self.world > 1:
            for p in self.other:
                p.register_post_accumulate_grad_hook(self._other_ready)
            self.head._before_forward_hook = self._head_weights_needed
            self.head.register_state_dict_pre_hook(lambda *a, **k: self.sync_master_weights())

    # ------------------------------------------------------------------ set-up
    def _packs_ready(self):
        """bf16 gradient buffers (the payload) once the flat buffers exist, i.e. after the first forward."""
        if self.world > 1:
            if self.head._pack.gb is None:      # (already there when the head built its buffers with bf16_matrix_grads set)
                self.head._pack.enable_bf16_grads(matrices_direct=True)
                self.head._pack.attach_grads()
            if self.vit._pack.gb is None:
                self.vit._pack.enable_bf16_grads(matrices_direct=False)

    def broadcast_parameters(self):
        """DDP-constructor semantics: every rank starts from rank 0's weights (valid after the first forward)."""
        if self.world > 1:
            for t in [self.vit.flat_buffers()[0], self.head.flat_buffers()[0]] + [p.data for p in self.other]:
                dist.broadcast(t, 0, group=self.group)
            for m in (self.vit, self.head):
                m._pack.refresh_bf16()       # (the in-place broadcast bumped w._version: the shadow follows)
        self._packs_ready()

    def _limit_sms(self, reserve):
        _lib.lib().avt_set_sm_limit(_sm_count() - reserve if reserve else 0)

    # ------------------------------------------------------------------ start of a step
    def begin_step(self):
        """Call before the forward. If this rank's optimizer updated only its shard of the AVT-h weights in the previous
        step, start the all-gather of the bf16 weights now: it runs under the whole backbone forward (the head's forward
        waits for it) on its own communicator, held to a few CTAs so that the forward keeps (almost) every SM."""
        if self.world > 1 and self._shadow_shards_stale:
            if self._gather_group is None:
                self._gather_group = self._make_gather_group()
            self._limit_sms(self.gather_ctas)
            pk = self.head._pack
            self._head_ag = all_gather_shards_(pk.b[pk.small_end:], self._gather_group, async_op=True)
            self._shadow_shards_stale = False

    def _make_gather_group(self):
        if _gloo(self.group):
            return self.group
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = self.gather_ctas
            opts.config.min_ctas = 1
            ranks = list(range(dist.get_world_size())) if self.group is None else dist.get_process_group_ranks(self.group)
            return dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
        except Exception:
            return self.group

    def _head_weights_needed(self):
        if self._head_ag is not None:
            self._head_ag.wait()
            self._head_ag = None
            self._limit_sms(0)

    # ------------------------------------------------------------------ backward hooks
    def _other_ready(self, p):
        if p.grad is not None:
            self._other_early.add(id(p))
            self._other_handles += allreduce_mean_([p.grad], self.group, async_op=True)

    def _head_ready(self):
        """AVT-h backward finished (the backbone backward is about to start): reduce-scatter its bf16 gradients."""
        pk = self.head._pack
        if self.world == 1:
            if pk.gb is not None:
                for lo, hi in pk.fp32_grad_ranges():
                    ops.cast_bf16(pk.g[lo:hi], pk.gb[lo:hi])
            return
        self._packs_ready()
        for lo, hi in pk.fp32_grad_ranges():        # e.g. the position-embedding gradient, accumulated in fp32
            ops.cast_bf16(pk.g[lo:hi], pk.gb[lo:hi])
        self._limit_sms(self.comm_sms)
        self._head_small = allreduce_mean_([pk.g[:pk.small_end]], self.group, async_op=True)   # vectors: fp32, replicated
        self._head_rs = reduce_scatter_mean(pk.gb[pk.small_end:], self.group, async_op=True)   # matrices: bf16, sharded

    def _vit_layer_done(self, i):
        """Layer i's weight gradients are final: reduce that slice now, overlapped with the rest of the backward."""
        lo, hi = self.vit._stack.layer_grad_range(i)
        self._layer_ranges.append((lo, hi))
        pk = self.vit._pack
        ops.cast_bf16(pk.g[lo:hi], pk.gb[lo:hi])
        self._handles += allreduce_mean_([pk.gb[lo:hi]], self.group, async_op=True)

    def _vit_ready(self):
        if self.world == 1:
            return
        pk = self.vit._pack
        self._packs_ready()
        # the vectors (biases, LayerNorm: 0.1 % of the elements, many of them zero-initialised, so a bf16-rounded gradient would
        # be a bf16-rounded parameter) travel in fp32; every matrix in bf16
        self._handles += allreduce_mean_([pk.g[:pk.small_end]], self.group, async_op=True)
        if self.vit._stack.layer_done_hook is None:      # first backward: per-layer overlap is armed from the next step on
            self.vit._stack.layer_done_hook = self._vit_layer_done
            self._layer_ranges = []
        # every matrix not covered by the per-layer slices (cls / pos, patch embedding)
        pos = pk.small_end
        for lo, hi in sorted(self._layer_ranges) + [(pk.total, pk.total)]:
            if lo > pos:
                ops.cast_bf16(pk.g[pos:lo], pk.gb[pos:lo])
                self._handles += allreduce_mean_([pk.gb[pos:lo]], self.group, async_op=True)
            pos = hi
        self._layer_ranges = []

    # ------------------------------------------------------------------ end of a step
    def finish_backward(self, optimizer=None):
        """Call after loss.backward(): reduces the remaining (torch-owned) gradients and waits for all handles.
        With `optimizer` (an avt_b200.optim.FlatSGD over [backbone.model, future_predictor] + the other parameters) the
        update is interleaved with the waits: this rank's shard of the AVT-h buffer - reduce-scattered while the backbone
        backward ran - is updated first, which hides the tail of the collectives (last backbone slices, small tensors)
        that would otherwise sit between the backward and the optimizer."""
        grads = [p.grad for p in self.other if p.grad is not None and id(p) not in self._other_early]
        other_handles = self._other_handles + allreduce_mean_(grads, self.group, async_op=True)
        self._other_handles, self._other_early = [], set()
        if optimizer is not None:
            assert [m for m in optimizer.mods] == [self.vit, self.head], "FlatSGD([dp.vit, dp.head], dp.other, ...) expected"
            optimizer.sync_lr()
        head_shard = None
        if self._head_rs is not None:
            head_shard, h = self._head_rs
            if h is not None:
                h.wait()
            self._head_rs = None
        for h in self._head_small:
            h.wait()
        self._head_small = []
        if optimizer is not None:
            if head_shard is not None:
                pk = self.head._pack
                optimizer.step_flat_vectors(1)
                optimizer.step_flat_shard(1, head_shard, shard_range(pk.total, self.group, start=pk.small_end))
                self._shadow_shards_stale = self._master_stale = True
            else:
                optimizer.step_flat(1)
        elif head_shard is not None:
            raise RuntimeError("data-parallel AVT-h gradients are reduce-scattered: pass the FlatSGD optimizer to finish_backward")
        for h in self._handles:
            h.wait()
        if optimizer is not None:
            optimizer.step_flat(0)
        for h in other_handles:
            h.wait()
        if optimizer is not None:
            optimizer.step_other()
        self._handles = []
        if self.comm_sms:
            self._limit_sms(0)

    def sync_master_weights(self):
        """All ranks: bring every rank's fp32 master copy (and bf16 shadow) of the AVT-h weights up to date (the sharded
        optimizer only touches the local 1/N). Collective - the reference calls model.state_dict() on every rank."""
        if self.world > 1 and self._master_stale:
            self._head_weights_needed()
            pk = self.head._pack
            all_gather_shards_(pk.w[pk.small_end:], self.group)
            if self._shadow_shards_stale:
                all_gather_shards_(pk.b[pk.small_end:], self.group)
                self._shadow_shards_stale = False
            self.head._pack.shadow_is_current()
            self._master_stale = False

    def flat_parameter_groups(self):
        """Three 'parameters' for a stock torch.optim optimizer: the two flat buffers (as leaf tensors whose .grad is
        the flat gradient buffer) + the torch-owned rest. Single process only (fp32 gradients). Valid after the first
        forward; the in-place optimizer step bumps the flat buffers' version counters, which refreshes the bf16 shadow."""
        assert self.world == 1, "the data-parallel path keeps bf16 gradients and a sharded AVT-h optimizer: use FlatSGD"
        out = []
        for m in (self.vit, self.head):
            w, g = m.flat_buffers()
            w.grad = g
            out.append(w)
        return out, self.other
