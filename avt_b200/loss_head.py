"""Fused classifier + cross-entropy head of the AVT training step.

Reference: `models/base_model.py:203-216` applies Dropout(0.2) and the classifier Linear(768 -> 3806) to the T past
features and the 1 future feature of every clip; `func/train_eval_ops.py:57-85` turns the logits into
CrossEntropyLoss(ignore_index=-1, reduction='none') losses (`loss_fn/multidim_xentropy.py:10-25`) and top-1 / top-5
accuracies (`common/utils.py:17-44`), `func/train.py:207-217` takes their means. In eager PyTorch that is ~40 small
ATen / cutlass-simt launches per step (fp32 sgemm x3, log_softmax fwd/bwd, nll fwd/bwd, topk, masked fills, sums).

Here: ONE dropout+down-cast kernel, the library's tcgen05 GEMM for the logits of all B*(T+1) rows at once, ONE kernel that
turns each row of logits into loss, top-k rank AND the bf16 logits gradient (avt_softmax_xent), and in the backward the
two gradient GEMMs (the weight-gradient GEMM also sums the bias gradient, the input-gradient GEMM re-applies the dropout
mask in its epilogue): 8 launches. The classifier stays an ordinary nn.Linear (its fp32 weight is the master; a bf16 copy
padded to a multiple of 32 classes (GEMM tile / TMA stride granularity) follows it through the version counter or the fused
SGD). The logits exist only as one fp32 [rows, 3808] scratch tensor between two kernels (1.3 MB, L2-resident).
"""
import torch

from . import ops


class _FusedLinearXent(torch.autograd.Function):
    """(x fp32 [R, K], weight [C, K], bias [C], target int64 [R], row_scale fp32 [R]) -> (loss [R], rank int32 [R]).
    row_scale[r] = the coefficient of loss[r] in the training loss (1/B, 1/(B*T): the means of func/train.py:207-209): it is
    folded into the logits gradient at forward time; the backward re-scales by (incoming gradient / row_scale), which is
    exactly 1 in the reference's loss, so any other use of `loss` still differentiates correctly."""

    @staticmethod
    def forward(ctx, x, weight, bias, target, row_scale, head, p_drop):
        from . import engine
        R, K = x.shape
        C = weight.shape[0]
        Cp = head.classes_padded
        dev = x.device
        wb = head.weight_bf16(weight)
        seed, off = head.next_philox(p_drop > 0.0)
        xd = torch.empty(R, K, dtype=torch.bfloat16, device=dev)
        ops.dropout_apply(x.contiguous().float(), p_drop, seed, 0, y_bf16=xd, offset_dev=off)       # dropout + bf16
        logits = head.scratch("logits", (R, Cp), torch.float32, dev)
        sk = engine.small_m_split(R, Cp, K)
        ops.gemm(xd, wb, logits, bias=head.bias_padded(bias), split_k=sk, workspace=head.gemm_ws(sk * R * Cp, dev))
        loss = torch.empty(R, dtype=torch.float32, device=dev)
        rank = torch.empty(R, dtype=torch.int32, device=dev)
        need_grad = x.requires_grad or weight.requires_grad
        dlogits = torch.empty(R, Cp, dtype=torch.bfloat16, device=dev) if need_grad else None
        ops.softmax_xent(logits, C, target, loss, row_scale=row_scale, rank=rank, dlogits=dlogits)
        if need_grad:
            ctx.save_for_backward(xd, dlogits, wb, row_scale, off if off is not None else torch.empty(0))
        ctx.meta = (head, p_drop, seed, off is not None, C, K)
        ctx.mark_non_differentiable(rank)
        return loss, rank

    @staticmethod
    def backward(ctx, gloss, _grank):
        from . import engine
        xd, dlogits, wb, row_scale, off = ctx.saved_tensors
        head, p_drop, seed, has_off, C, K = ctx.meta
        R, Cp = dlogits.shape
        dev = xd.device
        dlogits = dlogits * (gloss / row_scale).to(torch.bfloat16)[:, None]      # x 1 in the reference loss (see class doc)
        dWp = head.scratch("dW", (Cp, K), torch.float32, dev)
        dbp = head.scratch("db", (Cp,), torch.float32, dev)
        dbp.zero_()
        ops.gemm(dlogits, xd, dWp, a_mn=True, b_mn=True, a_colsum=dbp)          # dW = dlogits^T x~ ; db = column sums of dlogits
        dx = torch.empty(R, K, dtype=torch.float32, device=dev)
        sk = engine.small_m_split(R, K, Cp)
        ops.gemm(dlogits, wb, dx, b_mn=True, drop_p=p_drop, drop_seed=seed, drop_offset=0,
                 drop_offset_dev=off if has_off else None, split_k=sk,
                 workspace=head.gemm_ws(sk * R * K, dev))                       # dx = (dlogits W) o mask / (1 - p)
        # (views of the scratch buffers: autograd accumulates them into .grad right away; the next backward overwrites them)
        return dx, dWp[:C], dbp[:C], None, None, None, None


class FusedClassifierLoss:
    """Holds the bf16 shadow / scratch buffers for one classifier nn.Linear and evaluates the AVT classification losses."""

    def __init__(self, linear):
        self.linear = linear
        C = linear.weight.shape[0]
        self.classes = C
        self.classes_padded = (C + 31) // 32 * 32      # the GEMM wants N % 32 == 0 (TMA: 16-byte row strides)
        self._wb = self._bp = None
        self._sig = None
        self._scratch = {}
        self._rng_dev = None

    # ------------------------------------------------------------------ buffers
    def scratch(self, name, shape, dtype, dev):
        key = (name, tuple(shape), dtype)
        t = self._scratch.get(key)
        if t is None or t.device != dev:
            t = self._scratch[key] = torch.zeros(*shape, dtype=dtype, device=dev)
        return t

    def gemm_ws(self, n, dev):
        if n <= 0:
            return None
        t = self._scratch.get("gemm_ws")
        if t is None or t.numel() < n or t.device != dev:
            t = self._scratch["gemm_ws"] = torch.empty(n, dtype=torch.float32, device=dev)
        return t

    def weight_bf16(self, weight):
        sig = (weight._version, weight.data_ptr())
        if self._wb is None or self._wb.device != weight.device or sig != self._sig:
            if self._wb is None or self._wb.device != weight.device:
                self._wb = torch.zeros(self.classes_padded, weight.shape[1], dtype=torch.bfloat16, device=weight.device)
            with torch.no_grad():
                ops.cast_bf16(weight.detach().contiguous().view(-1), self._wb[:self.classes].view(-1))
            self._sig = sig
        return self._wb

    def shadow_written_by_optimizer(self):
        """The fused SGD refreshed the bf16 rows itself (avt_sgd_step's shadow output)."""
        w = self.linear.weight
        self._sig = (w._version, w.data_ptr())

    def bias_padded(self, bias):
        if self._bp is None or self._bp.device != bias.device:
            self._bp = torch.zeros(self.classes_padded, dtype=torch.float32, device=bias.device)
        with torch.no_grad():
            self._bp[:self.classes].copy_(bias.detach())
        return self._bp

    def next_philox(self, active):
        """(seed, device-resident offset or None): a fresh dropout mask per call, also under CUDA-graph replay."""
        seed = (torch.initial_seed() ^ 0x2545F4914F6C) & 0xFFFFFFFFFFFF
        if not active:
            return seed, None
        dev = self.linear.weight.device
        if self._rng_dev is None or self._rng_dev.device != dev:
            self._rng_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._rng_dev.add_(1 << 36)
        return seed, self._rng_dev.clone()

    # ------------------------------------------------------------------ the op
    def __call__(self, past, future, past_target, target, p_drop):
        """past (B, T, K), future (B, K); past_target (B, T) and target (B,) int64 with -1 = ignored.
        Returns (mean CE of the future logits over B, mean CE of the past logits over B*T, acc1, acc5 of the future logits
        in percent) - `func/train_eval_ops.py:57-85` + the means of `func/train.py:207-209`."""
        B, T, K = past.shape
        x = torch.cat([future.reshape(B, K), past.reshape(B * T, K)], dim=0)
        tgt = torch.cat([target.reshape(B), past_target.reshape(B * T)], dim=0).contiguous()
        scale = torch.empty(B * (T + 1), dtype=torch.float32, device=x.device)
        scale[:B] = 1.0 / B
        scale[B:] = 1.0 / (B * T)
        loss, rank = _FusedLinearXent.apply(x, self.linear.weight, self.linear.bias, tgt, scale, self, float(p_drop))
        loss_future = (loss[:B] * scale[:B]).sum()
        loss_past = (loss[B:] * scale[B:]).sum()
        with torch.no_grad():
            acc1 = (rank[:B] < 1).sum(dtype=torch.float32) * (100.0 / B)
            acc5 = (rank[:B] < min(5, self.classes)).sum(dtype=torch.float32) * (100.0 / B)
        return loss_future, loss_past, acc1, acc5
