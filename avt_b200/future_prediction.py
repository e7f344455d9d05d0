"""AVT-h: drop-in replacement for `models.future_prediction.AVTh` (reference :51-258).

Same constructor arguments, forward contract `(feats (B,T,C), target_shape) -> (updated_past (B,T,C),
final, {'feat': ...}, {})`, `output_dim` property and state-dict names (`encoder.weight`, `decoder.weight`,
`gpt_model.wpe.weight`, `gpt_model.h.{i}.{ln_1,attn.c_attn,attn.c_proj,ln_2,mlp.c_fc,mlp.c_proj}.*`,
`gpt_model.ln_f.*`; HF Conv1D weights are [in, out]). The encoder -> GPT-2 blocks -> ln_f -> decoder chain
runs as sm_100a kernels in one autograd.Function; the cheap output assembly (slices / cat / mean / MSE,
reference :207-251) stays PyTorch glue exactly as in the reference.

Hydra: `conf/model/future_predictor/avth_b200.yaml` -> `_target_: avt_b200.future_prediction.AVTh`.
"""
import importlib

import torch
import torch.nn as nn

from . import engine, ops

_HF_NAMES = dict(ln1="gpt_model.h.{i}.ln_1", qkv="gpt_model.h.{i}.attn.c_attn", proj="gpt_model.h.{i}.attn.c_proj",
                 ln2="gpt_model.h.{i}.ln_2", fc1="gpt_model.h.{i}.mlp.c_fc", fc2="gpt_model.h.{i}.mlp.c_proj")


class Conv1D(nn.Module):
    """Parameter container with HF Conv1D's layout: weight [in, out], bias [out] (not an nn.Linear, so
    BaseModel._initialize_weights (models/base_model.py:124-127) leaves it at HF's N(0, 0.02) init)."""

    def __init__(self, nf, nx):
        super().__init__()
        self.nf = nf
        self.weight = nn.Parameter(torch.empty(nx, nf).normal_(std=0.02))
        self.bias = nn.Parameter(torch.zeros(nf))


class _Attn(nn.Module):
    def __init__(self, nx):
        super().__init__()
        self.c_attn = Conv1D(3 * nx, nx)
        self.c_proj = Conv1D(nx, nx)


class _MLP(nn.Module):
    def __init__(self, n_inner, nx):
        super().__init__()
        self.c_fc = Conv1D(n_inner, nx)
        self.c_proj = Conv1D(nx, n_inner)


class _Block(nn.Module):
    def __init__(self, nx, n_inner, eps):
        super().__init__()
        self.ln_1 = nn.LayerNorm(nx, eps=eps)
        self.attn = _Attn(nx)
        self.ln_2 = nn.LayerNorm(nx, eps=eps)
        self.mlp = _MLP(n_inner, nx)


class _GPT2(nn.Module):
    """Parameter tree of transformers.GPT2Model after `del wte` (reference :89-95)."""

    def __init__(self, n_embd, n_layer, n_positions, n_inner, eps):
        super().__init__()
        self.wpe = nn.Embedding(n_positions, n_embd)
        nn.init.normal_(self.wpe.weight, std=0.02)
        self.h = nn.ModuleList([_Block(n_embd, n_inner, eps) for _ in range(n_layer)])
        self.ln_f = nn.LayerNorm(n_embd, eps=eps)


class _HeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats2d, head, B, T, train_graph, *params):
        decoded, saved = head._run_forward(feats2d, B, T, train_graph)
        ctx.head, ctx.saved = head, saved
        return decoded

    @staticmethod
    def backward(ctx, ddec):
        head = ctx.head
        dfeats = head._run_backward(ctx.saved, ddec)
        return (dfeats, None, None, None, None) + engine.param_grads(head._pack, head._param_order, head._param_list,
                                                              head.direct_grads)


def _instantiate_loss(cfg):
    """hydra.utils.instantiate(future_pred_loss, reduction='none') (reference :101-105) without requiring hydra."""
    if cfg is None:
        return None
    if isinstance(cfg, nn.Module):
        return cfg
    if isinstance(cfg, str):
        cfg = {"_target_": cfg}
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    mod, _, attr = target.rpartition(".")
    return getattr(importlib.import_module(mod), attr)(reduction="none", **cfg)


class AVTh(nn.Module):
    def __init__(self, in_features, output_len=-1, output_len_eval=-1, avg_last_n=-1, inter_dim=768,
                 future_pred_loss=None, return_past_too=False, drop_last_n=0, quantize_before_rollout=False,
                 assign_to_centroids=None, num_cluster_centers=50000, freeze_encoder_decoder=False, **kwargs):
        super().__init__()
        if in_features == 1 or assign_to_centroids:
            raise NotImplementedError("clustered / quantized inputs (nn.Embedding encoder) are not used by any AVT config")
        if quantize_before_rollout or drop_last_n != 0:
            raise NotImplementedError("quantize_before_rollout / drop_last_n are debugging options of the reference")
        self.encoder = nn.Linear(in_features, inter_dim, bias=False)
        self.decoder = nn.Linear(inter_dim, in_features, bias=False)
        if freeze_encoder_decoder:
            self.encoder.weight.requires_grad = False
            self.decoder.weight.requires_grad = False
        # GPT2Config(n_embd=inter_dim, vocab_size=in_features, use_cache=True, **kwargs) defaults
        kwargs = dict(kwargs)
        kwargs.pop("future_pred_loss_wt", None)
        self.n_head = int(kwargs.pop("n_head", 12))
        self.n_layer = int(kwargs.pop("n_layer", 12))
        n_positions = int(kwargs.pop("n_positions", 1024))
        n_inner = kwargs.pop("n_inner", None) or 4 * inter_dim
        if n_inner != 4 * inter_dim:
            raise NotImplementedError("n_inner != 4 * n_embd")
        self.eps = float(kwargs.pop("layer_norm_epsilon", 1e-5))
        self.embd_pdrop = float(kwargs.pop("embd_pdrop", 0.1))
        self.attn_pdrop = float(kwargs.pop("attn_pdrop", 0.1))
        self.resid_pdrop = float(kwargs.pop("resid_pdrop", 0.1))
        act = kwargs.pop("activation_function", "gelu_new")
        if act != "gelu_new":
            raise NotImplementedError("only GPT-2's gelu_new activation is implemented")
        if kwargs.pop("output_attentions", False):
            raise NotImplementedError("output_attentions (visualisation only)")
        self.extra_gpt_kwargs = kwargs  # remaining GPT2Config keys have no effect on this path
        assert inter_dim % self.n_head == 0
        self.gpt_model = _GPT2(inter_dim, self.n_layer, n_positions, n_inner, self.eps)
        self.output_len, self.output_len_eval = output_len, output_len_eval
        self.avg_last_n, self.inter_dim, self.in_features = avg_last_n, inter_dim, in_features
        self.future_pred_loss = _instantiate_loss(future_pred_loss)
        self.return_past_too = return_past_too
        self.precision = "bf16"            # "fp32": inference in fp32 on the CUDA cores (engine.forward_fp32)
        self.direct_grads = False
        self._grads_ready_hook = None
        self._before_forward_hook = None   # FlatDataParallel: wait for the all-gather of the sharded-optimizer weights
        self.bf16_matrix_grads = False     # weight-gradient GEMMs store bf16 (FlatDataParallel + FlatSGD; needs direct_grads)
        self._pack = None
        self._stack = None
        self._rng_dev = None
        # Checkpoints written with the reference's transformers==4.2.2 carry the causal-mask buffers of every
        # GPT2Attention (`h.{i}.attn.bias` [1,1,n_ctx,n_ctx], `h.{i}.attn.masked_bias`); the mask is implicit in the
        # kernels here, so those keys are dropped before a (strict) load instead of failing the resume of
        # func/train.py:764 / the released checkpoint.pth files (README.md:193).
        self._register_load_state_dict_pre_hook(self._drop_hf_mask_buffers)
        self._aux = {}

    @staticmethod
    def _drop_hf_mask_buffers(state_dict, prefix, *unused):
        for k in [k for k in state_dict if k.startswith(prefix + "gpt_model.h.") and
                  (k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"))]:
            del state_dict[k]

    # ------------------------------------------------------------------ plumbing
    def _ensure_pack(self, device):
        if self._pack is not None and self._pack.intact():
            return
        named = list(self.named_parameters())
        self._param_order = [n for n, _ in named]
        self._param_list = [p for _, p in named]
        self._pack = engine.ParamPack(named, device)
        self._pack.fp32_grad_names = ("gpt_model.wpe.weight",)     # summed over the batch in fp32 (avt_frame_sum_grads)
        spec = engine.StackSpec(dim=self.inter_dim, heads=self.n_head, layers=self.n_layer, eps=self.eps,
                                act=ops.ACT_GELU_TANH, conv1d=True, causal=True, names=_HF_NAMES,
                                p_attn=self.attn_pdrop, p_resid=self.resid_pdrop, attn_impl="simt")
        self._stack = engine.BlockStack(spec, self._pack)
        self._aux = {}
        if self.bf16_matrix_grads and self.direct_grads:
            self._pack.enable_bf16_grads(matrices_direct=True)
        if self.direct_grads:
            self._pack.attach_grads()

    def flat_buffers(self):
        return self._pack.w, self._pack.g

    # ------------------------------------------------------------------ kernels
    def _run_forward(self, feats2d, B, T, train_graph):
        pk, st = self._pack, self._stack
        M, C, Dh = B * T, self.in_features, self.inter_dim
        dev = feats2d.device
        drop = self.training  # nn.Dropout semantics: active in train() mode only
        pk.refresh_bf16()
        w = st.workspace(M, B, T, train_graph)
        if w["aux"] is None:   # buffers the backward reads: they belong to the workspace, not to the module
            w["aux"] = dict(xb=torch.empty(M, C, dtype=torch.bfloat16, device=dev),
                                  lnf=torch.empty(M, Dh, dtype=torch.bfloat16, device=dev),
                                  fst=torch.empty(2, M, dtype=torch.float32, device=dev),
                                  db=torch.empty(M, C, dtype=torch.bfloat16, device=dev),
                                  dlnf=torch.empty(M, Dh, dtype=torch.bfloat16, device=dev),
                                  g32=torch.empty(M, Dh, dtype=torch.float32, device=dev),
                                  gb=torch.empty(M, Dh, dtype=torch.bfloat16, device=dev),
                                  fsum=torch.empty(T * Dh, dtype=torch.float32, device=dev))
        a = w["aux"]
        # Philox stream: seed from torch's global seed; the per-step offset lives in DEVICE memory and is advanced by a
        # (capturable) in-place add, so a CUDA-graph replay of the step draws fresh dropout masks. Each forward keeps
        # its own snapshot for its backward.
        seed = (torch.initial_seed() ^ 0x5DEECE66D) & 0xFFFFFFFFFFFF
        off, off_dev = 0, None
        if drop:
            if self._rng_dev is None or self._rng_dev.device != dev:
                self._rng_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            self._rng_dev.add_(1 << 38)
            off_dev = self._rng_dev.clone()
        p_embd = self.embd_pdrop if drop else 0.0
        ops.cast_bf16(feats2d.contiguous().float(), a["xb"])
        # h0 = dropout(encoder(feats) + wpe[0:T])   (reference :163 + HF GPT2Model.forward)
        sk = engine.small_m_split(M, Dh, C)
        ops.gemm(a["xb"], pk.bv("encoder.weight"), w["x"][0], pos=pk.wv("gpt_model.wpe.weight")[:T], pos_period=T,
                 drop_p=p_embd, drop_seed=seed, drop_offset=off + (255 << 28), drop_offset_dev=off_dev, split_k=sk,
                 workspace=st._gemm_ws(w["x"][0], sk))
        xmid, y = st.forward(w, B, T, train_graph, rng=(seed, off, off_dev), dropout=drop)
        xf = st._xbuf(w, train_graph, 2 * self.n_layer)
        ops.layernorm_fwd(xmid, pk.wv("gpt_model.ln_f.weight"), pk.wv("gpt_model.ln_f.bias"), self.eps, a["lnf"],
                          a["fst"][0], a["fst"][1], add=y, x_out=xf if y is not None else None)   # (y None: xmid IS xf)
        decoded = torch.empty(M, C, dtype=torch.float32, device=dev)
        sk = engine.small_m_split(M, C, Dh)
        ops.gemm(a["lnf"], pk.bv("decoder.weight"), decoded, split_k=sk, workspace=st._gemm_ws(decoded, sk))
        return decoded, (w, st.lease(w) if train_graph else None, xf, B, T, p_embd, seed, off, off_dev)

    def _run_forward_fp32(self, feats2d, B, T):
        pk, st = self._pack, self._stack
        M, C, Dh = B * T, self.in_features, self.inter_dim
        x = torch.empty(M, Dh, dtype=torch.float32, device=feats2d.device)
        ops.sgemm_f32(feats2d.contiguous().float(), pk.wv("encoder.weight"), x, pos=pk.wv("gpt_model.wpe.weight")[:T], pos_period=T)
        x = st.forward_fp32(x, B, T)
        lnf = torch.empty_like(x)
        ops.layernorm_fwd(x, pk.wv("gpt_model.ln_f.weight"), pk.wv("gpt_model.ln_f.bias"), self.eps, lnf)
        decoded = torch.empty(M, C, dtype=torch.float32, device=x.device)
        ops.sgemm_f32(lnf, pk.wv("decoder.weight"), decoded)
        return decoded

    def _run_backward(self, saved, ddec):
        pk, st = self._pack, self._stack
        w, lease, xf, B, T, p_embd, seed, off, off_dev = saved
        st.check_lease(w, lease)
        M, C, Dh = B * T, self.in_features, self.inter_dim
        a = w["aux"]
        pk.zero_small_grads()
        ops.cast_bf16(ddec.contiguous().float(), a["db"])
        ops.gemm(a["db"], a["lnf"], pk.grad_out("decoder.weight"), a_mn=True, b_mn=True)    # dWdec = ddec^T lnf
        sk = engine.small_m_split(M, Dh, C)
        ops.gemm(a["db"], pk.bv("decoder.weight"), a["dlnf"], b_mn=True, split_k=sk,        # dlnf = ddec Wdec
                 workspace=st._gemm_ws(a["dlnf"], sk))
        dx, dxb = w["dx"], w["dxb"]
        ops.layernorm_bwd(a["dlnf"], xf, a["fst"][0], a["fst"][1], pk.wv("gpt_model.ln_f.weight"), dx,
                          pk.gv("gpt_model.ln_f.weight"), pk.gv("gpt_model.ln_f.bias"), w["lnws"], dx_bf16=dxb)
        st.backward(w, dx, dxb)
        g32, gb = dx, dxb
        if p_embd > 0.0:
            g32, gb = a["g32"], a["gb"]
            ops.dropout_apply(dx, p_embd, seed, off + (255 << 28), y_f32=g32, y_bf16=gb, offset_dev=off_dev)
        ops.frame_sum_grads(g32, B, T, Dh, a["fsum"], dpos=pk.gv("gpt_model.wpe.weight")[:T], accumulate=False)
        ops.gemm(gb, a["xb"], pk.grad_out("encoder.weight"), a_mn=True, b_mn=True)           # dWenc = g^T feats
        dfeats = torch.empty(M, C, dtype=torch.float32, device=ddec.device)
        sk = engine.small_m_split(M, C, Dh)
        ops.gemm(gb, pk.bv("encoder.weight"), dfeats, b_mn=True, split_k=sk,                 # dfeats = g Wenc
                 workspace=st._gemm_ws(dfeats, sk))
        lease.done = True
        if self._grads_ready_hook is not None:
            self._grads_ready_hook()
        return dfeats

    def _run_rollout(self, feats2d, B, T, output_len):
        """Autoregressive rollout for evaluation (reference :168-202: `output_len` GPT-2 calls, each after the first fed the
        LAST hidden state of the previous one with `past_key_values` and continuing position ids). KV-cached: the first call
        is one causal pass over the T input tokens that keeps every layer's packed qkv rows (they ARE the KV cache); every
        further call pushes ONE token per clip through the stack (engine.BlockStack.decode_step: M = B rows, the new q/k/v
        written into the cache, a single-row attention against the cached keys). Returns decoded outputs
        [B, T + output_len - 1, C] (fp32). No autograd graph, dropout off (eval)."""
        pk, st = self._pack, self._stack
        C, Dh, L = self.in_features, self.inter_dim, self.n_layer
        dev = feats2d.device
        tmax = T + output_len - 1
        pk.refresh_bf16()
        wpe = pk.wv("gpt_model.wpe.weight")
        # ---- call 0: prefill (the training-mode workspace keeps one qkv buffer per layer; dropout stays off)
        w = st.workspace(B * T, B, T, True)
        xb = torch.empty(B * T, C, dtype=torch.bfloat16, device=dev)
        ops.cast_bf16(feats2d.contiguous().float(), xb)
        sk = engine.small_m_split(B * T, Dh, C)
        ops.gemm(xb, pk.bv("encoder.weight"), w["x"][0], pos=wpe[:T], pos_period=T, split_k=sk,
                 workspace=st._gemm_ws(w["x"][0], sk))
        xmid, y = st.forward(w, B, T, True)
        hid = torch.empty(B * T, Dh, dtype=torch.float32, device=dev)           # ln_f output, fp32 (fed back below)
        ops.layernorm_fwd(xmid, pk.wv("gpt_model.ln_f.weight"), pk.wv("gpt_model.ln_f.bias"), self.eps, hid, add=y,
                          x_out=st._xbuf(w, True, 2 * L) if y is not None else None)

        def decode(rows32):
            hb = torch.empty(rows32.shape[0], Dh, dtype=torch.bfloat16, device=dev)
            ops.cast_bf16(rows32.contiguous(), hb)
            dec = torch.empty(rows32.shape[0], C, dtype=torch.float32, device=dev)
            sk = engine.small_m_split(rows32.shape[0], C, Dh)
            ops.gemm(hb, pk.bv("decoder.weight"), dec, split_k=sk, workspace=st._gemm_ws(dec, sk))
            return dec

        decoded_all = decode(hid).view(B, T, C)
        if output_len == 1:
            return decoded_all
        caches = []
        for l in range(L):                      # KV cache = the prefill's packed qkv rows, padded to tmax positions per clip
            c = torch.zeros(B * tmax, 3 * Dh, dtype=torch.bfloat16, device=dev)
            c.view(B, tmax, 3 * Dh)[:, :T].copy_(w["qkv"][l].view(B, T, 3 * Dh))
            caches.append(c)
        att = torch.zeros(B * tmax, Dh, dtype=torch.bfloat16, device=dev)
        last = hid.view(B, T, Dh)[:, -1]
        for i in range(1, output_len):
            pos = T + i - 1
            x = (last + wpe[pos].view(1, Dh)).contiguous()                      # next input: last hidden state + wpe[pos]  (:202, :169-173)
            xmid, y = st.decode_step(x, caches, att, B, tmax, pos)
            last = torch.empty(B, Dh, dtype=torch.float32, device=dev)
            ops.layernorm_fwd(xmid, pk.wv("gpt_model.ln_f.weight"), pk.wv("gpt_model.ln_f.bias"), self.eps, last, add=y)
            decoded_all = torch.cat([decoded_all, decode(last).view(B, 1, C)], dim=1)
        return decoded_all

    # ------------------------------------------------------------------ reference-compatible forward
    def forward(self, feats, target_shape):
        if not feats.is_cuda:
            raise RuntimeError("avt_b200 AVTh runs on CUDA (sm_100a) only; there is no CPU path")
        if feats.ndim == 2:
            feats = feats.unsqueeze(1)                                   # reference :119-121
        if len(target_shape) == 3:                                       # :123-130
            output_len = target_shape[1]
        elif self.training or self.output_len_eval < 0:
            output_len = self.output_len
        else:
            output_len = self.output_len_eval
        if output_len < 1:
            raise NotImplementedError("output_len must be >= 1 (the reference's output_len <= 0 returns nothing to decode)")
        B, T, C = feats.shape
        if self._before_forward_hook is not None:
            self._before_forward_hook()
        self._ensure_pack(feats.device)
        full_orig_feats = inp_feats = feats
        orig_feats_len = T
        train_graph = torch.is_grad_enabled() and (feats.requires_grad or any(p.requires_grad for p in self._param_list))
        if self.precision == "fp32":
            if train_graph or self.training or output_len != 1:
                raise NotImplementedError("precision='fp32' is the inference-only validation mode (eval(), no_grad, output_len 1)")
            decoded = self._run_forward_fp32(feats.reshape(B * T, C), B, T).view(B, T, C)
        elif output_len == 1:
            decoded = _HeadFunction.apply(feats.reshape(B * T, C), self, B, T, train_graph, *self._param_list).view(B, T, C)
        else:
            if train_graph or self.training:
                raise NotImplementedError("autoregressive rollout (output_len > 1) is implemented for evaluation "
                                          "(model.eval() under torch.no_grad(), func/train.py:357); every shipped AVT "
                                          "experiment trains with output_len=1")
            if T + output_len - 1 > self.gpt_model.wpe.weight.shape[0]:
                raise ValueError("rollout exceeds n_positions")
            decoded = self._run_rollout(feats.reshape(B * T, C), B, T, output_len)
        all_outputs = decoded                                            # :227-229
        losses = {}
        if self.future_pred_loss is not None:                            # :207-215
            n = min(full_orig_feats.size(1), all_outputs.size(1))
            losses = {"feat": self.future_pred_loss(all_outputs[:, :n - 1], full_orig_feats[:, 1:n])}
        prev = inp_feats
        if self.return_past_too:                                         # :232-236
            final = torch.cat((prev, all_outputs[:, orig_feats_len - 1:, :]), dim=1)
        elif output_len > 0:
            final = all_outputs[:, -output_len:]
        else:
            final = all_outputs
        if self.avg_last_n > 0:                                          # :241-242
            final = torch.mean(final[:, -self.avg_last_n:, :], dim=1)
        updated_past_feat = torch.cat([prev[:, :1, :], all_outputs[:, :(orig_feats_len - 1)]], dim=1)  # :249-250
        return updated_past_feat, final, losses, {}

    @property
    def output_dim(self):
        return self.in_features
