"""AVT-b: drop-in replacement for the reference's ViT backbone slot.

`TIMMModel` mirrors `models/video_classification.py:249-257` (constructor signature, `.model` attribute,
(B, C, T, H, W) -> (B, C', T, 1, 1) contract of `process_each_frame`, :213-227); `VisionTransformer` mirrors the
parameter names/shapes of timm 0.4.12's ViT so `train.init_from_model=[[backbone.model, *.pth]]`
(func/train.py:669-688) and checkpoints load unchanged. The modules only *own* parameters: forward and
backward are sequences of sm_100a kernels (avt_b200.engine), wrapped in one autograd.Function.

Hydra: `conf/model/backbone/avt_b_b200.yaml` -> `_target_: avt_b200.backbone.TIMMModel`.
"""
import torch
import torch.nn as nn

from . import engine, ops

# model_type -> (img, patch, dim, depth, heads, representation_size)   [timm 0.4.12 default_cfgs / model fns]
VIT_CONFIGS = {
    "vit_base_patch16_224": (224, 16, 768, 12, 12, None),
    "vit_base_patch16_224_in21k": (224, 16, 768, 12, 12, None),
    "vit_large_patch16_224": (224, 16, 1024, 24, 16, None),
    "vit_large_patch16_224_in21k": (224, 16, 1024, 24, 16, None),
    "vit_small_patch16_224": (224, 16, 384, 12, 6, None),
    "vit_test_patch16_32": (32, 16, 64, 2, 2, None),
    "vit_test_patch16_64": (64, 16, 128, 3, 2, None),
}

_TIMM_NAMES = dict(ln1="blocks.{i}.norm1", qkv="blocks.{i}.attn.qkv", proj="blocks.{i}.attn.proj",
                   ln2="blocks.{i}.norm2", fc1="blocks.{i}.mlp.fc1", fc2="blocks.{i}.mlp.fc2")


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, mlp_ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _PatchEmbed(nn.Module):
    def __init__(self, patch, in_chans, dim):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)


class _ViTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, vit, train, *params):
        feats, saved = vit._run_forward(x, train)
        ctx.vit, ctx.saved = vit, saved
        return feats

    @staticmethod
    def backward(ctx, dfeats):
        vit = ctx.vit
        vit._run_backward(ctx.saved, dfeats)
        return (None, None, None) + engine.param_grads(vit._pack, vit._param_order, vit._param_list, vit.direct_grads)


class VisionTransformer(nn.Module):
    """Parameter container with timm's names + the kernel-sequenced forward/backward."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 representation_size=None):
        super().__init__()
        if representation_size:
            raise NotImplementedError("pre_logits (representation_size) is not on the AVT hot path (timm 0.4.12 "
                                      "vit_*_in21k defs have none)")
        self.img_size, self.patch_size, self.in_chans = img_size, patch_size, in_chans
        self.embed_dim = self.num_features = embed_dim
        self.depth, self.num_heads = depth, num_heads
        self.num_tokens = (img_size // patch_size) ** 2 + 1
        self.patch_embed = _PatchEmbed(patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_tokens, embed_dim))
        self.blocks = nn.Sequential(*[_Block(embed_dim, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)
        self.precision = "bf16"     # "fp32": inference in fp32 on the CUDA cores (north star's 1e-5 bound; engine.forward_fp32)
        self.direct_grads = False   # True: backward writes param.grad (views of one flat buffer) itself
        self._grads_ready_hook = None  # called at the end of backward (parallel.FlatDataParallel)
        self.attn_impl = "tc"
        self._pack = None
        self._stack = None
        self._aux = {}

    # ------------------------------------------------------------------ plumbing
    def _ensure_pack(self, device):
        if self._pack is not None and self._pack.intact():
            return
        named = list(self.named_parameters())
        self._param_order = [n for n, _ in named]
        self._param_list = [p for _, p in named]
        self._pack = engine.ParamPack(named, device)
        spec = engine.StackSpec(dim=self.embed_dim, heads=self.num_heads, layers=self.depth, eps=1e-6,
                                act=ops.ACT_GELU_ERF, conv1d=False, causal=False, names=_TIMM_NAMES,
                                attn_impl=self.attn_impl)
        self._stack = engine.BlockStack(spec, self._pack)
        self._stack.grads_prezeroed = True     # _run_backward zeroes the whole flat gradient buffer first
        self._aux = {}
        if self.direct_grads:
            self._pack.attach_grads()

    def flat_buffers(self):
        """(weights fp32, grads fp32) flat tensors (valid after the first forward)."""
        return self._pack.w, self._pack.g

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("avt_b200 VisionTransformer runs on CUDA (sm_100a) only; there is no CPU path")
        self._ensure_pack(x.device)
        # (grad mode is off inside autograd.Function.forward, so decide here whether to keep activations)
        train = torch.is_grad_enabled() and any(p.requires_grad for p in self._param_list)
        if self.precision == "fp32":
            if train:
                raise NotImplementedError("precision='fp32' is the inference-only validation mode: call it under torch.no_grad()")
            return self._run_forward_fp32(x)
        return _ViTFunction.apply(x, self, train, *self._param_list)

    def _run_forward_fp32(self, x):
        pk, st = self._pack, self._stack
        F, ntok, D = x.shape[0], self.num_tokens, self.embed_dim
        Kp = self.in_chans * self.patch_size ** 2
        x = x.contiguous().float()
        A = torch.empty(F * ntok, Kp, dtype=torch.float32, device=x.device)
        ops.patchify_f32(x, A, self.patch_size)
        h = torch.empty(F * ntok, D, dtype=torch.float32, device=x.device)
        ops.sgemm_f32(A, pk.wv("patch_embed.proj.weight").view(D, Kp), h, bias=pk.wv("patch_embed.proj.bias"),
                      pos=pk.wv("pos_embed").view(ntok, D), cls=pk.wv("cls_token").view(D), pos_period=ntok)
        h = st.forward_fp32(h, F, ntok)
        feats = torch.empty(F, D, dtype=torch.float32, device=x.device)
        ops.layernorm_fwd(h, pk.wv("norm.weight"), pk.wv("norm.bias"), 1e-6, feats, rows=F, x_stride=ntok * D)   # CLS rows
        return feats

    # ------------------------------------------------------------------ kernels
    def _run_forward(self, x, train):
        pk, st = self._pack, self._stack
        F = x.shape[0]
        ntok, D = self.num_tokens, self.embed_dim
        M = F * ntok
        Kp = self.in_chans * self.patch_size ** 2
        x = x.contiguous().float()
        assert tuple(x.shape[1:]) == (self.in_chans, self.img_size, self.img_size), x.shape
        pk.refresh_bf16()
        w = st.workspace(M, F, ntok, train)
        if w["aux"] is None:   # buffers the backward reads: they belong to the workspace, not to the module
            w["aux"] = dict(patch=torch.empty(M, Kp, dtype=torch.bfloat16, device=x.device),
                            fstat=torch.empty(2, F, dtype=torch.float32, device=x.device),
                            xcls=torch.empty(F, D, dtype=torch.float32, device=x.device))
        if ("fsum", ntok) not in self._aux:   # scratch of one kernel call
            self._aux[("fsum", ntok)] = torch.empty(ntok * D, dtype=torch.float32, device=x.device)
        A = w["aux"]["patch"]
        ops.patchify(x, A, self.patch_size)
        ops.gemm(A, pk.bv("patch_embed.proj.weight").view(D, Kp), w["x"][0], bias=pk.wv("patch_embed.proj.bias"),
                 pos=pk.wv("pos_embed").view(ntok, D), cls=pk.wv("cls_token").view(D), pos_period=ntok)
        xmid, y = st.forward(w, F, ntok, train)
        feats = torch.empty(F, D, dtype=torch.float32, device=x.device)
        fst, xcls = w["aux"]["fstat"], w["aux"]["xcls"]
        # final residual add + LayerNorm on the CLS rows only (timm: norm(x)[:, 0]; the other 196 rows are dead)
        ops.layernorm_fwd(xmid, pk.wv("norm.weight"), pk.wv("norm.bias"), 1e-6, feats, fst[0], fst[1], rows=F,
                          x_stride=ntok * D, add=y, add_stride=ntok * D, x_out=xcls)
        return feats, (w, st.lease(w) if train else None, F)

    def _run_backward(self, saved, dfeats):
        pk, st = self._pack, self._stack
        w, lease, F = saved
        st.check_lease(w, lease)
        xcls = w["aux"]["xcls"]
        ntok, D = self.num_tokens, self.embed_dim
        M = F * ntok
        Kp = self.in_chans * self.patch_size ** 2
        dfeats = dfeats.contiguous().float()
        pk.zero_all_grads()
        dx, dxb = w["dx"], w["dxb"]
        ops.zero_(dx)
        ops.zero_(dxb)
        fst = w["aux"]["fstat"]
        # only the CLS rows of dx are non-zero here, so their column sums are the last fc2's bias gradient
        ops.layernorm_bwd(dfeats, xcls, fst[0], fst[1], pk.wv("norm.weight"), dx, pk.gv("norm.weight"), pk.gv("norm.bias"),
                          w["lnws"], dx_bf16=dxb, rows=F, dx_stride=ntok * D, dxb_stride=ntok * D,
                          dx_colsum=pk.gv(f"blocks.{len(self.blocks) - 1}.mlp.fc2.bias"))
        st.backward(w, dx, dxb, top_bias_done=True)
        A = w["aux"]["patch"]
        sk = engine._split_k_for(D, Kp, M, 256)
        ops.gemm(dxb, A, pk.gv("patch_embed.proj.weight").view(D, Kp), a_mn=True, b_mn=True, split_k=sk, accumulate=sk > 1)
        ops.frame_sum_grads(dx, F, ntok, D, self._aux[("fsum", ntok)], dpos=pk.gv("pos_embed"), dcls=pk.gv("cls_token"),
                            dbias=pk.gv("patch_embed.proj.bias"), accumulate=False)
        lease.done = True
        if self._grads_ready_hook is not None:
            self._grads_ready_hook()


def create_model(model_type, num_classes=0, **kw):
    """Counterpart of timm.create_model(model_type, num_classes=0) for the ViT families AVT uses."""
    if num_classes not in (0, None):
        raise NotImplementedError("classification head is not on the AVT hot path (drop_cls=True)")
    if model_type not in VIT_CONFIGS:
        raise NotImplementedError(f"unsupported model_type {model_type!r}; known: {sorted(VIT_CONFIGS)}")
    img, patch, dim, depth, heads, rep = VIT_CONFIGS[model_type]
    return VisionTransformer(img, patch, 3, dim, depth, heads, 4.0, rep)


class TIMMModel(nn.Module):
    """Same constructor / forward contract as models.video_classification.TIMMModel (:249-257)."""

    def __init__(self, num_classes, model_type="vit_base_patch16_224", drop_cls=True):
        super().__init__()
        if not drop_cls:
            raise NotImplementedError("drop_cls=False (ImageNet head) is not used by any AVT config")
        del num_classes
        self.model = create_model(model_type, num_classes=0)

    def forward(self, video):
        """video (B, C, T, H, W) -> (B, C', T, 1, 1)  [process_each_frame, video_classification.py:213-227]"""
        B, T = video.size(0), video.size(2)
        flat = video.transpose(1, 2).flatten(0, 1)
        feats = self.model(flat)
        return feats.view((B, T) + feats.shape[1:]).transpose(1, 2).unsqueeze(-1).unsqueeze(-1)
