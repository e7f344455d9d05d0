"""Restatement of timm==0.4.12 `timm/models/vision_transformer.py` (VisionTransformer, Block, Attention, Mlp)
and `timm/models/layers/patch_embed.py` (PatchEmbed) in plain PyTorch.

timm is pinned by the reference (env.yaml:302) but is not vendored and not installed here, so this file
restates its published algorithm; the reference reaches it through `timm.create_model(model_type,
num_classes=0)` (models/video_classification.py:255-256). Parameter names/shapes follow timm so that
`jx_vit_base_*.pth` checkpoints load (func/train.py:669-688). TEST INFRASTRUCTURE.
"""
from collections import OrderedDict
from functools import partial

import torch
import torch.nn as nn

# model_type -> (img, patch, dim, depth, heads, representation_size)
CONFIGS = {
    "vit_base_patch16_224": (224, 16, 768, 12, 12, None),
    "vit_base_patch16_224_in21k": (224, 16, 768, 12, 12, None),  # 0.4.12 AugReg def: no pre_logits
    "vit_large_patch16_224": (224, 16, 1024, 24, 16, None),
    "vit_large_patch16_224_in21k": (224, 16, 1024, 24, 16, None),
    "vit_small_patch16_224": (224, 16, 384, 12, 6, None),
    # tiny configs used only by tests / golden fixtures
    "vit_test_patch16_32": (32, 16, 64, 2, 2, None),
    "vit_test_patch16_64": (64, 16, 128, 3, 2, None),
    "vit_test_patch16_32_prelogits": (32, 16, 64, 2, 2, 64),
}


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()  # erf GELU
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) * self.scale
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=0, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0, representation_size=None):
        super().__init__()
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        if representation_size:
            self.num_features = representation_size
            self.pre_logits = nn.Sequential(OrderedDict([("fc", nn.Linear(embed_dim, representation_size)),
                                                         ("act", nn.Tanh())]))
        else:
            self.pre_logits = nn.Identity()
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)

    def forward_features(self, x):
        x = self.patch_embed(x)
        cls = self.cls_token.expand(x.shape[0], -1, -1)
        x = torch.cat((cls, x), dim=1) + self.pos_embed
        x = self.blocks(x)
        x = self.norm(x)
        return self.pre_logits(x[:, 0])

    def forward(self, x):
        return self.head(self.forward_features(x))


def create_model(model_type, num_classes=0, **kw):
    """Stand-in for timm.create_model (models/video_classification.py:255)."""
    img, patch, dim, depth, heads, rep = CONFIGS[model_type]
    return VisionTransformer(img, patch, 3, num_classes, dim, depth, heads, 4.0, rep)
