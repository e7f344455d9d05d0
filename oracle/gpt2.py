"""Restatement of transformers==4.2.2 `GPT2Model` (modeling_gpt2.py: Attention._attn, MLP, Block, GPT2Model)
restricted to what AVTh uses: `inputs_embeds` + `position_ids`, optional KV cache (`past_key_values`), no
`wte`, no attention/head masks. transformers is pinned by the reference (env.yaml:183), not vendored; the
reference calls it at models/future_prediction.py:89-95,178-182. Parameter names/shapes follow HF
(`Conv1D.weight` is [in, out]). TEST INFRASTRUCTURE.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class Conv1D(nn.Module):
    """HF Conv1D: y = x @ W + b with W stored [in, out]."""

    def __init__(self, nf, nx):
        super().__init__()
        self.nf = nf
        self.weight = nn.Parameter(torch.empty(nx, nf).normal_(std=0.02))
        self.bias = nn.Parameter(torch.zeros(nf))

    def forward(self, x):
        return torch.addmm(self.bias, x.reshape(-1, x.size(-1)), self.weight).view(x.shape[:-1] + (self.nf,))


def gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


class Attention(nn.Module):
    def __init__(self, nx, n_head, attn_pdrop, resid_pdrop):
        super().__init__()
        self.n_head, self.split_size = n_head, nx
        self.c_attn = Conv1D(3 * nx, nx)
        self.c_proj = Conv1D(nx, nx)
        self.attn_dropout = nn.Dropout(attn_pdrop)
        self.resid_dropout = nn.Dropout(resid_pdrop)

    def _heads(self, x):
        return x.view(x.shape[:-1] + (self.n_head, x.size(-1) // self.n_head)).permute(0, 2, 1, 3)

    def forward(self, x, layer_past=None):
        q, k, v = self.c_attn(x).split(self.split_size, dim=2)
        q, k, v = self._heads(q), self._heads(k), self._heads(v)
        if layer_past is not None:
            k = torch.cat((layer_past[0], k), dim=-2)
            v = torch.cat((layer_past[1], v), dim=-2)
        present = (k, v)
        w = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(v.size(-1))
        nd, ns = w.size(-2), w.size(-1)
        mask = torch.tril(torch.ones(ns, ns, dtype=torch.bool, device=w.device))[ns - nd:ns, :ns]
        w = torch.where(mask, w, torch.full_like(w, -1e4))  # 4.2.2: masked_bias = -1e4
        w = self.attn_dropout(F.softmax(w, dim=-1))
        a = torch.matmul(w, v).permute(0, 2, 1, 3).contiguous()
        a = a.view(a.shape[:-2] + (a.size(-2) * a.size(-1),))
        return self.resid_dropout(self.c_proj(a)), present


class MLP(nn.Module):
    def __init__(self, n_inner, nx, resid_pdrop):
        super().__init__()
        self.c_fc = Conv1D(n_inner, nx)
        self.c_proj = Conv1D(nx, n_inner)
        self.dropout = nn.Dropout(resid_pdrop)

    def forward(self, x):
        return self.dropout(self.c_proj(gelu_new(self.c_fc(x))))


class Block(nn.Module):
    def __init__(self, nx, n_head, eps, attn_pdrop, resid_pdrop):
        super().__init__()
        self.ln_1 = nn.LayerNorm(nx, eps=eps)
        self.attn = Attention(nx, n_head, attn_pdrop, resid_pdrop)
        self.ln_2 = nn.LayerNorm(nx, eps=eps)
        self.mlp = MLP(4 * nx, nx, resid_pdrop)

    def forward(self, x, layer_past=None):
        a, present = self.attn(self.ln_1(x), layer_past)
        x = x + a
        x = x + self.mlp(self.ln_2(x))
        return x, present


class GPT2Model(nn.Module):
    def __init__(self, n_embd=768, n_layer=12, n_head=12, n_positions=1024, layer_norm_epsilon=1e-5, embd_pdrop=0.1,
                 attn_pdrop=0.1, resid_pdrop=0.1, **unused):
        super().__init__()
        self.wpe = nn.Embedding(n_positions, n_embd)
        nn.init.normal_(self.wpe.weight, std=0.02)
        self.drop = nn.Dropout(embd_pdrop)
        self.h = nn.ModuleList(
            [Block(n_embd, n_head, layer_norm_epsilon, attn_pdrop, resid_pdrop) for _ in range(n_layer)])
        self.ln_f = nn.LayerNorm(n_embd, eps=layer_norm_epsilon)

    def forward(self, inputs_embeds, past_key_values=None, position_ids=None):
        """Returns (last_hidden_state, presents)."""
        if past_key_values is None:
            past_key_values = [None] * len(self.h)
        x = self.drop(inputs_embeds + self.wpe(position_ids))
        presents = []
        for blk, past in zip(self.h, past_key_values):
            x, present = blk(x, past)
            presents.append(present)
        return self.ln_f(x), presents
