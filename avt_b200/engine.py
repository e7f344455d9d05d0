"""Host-side sequencing of the CUDA kernels for the two transformer stacks on the AVT hot path.

`ParamPack`   flat fp32 master / bf16 shadow / fp32 gradient buffers (one allocation each, 256-byte aligned
              slices) so that weights are down-cast by one kernel and gradients all-reduce as one message.
`BlockStack`  pre-LN transformer blocks, forward + backward, shared by
              * AVT-b  = timm VisionTransformer blocks (nn.Linear [out,in] weights, erf-GELU, LN eps 1e-6,
                         full attention over 197 tokens)                       [SURVEY.md §3.3]
              * AVT-h  = HF GPT2Block (Conv1D [in,out] weights, gelu_new, LN eps 1e-5, causal attention,
                         attn/resid dropout)                                    [SURVEY.md §3.3b]
torch is used here only to own device memory and streams; every arithmetic step is an `ops.*` call into the
C-ABI. No autograd, no torch math.
"""
import math
import weakref
from dataclasses import dataclass

import torch

from . import ops

_ALIGN = 64  # elements: 256 B in fp32, 128 B in bf16 (TMA needs 16 B; swizzled tiles like 128 B)
import os as _os
_FUSE_RESIDUAL = _os.environ.get("AVT_FUSE_RES", "1") != "0"   # A/B switch: 0 = residual adds stay in the LayerNorm kernels


class ParamPack:
    """Owns flat buffers for a fixed list of named parameters and re-points the nn.Parameters into them."""

    def __init__(self, named_params, device):
        named_params = list(named_params)
        # vectors (biases, LayerNorm) first: their gradients are produced by atomics and need zero-initialisation
        # each backward; matrices follow in module order (one layer's matrices are contiguous, so a layer's weight
        # gradients can be all-reduced as one slice while the backward of earlier layers is still running).
        # Within the vectors the biases come first: the reference puts every parameter whose name ends in 'bias' into its
        # own optimizer group with weight decay x opt.bias_bn_wd_scale (func/train.py:704-731), so [0, bias_end) and
        # [bias_end, total) are the two weight-decay regions of the fused SGD.
        small = [(n, p) for n, p in named_params if p.dim() < 2]
        small = [(n, p) for n, p in small if n.endswith("bias")] + [(n, p) for n, p in small if not n.endswith("bias")]
        big = [(n, p) for n, p in named_params if p.dim() >= 2]
        self.names, self.slices, self.shapes = [], {}, {}
        off = 0
        self.bias_end = self.small_end = 0
        for n, p in small + big:
            if p.dim() >= 2 and off == self.small_end:
                # the vectors end on a 512-element boundary and so does the buffer: the matrix region [small_end, total)
                # divides into whole 64-element slices for up to 8 ranks (reduce-scatter / sharded optimizer)
                off = self.small_end = (off + 511) // 512 * 512
            self.names.append(n)
            self.slices[n] = (off, p.numel())
            self.shapes[n] = tuple(p.shape)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            if p.dim() < 2:
                self.small_end = off
                if n.endswith("bias"):
                    self.bias_end = off
        off = (off + 511) // 512 * 512
        self.total = off
        self.device = device
        self.w = torch.zeros(off, dtype=torch.float32, device=device)
        self.g = torch.zeros(off, dtype=torch.float32, device=device)
        self.b = torch.zeros(off, dtype=torch.bfloat16, device=device)
        self.gb = None                 # bf16 gradient buffer = the data-parallel payload (enable_bf16_grads)
        self.matrix_grads_bf16 = False  # weight-gradient GEMMs write straight into gb (no split-K accumulation: AVT-h)
        self.params = {}
        with torch.no_grad():
            for n, p in small + big:
                view = self.wv(n)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
                self.params[n] = p
        self._ptrs = {n: p.data_ptr() for n, p in self.params.items()}
        self._sig = None
        self.refresh_bf16()

    def _view(self, buf, n):
        o, k = self.slices[n]
        return buf[o:o + k].view(self.shapes[n])

    def wv(self, n):
        return self._view(self.w, n)

    def gv(self, n):
        return self._view(self.g, n)

    def bv(self, n):
        return self._view(self.b, n)

    def enable_bf16_grads(self, matrices_direct):
        """Data-parallel mode: gradients leave the GPU as bf16 (792 MB instead of 1.585 GB per step).
        matrices_direct: the weight-gradient GEMMs store bf16 straight into `gb` (possible when they do not accumulate
        split-K partial sums in fp32 atomics, i.e. for the 80-row AVT-h); otherwise `g` is down-cast slice by slice."""
        if self.gb is None:
            self.gb = torch.zeros(self.total, dtype=torch.bfloat16, device=self.device)
        self.matrix_grads_bf16 = bool(matrices_direct)

    def fp32_grad_ranges(self):
        """Element ranges of the MATRIX region whose gradients are produced in fp32 `g` even with bf16 matrix gradients: the
        matrices listed in `fp32_grad_names` (e.g. the position-embedding table, summed over the batch). (The vectors,
        [0, small_end), always stay fp32: their gradients come from atomics and the kernels read them from the master.)"""
        out = []
        for n in getattr(self, "fp32_grad_names", ()):
            o, k = self.slices[n]
            out.append((o, o + (k + _ALIGN - 1) // _ALIGN * _ALIGN))
        return out

    def grad_out(self, n):
        """Where a weight-gradient GEMM writes the gradient of matrix `n`."""
        direct = self.matrix_grads_bf16 and n not in getattr(self, "fp32_grad_names", ())
        return self._view(self.gb, n) if direct else self._view(self.g, n)

    def intact(self):
        """False if someone re-allocated a parameter (e.g. module.to()) so the flat views are stale."""
        return all(p.data_ptr() == self._ptrs[n] and p.device == self.w.device for n, p in self.params.items())

    def _signature(self):
        # in-place writes to the flat buffer itself (a stock optimizer stepping on flat_parameter_groups(), a broadcast)
        # bump only w._version, writes through the parameter views only the parameters'
        return self.w._version + sum(p._version for p in self.params.values())

    def refresh_bf16(self):
        """Down-cast the master weights unless nothing touched them since the shadow was last written (torch
        in-place ops bump the parameters' version counters; avt_sgd_step refreshes the shadow itself)."""
        sig = self._signature()
        if sig != self._sig:
            ops.cast_bf16(self.w, self.b)
            self._sig = sig

    def shadow_is_current(self):
        self._sig = self._signature()

    def zero_small_grads(self):
        if self.small_end:
            ops.zero_(self.g[:self.small_end])

    def zero_all_grads(self):
        ops.zero_(self.g)

    def attach_grads(self):
        """Direct-gradient mode: make every param.grad a view of the flat gradient buffer (fp32; with bf16 matrix
        gradients the matrices' .grad stays None - the fused optimizer reads `gb`)."""
        for n, p in self.params.items():
            p.grad = None if (self.matrix_grads_bf16 and p.dim() >= 2) else self.gv(n)


def param_grads(pack, names, params, direct):
    """What an autograd.Function returns for its parameter inputs after the kernels filled pack.g.
    direct=True : gradients live in pack.g and param.grad are views of it (re-attached if the optimizer
                  set them to None); autograd gets None. One backward per step (grads are overwritten).
    direct=False: standard autograd semantics (accumulation, DDP hooks): views of ONE cloned buffer."""
    if direct:
        for n, p in zip(names, params):
            if p.requires_grad and p.grad is None and not (pack.matrix_grads_bf16 and p.dim() >= 2):
                p.grad = pack.gv(n)
        return (None,) * len(names)
    assert not pack.matrix_grads_bf16, "bf16 matrix gradients need direct_grads (FlatDataParallel + FlatSGD)"
    g = pack.g.clone()
    out = []
    for n, p in zip(names, params):
        o, k = pack.slices[n]
        out.append(g[o:o + k].view(pack.shapes[n]) if p.requires_grad else None)
    return tuple(out)


@dataclass
class StackSpec:
    dim: int
    heads: int
    layers: int
    eps: float
    act: int                 # ops.ACT_GELU_ERF | ops.ACT_GELU_TANH
    conv1d: bool             # weights stored [in, out] (HF Conv1D) instead of [out, in] (nn.Linear)
    causal: bool
    names: dict              # keys: ln1, qkv, proj, ln2, fc1, fc2 -> format strings with {i}
    p_attn: float = 0.0
    p_resid: float = 0.0
    attn_impl: str = "simt"  # "simt" | "tc" (tcgen05 kernel for N=197, hd=64)


def _best_split(tiles, kblocks, slots, max_split, unit_overhead, allowed=None):
    """Split-K factor minimising  waves x (k-blocks per unit + per-unit overhead): a persistent grid of `slots` CTAs (or
    CTA pairs) runs ceil(units / slots) rounds, so 160 units on 148 SMs cost two rounds, not 1.08. `allowed`: the effective
    split factors to choose from."""
    best_cost, best = None, 1
    for sk in range(1, max(1, max_split) + 1):
        kb = -(-kblocks // sk)
        eff = -(-kblocks // kb)                 # the C-ABI drops empty splits
        if allowed is not None and eff not in allowed:
            continue
        waves = -(-tiles * eff // slots)
        cost = waves * (kb + unit_overhead)
        if best_cost is None or cost < best_cost:
            best_cost, best = cost, eff
    return best


def small_m_split(M, N, K, sms=148):
    """Split-K factor for the weight-streaming GEMMs (M <= 128 rows: one 128 x 64 tile per CTA would leave most SMs
    idle while a few CTAs stream the whole weight matrix). 1 = no split."""
    if M > 128:
        return 1
    bn = ops.small_m_block_n(N)
    tiles = (N + bn - 1) // bn
    kblocks = (K + 63) // 64
    # 1 / 2 / 4: the split units of a tile run as one thread-block cluster and reduce through distributed shared memory
    # inside the GEMM kernel (no fp32 slices in HBM, no finishing launch). Clusters of 8 also work, but 16 of them do not
    # become co-resident on the 148 SMs (measured: proj 9.7 us with 4 slabs, 17.0 us with 8).
    return _best_split(tiles, kblocks, sms, min(16, kblocks // 2), 6, allowed=(1, 2, 4))


def _split_k_for(m_w, n_w, k_rows, block_n, sms=148):
    """Split-K factor for a weight gradient dW[m_w, n_w] contracted over k_rows activations rows (256 x block_n tiles on
    CTA pairs; the partial sums meet in fp32 atomics, so fewer splits win ties)."""
    tiles = ((m_w + 255) // 256) * ((n_w + block_n - 1) // block_n)
    kblocks = (k_rows + 63) // 64
    return _best_split(tiles, kblocks, sms // 2, min(32, kblocks // 8), 8)


class Lease:
    """Ownership of one activation workspace by the forward whose backward still has to read it. Lives on that forward's
    autograd ctx: the workspace is handed out again once the backward has run (`done`) or the graph was dropped (the
    weak reference died). `gen` detects a backward over activations a later forward has overwritten."""
    __slots__ = ("gen", "done", "__weakref__")


class BlockStack:
    MAX_LIVE_WORKSPACES = 8   # forwards of one shape whose backward is still pending (multi-crop, gradient accumulation)

    def __init__(self, spec: StackSpec, pack: ParamPack):
        self.s, self.pack = spec, pack
        self.ws = {}
        self.grads_prezeroed = False  # the owner zeroes pack.g before backward (split-K wgrads then just accumulate)
        self.layer_done_hook = None   # called with the layer index once that layer's parameter gradients are final

    def layer_grad_range(self, i):
        """[start, end) element range of layer i's weight matrices in the flat buffers (they are packed contiguously)."""
        names = [v.format(i=i) + ".weight" for k, v in self.s.names.items() if k in ("qkv", "proj", "fc1", "fc2")]
        lo = min(self.pack.slices[n][0] for n in names)
        hi = max(self.pack.slices[n][0] + self.pack.slices[n][1] for n in names)
        return lo, hi

    # ------------------------------------------------------------------ linear helpers (both weight layouts)
    def _fwd(self, x, wname, out, **ep):
        sk = small_m_split(out.shape[0], out.shape[1], x.shape[1])
        ops.gemm(x, self.pack.bv(wname), out, b_mn=self.s.conv1d, split_k=sk, workspace=self._gemm_ws(out, sk), **ep)

    def _dgrad(self, dy, wname, out, **ep):
        sk = small_m_split(out.shape[0], out.shape[1], dy.shape[1])
        ops.gemm(dy, self.pack.bv(wname), out, b_mn=not self.s.conv1d, split_k=sk, workspace=self._gemm_ws(out, sk), **ep)

    def _gemm_ws(self, out, sk):
        if sk <= 1:
            return None
        n = sk * out.shape[0] * out.shape[1]
        if getattr(self, "_ws_buf", None) is None or self._ws_buf.numel() < n:
            self._ws_buf = torch.empty(n, dtype=torch.float32, device=out.device)
        return self._ws_buf

    def _wgrad(self, x, dy, wname, bname):
        dW = self.pack.grad_out(wname)
        rows = x.shape[0]
        if self.s.conv1d:   # dW[in, out] = X^T dY
            sk = _split_k_for(x.shape[1], dy.shape[1], rows, 256)
            assert sk == 1 or dW.dtype == torch.float32
            ops.gemm(x, dy, dW, a_mn=True, b_mn=True, split_k=sk, accumulate=sk > 1 and self.grads_prezeroed)
        else:               # dW[out, in] = dY^T X; the bias gradient (column sums of dY) rides on the A tiles in smem
            sk = _split_k_for(dy.shape[1], x.shape[1], rows, 256)
            assert dW.dtype == torch.float32
            ops.gemm(dy, x, dW, a_mn=True, b_mn=True, split_k=sk, accumulate=sk > 1 and self.grads_prezeroed,
                     a_colsum=self.pack.gv(bname) if bname is not None else None)
            return
        if bname is not None:
            ops.colsum(dy, self.pack.gv(bname))

    # ------------------------------------------------------------------ workspace
    def _workspace(self, M, nb, ntok, train):
        key = (M, nb, ntok, train)
        pool = self.ws.setdefault(key, [])
        for w in pool:
            lease = w["lease"]() if w["lease"] is not None else None
            if lease is None or lease.done:
                return w
        if len(pool) >= self.MAX_LIVE_WORKSPACES:
            raise RuntimeError(f"{len(pool)} forwards of shape {key} are waiting for their backward: every one of them keeps "
                               "a full set of activations alive (run backward, or drop the graphs, before the next forward)")
        w = self._alloc_workspace(M, nb, ntok, train)
        pool.append(w)
        return w

    def lease(self, w):
        """Mark `w` as owned by the forward that just filled it (train mode); keep the returned object on the autograd ctx."""
        lease = Lease()
        w["gen"] += 1
        lease.gen, lease.done = w["gen"], False
        w["lease"] = weakref.ref(lease)
        return lease

    @staticmethod
    def check_lease(w, lease):
        if lease.gen != w["gen"]:
            raise RuntimeError("the activations saved by this forward were overwritten by a later forward of the same shape "
                               "(backward called twice with retain_graph after another forward ran)")

    def _alloc_workspace(self, M, nb, ntok, train):
        s, dev = self.s, self.pack.device
        D, L = s.dim, s.layers
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda *shape, dt=bf: torch.empty(*shape, dtype=dt, device=dev)
        nl = L if train else 1  # without grad only one layer's activations are live
        w = {
            "x": [e(M, D, dt=f32) for _ in range(2 * nl + 1)],   # x_in, x_mid per layer, final
            "ln1": [e(M, D) for _ in range(nl)], "qkv": [e(M, 3 * D) for _ in range(nl)],
            "att": [e(M, D) for _ in range(nl)], "lse": [e(nb * s.heads, ntok, dt=f32) for _ in range(nl)],
            "ln2": [e(M, D) for _ in range(nl)], "z": [e(M, 4 * D) for _ in range(nl)],
            "h": [e(M, 4 * D) for _ in range(nl)],
            "st": [e(4, M, dt=f32) for _ in range(nl)],           # mean1, rstd1, mean2, rstd2
            "y": e(M, D),                                          # branch output (attention / MLP), bf16
        }
        if train:
            w.update({"dx": e(M, D, dt=f32), "dxb": e(M, D), "dz": e(M, 4 * D), "dln": e(M, D), "datt": e(M, D),
                      "dqkv": e(M, 3 * D), "g": e(M, D),
                      "lnws": torch.empty(ops.layernorm_bwd_workspace(M, D), dtype=torch.uint8, device=dev)})
        w.update({"lease": None, "gen": 0, "aux": None})   # aux: the owning module's own per-forward buffers
        return w

    # ------------------------------------------------------------------ forward
    def workspace(self, M, nb, ntok, train):
        """Activation buffers for M = nb*ntok rows. The caller writes the stack input (fp32 residual
        stream) into w["x"][0] before calling forward()."""
        return self._workspace(M, nb, ntok, train)

    @staticmethod
    def _xbuf(w, train, k):
        return w["x"][max(k, 0)] if train else w["x"][k % 3]

    def forward(self, w, nb, ntok, train, rng=(0, 0), dropout=False):
        """Runs all blocks on w["x"][0]. With residual dropout the last residual add is left pending so the caller can fuse it
        into its final LayerNorm: returns (x_mid_last fp32, y_last bf16) with stack output = x_mid_last + y_last; without it
        returns (stack output fp32, None).
        train: keep every layer's activations for backward. dropout: apply spec.p_attn / p_resid."""
        s, pk = self.s, self.pack
        D = s.dim
        hd = D // s.heads
        scale = hd ** -0.5
        seed, off, off_dev = (tuple(rng) + (None,))[:3]   # off_dev: device-resident part of the Philox offset (CUDA graphs)
        p_attn = s.p_attn if dropout else 0.0
        p_res = s.p_resid if dropout else 0.0
        y = w["y"]
        # Without residual dropout the branch-closing GEMMs (proj, fc2) add the residual themselves and write the fp32 stream
        # (x_mid = x_in + proj(...), x_next = x_mid + fc2(...)): 8 extra bytes per element on tensor-bound kernels, and the
        # HBM-bound LayerNorms only read x and write their bf16 output (6 instead of 12 bytes per element).
        fuse_res = p_res == 0.0
        fuse_res = fuse_res and _FUSE_RESIDUAL
        for i in range(s.layers):
            j = i if train else 0
            nm = {k: v.format(i=i) for k, v in s.names.items()}
            st = w["st"][j]
            xprev, xin, xmid = (self._xbuf(w, train, 2 * i + k) for k in (-1, 0, 1))
            g1, b1 = pk.wv(nm["ln1"] + ".weight"), pk.wv(nm["ln1"] + ".bias")
            if i == 0 or fuse_res:     # (fuse_res: the previous layer's fc2 wrote x_in itself)
                ops.layernorm_fwd(xin, g1, b1, s.eps, w["ln1"][j], st[0], st[1])
            else:   # x_in = x_mid(prev) + mlp branch(prev), fused with this layer's LN1
                ops.layernorm_fwd(xprev, g1, b1, s.eps, w["ln1"][j], st[0], st[1], add=y, x_out=xin)
            self._fwd(w["ln1"][j], nm["qkv"] + ".weight", w["qkv"][j], bias=pk.wv(nm["qkv"] + ".bias"))
            self._attn_fwd(w["qkv"][j], w["att"][j], w["lse"][j], nb, ntok, hd, scale, p_attn, seed, off + (4 * i << 28),
                           off_dev)
            if fuse_res:
                self._fwd(w["att"][j], nm["proj"] + ".weight", xmid, bias=pk.wv(nm["proj"] + ".bias"), residual=xin)
                ops.layernorm_fwd(xmid, pk.wv(nm["ln2"] + ".weight"), pk.wv(nm["ln2"] + ".bias"), s.eps, w["ln2"][j], st[2], st[3])
            else:
                self._fwd(w["att"][j], nm["proj"] + ".weight", y, bias=pk.wv(nm["proj"] + ".bias"),
                          drop_p=p_res, drop_seed=seed, drop_offset=off + ((4 * i + 1) << 28), drop_offset_dev=off_dev)
                ops.layernorm_fwd(xin, pk.wv(nm["ln2"] + ".weight"), pk.wv(nm["ln2"] + ".bias"), s.eps, w["ln2"][j], st[2], st[3],
                                  add=y, x_out=xmid)
            # fc1 epilogue writes gelu(z) and gelu'(z): backward only multiplies
            self._fwd(w["ln2"][j], nm["fc1"] + ".weight", w["h"][j], bias=pk.wv(nm["fc1"] + ".bias"), act=s.act,
                      aux_z=w["z"][j] if train else None, aux_grad=True)
            if fuse_res:
                self._fwd(w["h"][j], nm["fc2"] + ".weight", self._xbuf(w, train, 2 * i + 2), bias=pk.wv(nm["fc2"] + ".bias"),
                          residual=xmid)
            else:
                self._fwd(w["h"][j], nm["fc2"] + ".weight", y, bias=pk.wv(nm["fc2"] + ".bias"),
                          drop_p=p_res, drop_seed=seed, drop_offset=off + ((4 * i + 2) << 28), drop_offset_dev=off_dev)
        w["rng"] = (seed, off, off_dev, p_attn, p_res)
        w["dims"] = (nb, ntok)
        if fuse_res:
            return self._xbuf(w, train, 2 * s.layers), None      # the stack output itself, nothing pending
        return self._xbuf(w, train, 2 * s.layers - 1), y

    def _attn_fwd(self, qkv, out, lse, nb, ntok, hd, scale, p, seed, off, off_dev=None):
        s = self.s
        if s.attn_impl == "tc" and p == 0.0 and not s.causal and hd == 64 and ntok <= 208:
            ops.attention_tc_fwd(qkv, out, lse, nb, s.heads, ntok, scale=scale)
        else:
            ops.attention_simt_fwd(qkv, out, lse, nb, s.heads, ntok, hd, causal=s.causal, scale=scale, drop_p=p,
                                   seed=seed, offset=off, offset_dev=off_dev)

    def _attn_bwd(self, qkv, att, dout, lse, dqkv, nb, ntok, hd, scale, p, seed, off, off_dev=None):
        s = self.s
        if s.attn_impl == "tc" and p == 0.0 and not s.causal and hd == 64 and ntok <= 208:
            ops.attention_tc_bwd(qkv, att, dout, lse, dqkv, nb, s.heads, ntok, scale=scale)
        else:
            ops.attention_simt_bwd(qkv, att, dout, lse, dqkv, nb, s.heads, ntok, hd, causal=s.causal, scale=scale, drop_p=p,
                                   seed=seed, offset=off, offset_dev=off_dev)

    # ------------------------------------------------------------------ KV-cached decode (evaluation-time rollout)
    def decode_step(self, x, caches, att, nb, tmax, pos):
        """One new token per batch item through all blocks: x fp32 [nb, D] = its input embedding (updated in place to the
        stack output BEFORE the final LayerNorm, the pending MLP branch in `y`: returns (x_mid, y) like forward()).
        caches[l] bf16 [nb*tmax, 3D]: layer l's packed qkv rows of all earlier positions (the prefill pass wrote rows
        0..T-1); the new token's q/k/v go to row `pos` of every item and only that row attends (avt_attention_simt_decode).
        att bf16 [nb*tmax, D]: attention output scratch. The reference does the same with HF GPT2Model's past_key_values
        (models/future_prediction.py:168-202)."""
        s, pk = self.s, self.pack
        D = s.dim
        hd = D // s.heads
        dev = x.device
        bf = torch.bfloat16
        ln = torch.empty(nb, D, dtype=bf, device=dev)
        h = torch.empty(nb, 4 * D, dtype=bf, device=dev)
        y = torch.empty(nb, D, dtype=bf, device=dev)
        xmid = torch.empty_like(x)
        xin = x
        for i in range(s.layers):
            nm = {k: v.format(i=i) for k, v in s.names.items()}
            g1, b1 = pk.wv(nm["ln1"] + ".weight"), pk.wv(nm["ln1"] + ".bias")
            if i == 0:
                ops.layernorm_fwd(xin, g1, b1, s.eps, ln)
            else:       # x_in = x_mid(prev) + mlp branch(prev), fused with this layer's LN1
                ops.layernorm_fwd(xmid, g1, b1, s.eps, ln, add=y, x_out=xin)
            qkv_row = caches[i].view(nb, tmax, 3 * D)[:, pos]                 # [nb, 3D], row stride tmax * 3D
            self._fwd(ln, nm["qkv"] + ".weight", qkv_row, bias=pk.wv(nm["qkv"] + ".bias"))
            ops.attention_simt_decode(caches[i], att, nb, s.heads, tmax, hd, pos, scale=hd ** -0.5)
            self._fwd(att.view(nb, tmax, D)[:, pos], nm["proj"] + ".weight", y, bias=pk.wv(nm["proj"] + ".bias"))
            ops.layernorm_fwd(xin, pk.wv(nm["ln2"] + ".weight"), pk.wv(nm["ln2"] + ".bias"), s.eps, ln, add=y, x_out=xmid)
            self._fwd(ln, nm["fc1"] + ".weight", h, bias=pk.wv(nm["fc1"] + ".bias"), act=s.act)
            self._fwd(h, nm["fc2"] + ".weight", y, bias=pk.wv(nm["fc2"] + ".bias"))
        return xmid, y

    # ------------------------------------------------------------------ fp32-accuracy mode (inference)
    def forward_fp32(self, x, nb, ntok):
        """All blocks in fp32 on x fp32 [M, D] (updated in place and returned): LayerNorm -> qkv -> attention -> proj (+x) ->
        LayerNorm -> fc1 + GELU -> fc2 (+x), CUDA-core kernels of csrc/fp32_path.cu, weights read from the fp32 masters.
        The validation path behind the north star's 1e-5 (fp32) bound; no dropout, no activations kept (eval / no_grad)."""
        s, pk = self.s, self.pack
        M, D = x.shape
        hd = D // s.heads
        f32 = dict(dtype=torch.float32, device=x.device)
        ln, qkv, att, h = (torch.empty(M, D, **f32), torch.empty(M, 3 * D, **f32), torch.empty(M, D, **f32),
                           torch.empty(M, 4 * D, **f32))
        for i in range(s.layers):
            nm = {k: v.format(i=i) for k, v in s.names.items()}
            W = lambda k: pk.wv(nm[k] + ".weight")
            Bv = lambda k: pk.wv(nm[k] + ".bias")
            ops.layernorm_fwd(x, W("ln1"), Bv("ln1"), s.eps, ln)
            ops.sgemm_f32(ln, W("qkv"), qkv, b_kn=s.conv1d, bias=Bv("qkv"))
            ops.attention_f32_fwd(qkv, att, nb, s.heads, ntok, hd, causal=s.causal, scale=hd ** -0.5)
            ops.sgemm_f32(att, W("proj"), x, b_kn=s.conv1d, bias=Bv("proj"), residual=x)
            ops.layernorm_fwd(x, W("ln2"), Bv("ln2"), s.eps, ln)
            ops.sgemm_f32(ln, W("fc1"), h, b_kn=s.conv1d, bias=Bv("fc1"), act=s.act)
            ops.sgemm_f32(h, W("fc2"), x, b_kn=s.conv1d, bias=Bv("fc2"), residual=x)
        return x

    # ------------------------------------------------------------------ backward
    def backward(self, w, dx, dxb, top_bias_done=False):
        """dx fp32 [M, D] / dxb bf16 copy: gradient w.r.t. the stack output (updated in place to the gradient
        w.r.t. the stack input). Parameter gradients are written into pack.g: matrices are overwritten
        (or atomically accumulated by split-K units), biases and LayerNorm parameters are accumulated by atomics (the LayerNorm
        backward adds its column sums straight into them) - the caller zeroes the corresponding region of pack.g first
        (zero_all_grads / zero_small_grads). top_bias_done: the caller's own
        LayerNorm backward already produced the last layer's fc2 bias gradient (column sums of dx)."""
        s, pk = self.s, self.pack
        M, D = dx.shape
        nb, ntok = w["dims"]
        seed, off, off_dev, p_attn, p_res = w["rng"]
        hd = D // s.heads
        scale = hd ** -0.5
        # Without residual dropout the gradient of a branch's closing Linear output IS the residual-stream gradient, so
        # its bias gradient (column sums of dx) comes out of the LayerNorm backward that produced dx: no extra pass.
        fuse_bias = p_res == 0.0
        for i in reversed(range(s.layers)):
            nm = {k: v.format(i=i) for k, v in s.names.items()}
            st = w["st"][i]
            xin, xmid = w["x"][2 * i], w["x"][2 * i + 1]
            # ---- MLP branch: x_out = x_mid + drop(fc2(act(fc1(LN2(x_mid)))))
            g = dxb
            if p_res > 0.0:
                g = w["g"]
                ops.dropout_apply(dx, p_res, seed, off + ((4 * i + 2) << 28), y_bf16=g, offset_dev=off_dev)
            self._wgrad(w["h"][i], g, nm["fc2"] + ".weight",
                        None if fuse_bias and (i < s.layers - 1 or top_bias_done) else nm["fc2"] + ".bias")
            self._dgrad(g, nm["fc2"] + ".weight", w["dz"], dact_z=w["z"][i], dact=s.act, dact_is_grad=True)
            self._wgrad(w["ln2"][i], w["dz"], nm["fc1"] + ".weight", nm["fc1"] + ".bias")
            self._dgrad(w["dz"], nm["fc1"] + ".weight", w["dln"])
            ops.layernorm_bwd(w["dln"], xmid, st[2], st[3], pk.wv(nm["ln2"] + ".weight"), dx,
                              pk.gv(nm["ln2"] + ".weight"), pk.gv(nm["ln2"] + ".bias"), w["lnws"], dx_in=dx, dx_bf16=dxb,
                              dx_colsum=pk.gv(nm["proj"] + ".bias") if fuse_bias else None, accumulate=True)
            # ---- attention branch: x_mid = x_in + drop(proj(attn(qkv(LN1(x_in)))))
            g = dxb
            if p_res > 0.0:
                g = w["g"]
                ops.dropout_apply(dx, p_res, seed, off + ((4 * i + 1) << 28), y_bf16=g, offset_dev=off_dev)
            self._wgrad(w["att"][i], g, nm["proj"] + ".weight", None if fuse_bias else nm["proj"] + ".bias")
            self._dgrad(g, nm["proj"] + ".weight", w["datt"])
            self._attn_bwd(w["qkv"][i], w["att"][i], w["datt"], w["lse"][i], w["dqkv"], nb, ntok, hd, scale, p_attn, seed,
                           off + (4 * i << 28), off_dev)
            self._wgrad(w["ln1"][i], w["dqkv"], nm["qkv"] + ".weight", nm["qkv"] + ".bias")
            self._dgrad(w["dqkv"], nm["qkv"] + ".weight", w["dln"])
            ops.layernorm_bwd(w["dln"], xin, st[0], st[1], pk.wv(nm["ln1"] + ".weight"), dx,
                              pk.gv(nm["ln1"] + ".weight"), pk.gv(nm["ln1"] + ".bias"), w["lnws"], dx_in=dx, dx_bf16=dxb,
                              dx_colsum=pk.gv(s.names["fc2"].format(i=i - 1) + ".bias") if fuse_bias and i > 0 else None,
                              accumulate=True)
            if self.layer_done_hook is not None:
                self.layer_done_hook(i)
        return dx, dxb
