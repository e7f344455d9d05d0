"""Tensor-level wrappers over the C-ABI (one Python function per entry point).

These take torch CUDA tensors only for their device pointers / shapes; no torch arithmetic happens here.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_GELU_ERF, ACT_GELU_TANH, ACT_NONE, Epilogue  # noqa: F401


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("avt_b200 ops need CUDA tensors (there is no CPU path)")


def small_m_block_n(N):
    """Tile width of the weight-streaming GEMMs (AVT-h, M = B*T <= 128 rows). Measured on B200 (tools/sweep.py head) with the
    cluster split-K reduction: N = 2048 outputs run best as 32 tiles of 64 columns x 4 k-slabs (128 CTAs in clusters of 4:
    8.9 / 17.7 us for K = 2048 / 8192 against 9.7 / 18.5 us with 128-wide tiles), the wide ones (N >= 4096) as 128-wide tiles
    x 2 k-slabs."""
    return 128 if N >= 4096 else 64


def gemm(a, b, out, *, a_mn=False, b_mn=False, bias=None, residual=None, act=ACT_NONE, aux_z=None, dact_z=None,
         dact=ACT_NONE, aux_grad=False, dact_is_grad=False, pos=None, cls=None, pos_period=0, alpha=1.0, drop_p=0.0,
         drop_seed=0, drop_offset=0, drop_offset_dev=None, accumulate=False, split_k=1, block_n=0, cta_group=0,
         workspace=None, a_colsum=None):
    """out[M,N] = epilogue(A[M,K] @ B[N,K]^T).  a_mn/b_mn: operand is stored transposed ([K,M] / [K,N]).
    a_colsum (fp32 [M], needs a_mn): += sum_k A[m, k], the bias gradient when A = dY^T of a weight-gradient GEMM."""
    _chk_cuda(a, b, out, bias, residual, aux_z, dact_z, pos, cls)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and out.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    assert tuple(out.shape) == (M, N), (out.shape, M, N)
    if block_n == 0 and M <= 128:
        block_n = small_m_block_n(N)
    ep = Epilogue()
    ep.bias = _ptr(bias)
    ep.residual = _ptr(residual)
    ep.ldr = residual.stride(0) if residual is not None else 0
    ep.dact_z = _ptr(dact_z)
    ep.aux_z = _ptr(aux_z)
    z = aux_z if aux_z is not None else dact_z
    ep.ldz = z.stride(0) if z is not None else 0
    ep.pos = _ptr(pos)
    ep.cls = _ptr(cls)
    ep.pos_period = pos_period
    ep.act = act
    ep.dact = dact
    ep.aux_mode = 1 if aux_grad else 0
    ep.dact_mode = 1 if dact_is_grad else 0
    ep.alpha = alpha
    ep.drop_p = drop_p
    ep.drop_seed = drop_seed
    ep.drop_offset = drop_offset
    ep.drop_offset_dev = _ptr(drop_offset_dev)
    ep.out = _ptr(out)
    ep.ldo = out.stride(0)
    ep.out_fp32 = 1 if out.dtype == torch.float32 else 0
    if not ep.out_fp32:
        assert out.dtype == torch.bfloat16
    ep.accumulate = 1 if accumulate else 0
    ws_bytes = 0 if workspace is None else workspace.numel() * workspace.element_size()
    if a_colsum is None:
        _lib.call("avt_gemm_bf16", _ptr(a), a.stride(0), int(a_mn), _ptr(b), b.stride(0), int(b_mn), M, N, K,
                  C.byref(ep), split_k, block_n, cta_group, _ptr(workspace), ws_bytes, _stream())
    else:
        _chk_cuda(a_colsum)
        assert a_mn and a_colsum.dtype == torch.float32 and a_colsum.numel() == M
        _lib.call("avt_gemm_bf16_colsum", _ptr(a), a.stride(0), int(a_mn), _ptr(b), b.stride(0), int(b_mn), M, N, K,
                  C.byref(ep), split_k, block_n, cta_group, _ptr(workspace), ws_bytes, _ptr(a_colsum), _stream())
    return out


def layernorm_fwd(x, gamma, beta, eps, y, mean=None, rstd=None, rows=None, x_stride=None, add=None, add_stride=None,
                  x_out=None, x_out_stride=None):
    """y = LN(x [+ add]) (x fp32 [rows, D] with row stride x_stride; add bf16 branch output; x_out fp32 receives
    x + add; y bf16 or fp32 [rows, D])."""
    _chk_cuda(x, gamma, beta, y, add, x_out)
    D = gamma.numel()
    rows = y.shape[0] if rows is None else rows
    x_stride = x.stride(0) if x_stride is None else x_stride
    if add is not None:
        assert add.dtype == torch.bfloat16
        add_stride = add.stride(0) if add_stride is None else add_stride
    if x_out is not None:
        assert x_out.dtype == torch.float32      # receives x + add (a plain copy of the normalised rows' input without add)
        x_out_stride = x_out.stride(0) if x_out_stride is None else x_out_stride
    _lib.call("avt_layernorm_fwd", _ptr(x), x_stride, _ptr(add), add_stride or 0, _ptr(x_out), x_out_stride or 0, _ptr(gamma),
              _ptr(beta), float(eps), rows, D, _ptr(y), int(y.dtype == torch.float32), y.stride(0), _ptr(mean), _ptr(rstd),
              _stream())
    return y


def layernorm_bwd_workspace(rows, D):
    return _lib.lib().avt_layernorm_bwd_workspace_bytes(rows, D)


def layernorm_bwd(dy, x, mean, rstd, gamma, dx_out, dgamma, dbeta, workspace, *, dx_in=None, dx_bf16=None, rows=None,
                  x_stride=None, dx_stride=None, dxb_stride=None, accumulate=False, dx_colsum=None):
    """dx_colsum (fp32 [D], optional) receives the column sums of dx_out (= bias gradient of the Linear that produced
    this residual stream update)."""
    _chk_cuda(dy, x, mean, rstd, gamma, dx_out, dgamma, dbeta, workspace, dx_colsum)
    D = gamma.numel()
    rows = dy.shape[0] if rows is None else rows
    x_stride = x.stride(0) if x_stride is None else x_stride
    dx_stride = dx_out.stride(0) if dx_stride is None else dx_stride
    if dx_bf16 is not None and dxb_stride is None:
        dxb_stride = dx_bf16.stride(0)
    _lib.call("avt_layernorm_bwd", _ptr(dy), int(dy.dtype == torch.float32), dy.stride(0), _ptr(x), x_stride, _ptr(mean),
              _ptr(rstd), _ptr(gamma), rows, D, _ptr(dx_in), _ptr(dx_out), dx_stride, _ptr(dx_bf16), dxb_stride or 0,
              _ptr(dgamma), _ptr(dbeta), _ptr(dx_colsum), int(accumulate), _ptr(workspace),
              workspace.numel() * workspace.element_size(),
              _stream())


def zero_(t):
    """t[...] = 0 as a stream-ordered memset (contiguous tensors)."""
    _chk_cuda(t)
    assert t.is_contiguous()
    _lib.call("avt_zero", _ptr(t), t.numel() * t.element_size(), _stream())
    return t


def cast_bf16(src, dst):
    _chk_cuda(src, dst)
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
    assert src.is_contiguous() and dst.is_contiguous()
    _lib.call("avt_cast_f32_to_bf16", _ptr(src), _ptr(dst), src.numel(), _stream())
    return dst


def patchify(video, out, patch):
    """video fp32 [F, C, H, W] (contiguous) -> out bf16 [F*(P+1), C*patch*patch]."""
    _chk_cuda(video, out)
    F, Cc, H, W = video.shape
    assert video.is_contiguous() and video.dtype == torch.float32 and out.dtype == torch.bfloat16
    _lib.call("avt_patchify_bf16", _ptr(video), _ptr(out), F, Cc, H, W, patch, _stream())
    return out


def colsum(x, out):
    """out[c] += sum_r x[r, c] (x bf16 2-D, out fp32)."""
    _chk_cuda(x, out)
    assert x.dtype == torch.bfloat16 and out.dtype == torch.float32 and x.stride(1) == 1
    _lib.call("avt_colsum_bf16", _ptr(x), x.shape[0], x.shape[1], x.stride(0), _ptr(out), _stream())


def frame_sum_grads(dx, F, period, D, workspace, dpos=None, dcls=None, dbias=None, accumulate=False):
    _chk_cuda(dx, workspace)
    assert dx.dtype == torch.float32 and workspace.numel() >= period * D
    _lib.call("avt_frame_sum_grads", _ptr(dx), F, period, D, _ptr(dpos), _ptr(dcls), _ptr(dbias), int(accumulate),
              _ptr(workspace), _stream())


def dropout_apply(x, p, seed, offset, y_f32=None, y_bf16=None, offset_dev=None):
    _chk_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    _lib.call("avt_dropout_apply", _ptr(x), x.numel(), float(p), int(seed), int(offset), _ptr(offset_dev), _ptr(y_f32),
              _ptr(y_bf16), _stream())


def attention_simt_fwd(qkv, out, lse, B, H, N, hd, *, causal, scale, drop_p=0.0, seed=0, offset=0, offset_dev=None):
    _chk_cuda(qkv, out, lse)
    assert qkv.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and qkv.is_contiguous() and out.is_contiguous()
    _lib.call("avt_attention_simt_fwd", _ptr(qkv), _ptr(out), _ptr(lse), B, H, N, hd, int(causal), float(scale),
              float(drop_p), int(seed), int(offset), _ptr(offset_dev), _stream())


def attention_simt_bwd(qkv, out, dout, lse, dqkv, B, H, N, hd, *, causal, scale, drop_p=0.0, seed=0, offset=0,
                       offset_dev=None):
    _chk_cuda(qkv, out, dout, lse, dqkv)
    assert dout.dtype == torch.bfloat16 and dqkv.dtype == torch.bfloat16 and dout.is_contiguous() and dqkv.is_contiguous()
    assert out.dtype == torch.bfloat16 and out.is_contiguous()
    _lib.call("avt_attention_simt_bwd", _ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv), B, H, N, hd, int(causal),
              float(scale), float(drop_p), int(seed), int(offset), _ptr(offset_dev), _stream())


def attention_tc_fwd(qkv, out, lse, F, H, N, *, scale):
    """tcgen05 attention forward (head_dim 64, N <= 208, no mask)."""
    _chk_cuda(qkv, out, lse)
    assert qkv.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and qkv.is_contiguous() and out.is_contiguous()
    assert qkv.shape == (F * N, 3 * H * 64)
    _lib.call("avt_attention_tc_fwd", _ptr(qkv), _ptr(out), _ptr(lse), F, H, N, float(scale), _stream())


def attention_tc_bwd(qkv, out, dout, lse, dqkv, F, H, N, *, scale):
    """tcgen05 attention backward (head_dim 64, N <= 208, no mask)."""
    _chk_cuda(qkv, out, dout, lse, dqkv)
    assert dout.dtype == torch.bfloat16 and dqkv.dtype == torch.bfloat16 and dout.is_contiguous() and dqkv.is_contiguous()
    assert out.is_contiguous() and qkv.is_contiguous()
    _lib.call("avt_attention_tc_bwd", _ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv), F, H, N, float(scale),
              _stream())


def sgd_step(p, g, m, shadow, lr, momentum, weight_decay, nesterov, first, *, weight_decay_lo=None, lo_elems=0, lr_dev=None):
    """p, m fp32 flat (or a contiguous shard of a flat buffer); g fp32 or bf16; elements [0, lo_elems) decay with
    weight_decay_lo; lr_dev: optional fp32 device scalar that overrides lr."""
    _chk_cuda(p, g, m, shadow, lr_dev)
    assert p.dtype == m.dtype == torch.float32 and p.numel() == g.numel() == m.numel()
    assert g.dtype in (torch.float32, torch.bfloat16) and p.is_contiguous() and g.is_contiguous() and m.is_contiguous()
    _lib.call("avt_sgd_step", _ptr(p), _ptr(g), int(g.dtype == torch.bfloat16), _ptr(m), _ptr(shadow), p.numel(), float(lr),
              _ptr(lr_dev), float(momentum), float(weight_decay),
              float(weight_decay if weight_decay_lo is None else weight_decay_lo), int(lo_elems), int(nesterov), int(first),
              _stream())


def softmax_xent(logits, classes, target, loss, *, row_scale=None, rank=None, dlogits=None):
    """logits fp32 [R, >= classes] (row stride may exceed `classes`), target int64 [R] (-1 = ignored), loss fp32 [R];
    optional rank int32 [R] and dlogits bf16 [R, classes_padded] = (softmax - onehot) * row_scale."""
    _chk_cuda(logits, target, loss, row_scale, rank, dlogits)
    assert logits.dtype == torch.float32 and logits.stride(1) == 1 and target.dtype == torch.int64 and target.is_contiguous()
    R = logits.shape[0]
    assert loss.numel() == R and target.numel() == R
    if dlogits is not None:
        assert dlogits.dtype == torch.bfloat16 and dlogits.stride(1) == 1 and row_scale is not None and row_scale.numel() == R
    _lib.call("avt_softmax_xent", _ptr(logits), logits.stride(0), R, int(classes), _ptr(target), _ptr(row_scale), _ptr(loss),
              _ptr(rank), _ptr(dlogits), dlogits.stride(0) if dlogits is not None else 0,
              dlogits.shape[1] if dlogits is not None else 0, _stream())


# ----------------------------------------------------------------------------- fp32-accuracy mode (inference)
def sgemm_f32(a, b, out, *, b_kn=False, bias=None, residual=None, act=ACT_NONE, pos=None, cls=None, pos_period=0):
    """out[M,N] = act(a[M,K] @ Bop + bias) + residual, all fp32; b is [N,K] (nn.Linear) or, with b_kn, [K,N] (HF Conv1D)."""
    _chk_cuda(a, b, out, bias, residual, pos, cls)
    assert a.dtype == b.dtype == out.dtype == torch.float32 and a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    M, K = a.shape
    N = b.shape[1] if b_kn else b.shape[0]
    assert (b.shape[0] if b_kn else b.shape[1]) == K and tuple(out.shape) == (M, N)
    _lib.call("avt_sgemm_f32", _ptr(a), a.stride(0), _ptr(b), b.stride(0), int(b_kn), M, N, K, _ptr(bias), _ptr(residual),
              residual.stride(0) if residual is not None else 0, act, _ptr(pos), _ptr(cls), pos_period, _ptr(out), out.stride(0),
              _stream())
    return out


def attention_f32_fwd(qkv, out, B, H, N, hd, *, causal, scale):
    _chk_cuda(qkv, out)
    assert qkv.dtype == out.dtype == torch.float32 and qkv.is_contiguous() and out.is_contiguous()
    _lib.call("avt_attention_f32_fwd", _ptr(qkv), _ptr(out), B, H, N, hd, int(causal), float(scale), _stream())


def patchify_f32(video, out, patch):
    _chk_cuda(video, out)
    F, Cc, H, W = video.shape
    assert video.is_contiguous() and video.dtype == out.dtype == torch.float32
    _lib.call("avt_patchify_f32", _ptr(video), _ptr(out), F, Cc, H, W, patch, _stream())


def preprocess_u8(frames, out, crop_y, crop_x, *, flip=None, scale=1.0, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """frames uint8 [F, Hin, Win, 3] (HWC, as decoded) -> out fp32 [F, 3, h, w]: /255, flip (uint8 [F] flags), * scale,
    normalise, crop - the tail of the reference's transform chain on the GPU (expts/01: mean = std = 0.5)."""
    _chk_cuda(frames, out, flip)
    assert frames.dtype == torch.uint8 and frames.is_contiguous() and frames.shape[-1] == 3 and out.dtype == torch.float32
    F, Hin, Win, _ = frames.shape
    assert out.is_contiguous() and out.shape[0] == F and out.shape[1] == 3
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    _lib.call("avt_preprocess_u8", _ptr(frames), F, Hin, Win, _ptr(out), out.shape[2], out.shape[3], int(crop_y), int(crop_x),
              _ptr(flip), float(scale), m3, s3, _stream())
    return out


def attention_simt_decode(qkv_cache, out, B, H, N, hd, q_row, *, scale):
    """KV-cached decode step: query row q_row of every batch item against keys 0..q_row of the packed qkv cache [B*N, 3*H*hd]."""
    _chk_cuda(qkv_cache, out)
    assert qkv_cache.dtype == out.dtype == torch.bfloat16 and qkv_cache.is_contiguous() and out.is_contiguous()
    _lib.call("avt_attention_simt_decode", _ptr(qkv_cache), _ptr(out), B, H, N, hd, int(q_row), float(scale), _stream())
