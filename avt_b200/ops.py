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


def gemm(a, b, out, *, a_mn=False, b_mn=False, bias=None, residual=None, act=ACT_NONE, aux_z=None, dact_z=None,
         dact=ACT_NONE, pos=None, cls=None, pos_period=0, alpha=1.0, drop_p=0.0, drop_seed=0, drop_offset=0,
         accumulate=False, split_k=1, block_n=0):
    """out[M,N] = epilogue(A[M,K] @ B[N,K]^T).  a_mn/b_mn: operand is stored transposed ([K,M] / [K,N])."""
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
    ep.alpha = alpha
    ep.drop_p = drop_p
    ep.drop_seed = drop_seed
    ep.drop_offset = drop_offset
    ep.out = _ptr(out)
    ep.ldo = out.stride(0)
    ep.out_fp32 = 1 if out.dtype == torch.float32 else 0
    if not ep.out_fp32:
        assert out.dtype == torch.bfloat16
    ep.accumulate = 1 if accumulate else 0
    _lib.call("avt_gemm_bf16", _ptr(a), a.stride(0), int(a_mn), _ptr(b), b.stride(0), int(b_mn), M, N, K,
              C.byref(ep), split_k, block_n, _stream())
    return out
