"""ctypes binding of the C-ABI in include/avt_b200.h (the only way Python reaches the CUDA kernels).

The shared library is built in-tree by `avt_b200.build` (nvcc, sm_100a). There is no fallback: if the
library is missing or a call fails, a RuntimeError carrying avt_last_error() is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AVT_B200_LIB: load another build of the same C-ABI instead (kernel A/B experiments, tools/build_variant.sh)
LIB_PATH = os.environ.get("AVT_B200_LIB") or os.path.join(_HERE, "libavt_b200.so")

ACT_NONE, ACT_GELU_ERF, ACT_GELU_TANH = 0, 1, 2


class Epilogue(C.Structure):
    """Mirror of avt_epilogue_t."""
    _fields_ = [
        ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("ldr", C.c_int64),
        ("dact_z", C.c_void_p),
        ("aux_z", C.c_void_p),
        ("ldz", C.c_int64),
        ("pos", C.c_void_p),
        ("cls", C.c_void_p),
        ("pos_period", C.c_int32),
        ("act", C.c_int32),
        ("dact", C.c_int32),
        ("aux_mode", C.c_int32),
        ("dact_mode", C.c_int32),
        ("alpha", C.c_float),
        ("drop_p", C.c_float),
        ("drop_seed", C.c_uint64),
        ("drop_offset", C.c_uint64),
        ("drop_offset_dev", C.c_void_p),
        ("out", C.c_void_p),
        ("ldo", C.c_int64),
        ("out_fp32", C.c_int32),
        ("accumulate", C.c_int32),
    ]


_lib = None

_i64, _i32, _f32, _u64, _vp = C.c_int64, C.c_int, C.c_float, C.c_uint64, C.c_void_p

# name -> argtypes; every function returns int (0 = ok) unless listed in _SPECIAL.
SIGNATURES = {
    "avt_check_device": [],
    "avt_set_sm_limit": [_i32],
    "avt_set_pdl": [_i32],
    "avt_set_gemm_specialized_epilogues": [_i32],
    "avt_gemm_bf16": [_vp, _i64, _i32, _vp, _i64, _i32, _i64, _i64, _i64, C.POINTER(Epilogue), _i32, _i32, _i32, _vp, _i64,
                      _vp],
    "avt_gemm_bf16_colsum": [_vp, _i64, _i32, _vp, _i64, _i32, _i64, _i64, _i64, C.POINTER(Epilogue), _i32, _i32, _i32, _vp,
                             _i64, _vp, _vp],
    "avt_layernorm_fwd": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _f32, _i64, _i32, _vp, _i32, _i64, _vp, _vp, _vp],
    "avt_layernorm_bwd": [_vp, _i32, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _i64, _vp, _vp,
                          _vp, _i32, _vp, _i64, _vp],
    "avt_cast_f32_to_bf16": [_vp, _vp, _i64, _vp],
    "avt_zero": [_vp, _i64, _vp],
    "avt_patchify_bf16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "avt_colsum_bf16": [_vp, _i64, _i32, _i64, _vp, _vp],
    "avt_frame_sum_grads": [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp],
    "avt_dropout_apply": [_vp, _i64, _f32, _u64, _u64, _vp, _vp, _vp, _vp],
    "avt_preprocess_u8": [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _f32, C.POINTER(C.c_float), C.POINTER(C.c_float), _vp],
    "avt_sgemm_f32": [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _i64, _vp],
    "avt_attention_f32_fwd": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "avt_patchify_f32": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "avt_softmax_xent": [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp],
    "avt_sgd_step": [_vp, _vp, _i32, _vp, _vp, _i64, _f32, _vp, _f32, _f32, _f32, _i64, _i32, _i32, _vp],
    "avt_attention_simt_fwd": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _u64, _u64, _vp, _vp],
    "avt_attention_simt_decode": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "avt_attention_tc_fwd": [_vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp],
    "avt_attention_tc_bwd": [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp],
    "avt_attention_simt_bwd": [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _u64, _u64, _vp, _vp],
}
_SPECIAL = {"avt_abi_version": ([], C.c_int), "avt_last_error": ([], C.c_char_p), "avt_kernel_launch_count": ([], C.c_longlong),
            "avt_layernorm_bwd_workspace_bytes": ([_i64, _i32], C.c_int64)}


def lib():
    """Load (once) and return the ctypes handle; raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m avt_b200.build` (nvcc, sm_100a). "
                "avt_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        for name, (argtypes, restype) in _SPECIAL.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().avt_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"avt_b200 {what} failed (code {rc}): {msg}")


# kernels launched by the library so far (for bench.py's `gpu_launches` claim): the C side counts every launch it makes
launch_count = 0


def call(name, *args):
    global launch_count
    h = lib()
    check(getattr(h, name)(*args), name)
    launch_count = h.avt_kernel_launch_count()
