"""ctypes binding of libvitunet_b200.so (the C ABI declared in include/vit_unet_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or ``sh vit_unet_b200/csrc/build.sh``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VU_LIB_PATH") or os.path.join(_HERE, "libvitunet_b200.so")     # VU_LIB_PATH: A/B builds of the kernels
ABI_VERSION = 9


class VuError(RuntimeError):
    """A vu_* entry point returned non-zero (bad shape, unsupported config, CUDA error)."""


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("aux_in", C.c_void_p), ("aux_out", C.c_void_p),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("trans_a", C.c_int), ("trans_b", C.c_int),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64), ("ldr", C.c_int64), ("ldaux", C.c_int64),
        ("batch_outer", C.c_int), ("batch_inner", C.c_int),
        ("sAo", C.c_int64), ("sAi", C.c_int64), ("sBo", C.c_int64), ("sBi", C.c_int64),
        ("sCo", C.c_int64), ("sCi", C.c_int64),
        ("alpha", C.c_float), ("act", C.c_int), ("accumulate", C.c_int), ("split_k", C.c_int),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint64), ("drop_stream", C.c_uint32),
        ("precision", C.c_int),
        ("a_bf16", C.c_int), ("b_bf16", C.c_int), ("c_bf16", C.c_int), ("aux_bf16", C.c_int),
    ]


_p, _i, _l, _f, _u64, _u32 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_uint32

# name -> argtypes; every function returns int (status) unless listed in _SPECIAL
SIGNATURES = {
    "vu_repatch": [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    "vu_heads_transpose_bf16": [_p, _p, _i, _i, _i, _i, _i, _p],
    "vu_pe_fwd": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _p],
    "vu_pe_bwd_table": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p],
    "vu_conv3x3_fwd": [_p, _i, _p, _p, _p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "vu_conv3x3_bwd_data": [_p, _p, _p, _i, _p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "vu_conv3x3_bwd_weight": [_p, _i, _p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "vu_gemm": [C.POINTER(GemmDesc), _p],
    "vu_colsum": [_p, _i, _l, _i, _l, _p, _i, _p],
    "vu_softmax_rows": [_p, _l, _i, _i, _f, _p],
    "vu_softmax_stats": [_p, _p, _i, _i, _i, _i, _f, _f, _u64, _u32, _p, _i, _p],
    "vu_reattn_mix_reduce": [_p, _p, _p, _i, _p, _i, _i, _i, _i, _f, _u64, _u32, _p, _p],
    "vu_reattn_stats": [_p, _i, _i, _i, _i, _f, _u64, _u32, _p, _p],
    "vu_reattn_bn_finalize": [_p, _l, _i, _i, _p, _p, _p, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p],
    "vu_reattn_mix": [_p, _p, _i, _p, _i, _i, _i, _i, _f, _u64, _u32, _p],
    "vu_reattn_bwd_reduce": [_p, _p, _i, _i, _i, _i, _f, _u64, _u32, _p, _p],
    "vu_reattn_bwd_params": [_p, _p, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p],
    "vu_reattn_bwd_rows": [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _f, _f, _u64, _u32, _p],
    "vu_reattn_stream_fwd": [_i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _u64, _u32, _p],
    "vu_reattn_stream_bwd_reduce": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _u64, _u32, _p],
    "vu_reattn_stream_bwd_ds": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _u64, _u32, _p],
    "vu_ln_stats": [_p, _i, _l, _f, _p, _p, _p],
    "vu_ln_apply": [_p, _p, _p, _p, _p, _p, _i, _l, _p],
    "vu_ln_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _l, _p],
    "vu_loss_fwd": [_i, _p, _p, _l, _p, _p, _p],
    "vu_loss_finalize": [_i, _l, _p, _p, _p],
    "vu_loss_bwd": [_i, _p, _p, _l, _p, _p, _p, _p],
    "vu_psnr": [_p, _p, _i, _l, _f, _p, _p, _p],
    "vu_u8hwc_to_chw": [_p, _p, _i, _i, _i, _i, _f, _f, _f, _p],
    "vu_resize_u8hwc": [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    "vu_warp_u8hwc_to_chw": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _i, _f, _f, _f, _f, _p],
    "vu_dropout": [_p, _p, _i, _l, _f, _u64, _u32, _p],
    "vu_cast_bf16": [_p, _p, _p, _i, _i, _p],
    "vu_axpby": [_p, _p, _l, _f, _f, _p],
    "vu_zero": [_p, _l, _p],
    "vu_adamw": [_p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _i, _f, _p],
}
_SPECIAL = {
    "vu_version": ([], C.c_int),
    "vu_last_error": ([], C.c_char_p),
    "vu_device_sm_count": ([_i], C.c_int),
    "vu_reattn_tensor_core_path": ([_i, _i, _i], C.c_int),
    "vu_reattn_stream_supported": ([_i, _i, _i], C.c_int),
    "vu_gemm_tf32_fallbacks": ([], C.c_int),
}

_lib = None

# kernels launched per entry point (for bench.py's gpu_launches claim); memsets are not counted
_KERNELS_PER_CALL = {"vu_ln_bwd": 3, "vu_ln_stats": 2, "vu_loss_fwd": 2, "vu_psnr": 3, "vu_zero": 0}
LN_SPLIT = 8
_launches = 0


def reset_launch_count() -> None:
    global _launches
    _launches = 0


def launch_count() -> int:
    return _launches


def load() -> C.CDLL:
    """Load the shared library once; raise loudly if it is absent or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"vit_unet_b200: CUDA extension not built ({LIB_PATH} missing). "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (args, res) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, res
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.argtypes, fn.restype = args, C.c_int
    v = lib.vu_version()
    if v != ABI_VERSION:
        raise ImportError(f"vit_unet_b200: ABI mismatch (library {v}, binding {ABI_VERSION}); rebuild")
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    global _launches
    lib = load()
    rc = getattr(lib, name)(*args)
    _launches += _KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        msg = lib.vu_last_error()
        raise VuError(f"{name} failed (code {rc}): {msg.decode() if msg else '?'}")
