"""ctypes binding of libb200decode.so (the C ABI in include/b200_decode.h).

The library must have been built in-tree (``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C scalellm_b200/csrc``).  There is NO fallback: if the shared object
is missing, importing the kernels raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200decode.so")

B200_BF16, B200_FP16, B200_FP32 = 0, 1, 2
AR_HANDLE_BYTES = 128


class B200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200decode error {code}: {msg}")
        self.code = code


_i64, _i32, _f32, _vp, _int = C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_int

# name -> (restype, argtypes).  Kept in lock-step with include/b200_decode.h;
# tests/test_abi.py checks that every B200_API symbol in the header is listed here
# and exported by the shared object.
SIGNATURES = {
    "b200_abi_version": (_int, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_launch_count": (_i64, []),
    "b200_launch_count_reset": (None, []),
    "b200_rms_norm": (_int, [_vp, _vp, _vp, _i64, _i64, _f32, _int, _vp]),
    "b200_rms_norm_residual": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _int, _vp]),
    "b200_rope_inplace": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                                 _int, _int, _vp]),
    "b200_kv_write": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _int, _vp]),
    "b200_rope_kv_write": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64,
                                  _i64, _i64, _i64, _i64, _int, _int, _vp]),
    "b200_kv_gather": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "b200_silu": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "b200_silu_mul": (_int, [_vp, _vp, _i64, _i64, _int, _vp]),
    "b200_silu_mul_strided": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _vp]),
    "b200_paged_attn_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64, _i64, _i64]),
    "b200_paged_attn_decode": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,  # ptrs
                                      _i64, _i64, _i64, _i64, _i64,                # batch..n_slots
                                      _i64, _i64, _i64, _i64, _i64, _i64,          # strides
                                      _int, _int, _int, _f32, _f32, _int,          # bs..window
                                      _vp, _i64, _int, _vp]),
    "b200_w4a16_packed_bytes": (_i64, [_i64, _i64, _int]),
    "b200_w4a16_prepack_awq": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _int, _vp]),
    "b200_w4a16_prepack_gptq": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _vp]),
    "b200_w4a16_dequant": (_int, [_vp, _vp, _i64, _i64, _int, _vp]),
    "b200_w4a16_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "b200_w4a16_gemm": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _int, _vp, _i64,
                               _vp]),
    "b200_w4a16_splitk_splits": (_int, [_i64, _i64, _i64]),
    "b200_w4a16_gemm_splitk": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _vp]),
    "b200_w4a16_reduce_partials": (_int, [_vp, _vp, _int, _i64, _vp, _i64, _i64, _i64, _vp]),
    "b200_debug_attn_plan": (_int, [_i64, _int, _int, _int, _int, _int, _int, _vp]),
    "b200_debug_w4a16_plan": (_int, [_i64, _i64, _int, _int, _vp, _vp, _vp]),
    "b200_gemma_rms_norm": (_int, [_vp, _vp, _vp, _i64, _i64, _f32, _int, _vp]),
    "b200_layer_norm": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _int, _vp]),
    "b200_gelu": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _int, _int, _vp]),
    "b200_apply_temperature": (_int, [_vp, _vp, _i64, _i64, _int, _vp]),
    "b200_apply_repetition_penalty": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "b200_apply_frequency_presence_penalty": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "b200_softmax": (_int, [_vp, _i64, _i64, _int, _vp]),
    "b200_topk_topp_filter": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "b200_argmax": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "b200_rope_kv_write_splitk": (_int, [_vp, _vp, _int, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64,
                                         _i64, _i64, _int, _int, _vp]),
    "b200_silu_mul_splitk": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _int, _vp]),
    "b200_rms_norm_residual_splitk": (_int, [_vp, _vp, _vp, _int, _i64, _vp, _i64, _i64, _f32, _int, _vp]),
    "b200_debug_set_trace": (None, [_vp]),
    "b200_w4a16_prepack_gptq_actorder": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _vp]),
    "b200_w4a16_repack_awq": (_int, [_vp, _vp, _i64, _i64, _vp]),
    "b200_w4a16_repack_gptq": (_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "b200_w4a16_assemble_marlin": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _int, _vp]),
    "b200_permute_cols": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _vp]),
    "b200_dense_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "b200_dense_gemm": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _vp]),
    "b200_dense_splitk_splits": (_int, [_i64, _i64, _i64]),
    "b200_dense_gemm_splitk": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _int, _vp]),
    "b200_ar_create": (_int, [C.POINTER(_vp), _int, _int, _i64, _vp]),
    "b200_ar_open_peers": (_int, [_vp, _vp]),
    "b200_ar_create_all": (_int, [C.POINTER(_vp), C.POINTER(C.c_int), _int, _i64]),
    "b200_ar_allreduce": (_int, [_vp, _vp, _i64, _int, _vp]),
    "b200_ar_allgather": (_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "b200_ar_allreduce_splitk": (_int, [_vp, _vp, _vp, _int, _i64, _i64, _i64, _int, _vp]),
    "b200_ar_allreduce_splitk_norm": (_int, [_vp, _vp, _vp, _vp, _int, _i64, _vp, _i64, _i64, _f32, _int, _vp]),
    "b200_ar_argmax": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "b200_ar_destroy": (_int, [_vp]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load (once) and return the shared library; raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Build it with "
            "`python -c \"import __graft_entry__ as g; g.build()\"` (needs nvcc). "
            "There is no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise B200Error(rc, load().b200_last_error().decode("utf-8", "replace"))
