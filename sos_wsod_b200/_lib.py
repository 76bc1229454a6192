"""ctypes binding of libsoswsod_b200.so (C ABI: include/soswsod_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised
(BASELINE.json north_star: "no CPU fallback").  Build it with ``python -m sos_wsod_b200.build``."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsoswsod_b200.so")
ABI_VERSION = 10

DTYPE_F32, DTYPE_BF16 = 0, 1
ARGMAX_I32, ARGMAX_U16 = 0, 1

_P = c_void_p
_LL = c_longlong

# name -> (restype, argtypes); mirrors include/soswsod_b200.h one to one
SIGNATURES = {
    "soswsod_abi_version": (c_int, []),
    "soswsod_last_error": (c_char_p, []),
    "soswsod_roi_pool_forward": (c_int, [_P, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_float, _P, c_float,
                                          _P, _P, c_int, _P, _LL, _P, c_size_t, _P]),
    "soswsod_roi_pool_plan_bytes": (c_size_t, [c_int, c_int, c_int]),
    "soswsod_roi_pool_plan": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, _P, c_float, _P, c_size_t, _P]),
    "soswsod_roi_pool_backward": (c_int, [_P, c_int, _LL, _P, c_int, _P, c_int, _P, c_float, c_int, c_int, c_int, c_int,
                                           c_int, c_int, c_float, _P, _P, c_size_t, _P]),
    "soswsod_gemm_bf16": (c_int, [_P, _LL, c_int, _P, _LL, c_int, _P, _LL, c_int, c_int, c_int, c_int, _P, c_int, _P,
                                   _LL, c_float, c_float, c_ulonglong, _P, _P]),
    "soswsod_dropout_mask": (c_int, [_P, c_int, c_int, c_float, c_ulonglong, _P]),
    "soswsod_cast_f32_bf16": (c_int, [_P, _LL, c_int, c_int, _P, _P, _LL, _P, _LL, _P, _LL, c_float, _P]),
    "soswsod_transpose_bf16": (c_int, [_P, _LL, c_int, c_int, _P, _LL, _P]),
    "soswsod_colsum_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "soswsod_colsum": (c_int, [_P, c_int, _LL, c_int, c_int, _P, _P, c_size_t, _P]),
    "soswsod_wsddn_forward": (c_int, [_P, _LL, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _LL, _P]),
    "soswsod_oicr_avg_scores": (c_int, [_P, _P, _LL, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "soswsod_oicr_mine_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "soswsod_image_level_gt": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "soswsod_oicr_mine_label": (c_int, [_P, _LL, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_float, c_float, c_float,
                                         c_float, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "soswsod_oicr_loss": (c_int, [_P, _LL, c_int, c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float,
                                   c_float, c_float, c_float, _P, _P, _P, _P, _LL, _P]),
    "soswsod_predict": (c_int, [_P, _LL, c_int, c_int, _P, c_int, c_int, c_int, c_float, c_float, c_float, c_float, _P,
                                 _P, _P]),
    "soswsod_tta_accumulate": (c_int, [_P, _P, c_int, c_int, c_float, c_float, c_int, c_float, c_int, c_float, _P, _P,
                                        _P]),
    "soswsod_tta_views": (c_int, [_P, c_int, _P, c_int, c_float, _P, _P, _P, _P]),
    "soswsod_tta_merge": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "soswsod_nms_workspace_bytes": (c_size_t, [c_int]),
    "soswsod_nms": (c_int, [_P, _P, c_int, c_float, _P, _P, _P, c_size_t, _P]),
    "soswsod_detect_workspace_bytes": (c_size_t, [c_int, c_int]),
    "soswsod_detect": (c_int, [_P, _P, c_int, c_int, c_float, c_float, c_float, c_float, c_int, _P, _P, _P, _P, _P, _P,
                                c_size_t, _P]),
    "soswsod_pgf": (c_int, [_P, _P, _P, _P, c_int, c_double, c_double, c_int, c_ulonglong, c_ulonglong, _P, _P]),
    "soswsod_sgd_step": (c_int, [_P, _P, _P, c_longlong, c_float, c_float, c_float, c_float, _P, _P]),
    "soswsod_sgd_multi": (c_int, [_P, c_int, c_float, c_float, _P]),
    "soswsod_sgd_nvls": (c_int, [_P, c_int, c_float, c_float, c_int, _P]),
}


class SgdNvlsTensor(ctypes.Structure):
    """soswsod_sgd_nvls_tensor (include/soswsod_b200.h)."""
    _fields_ = [("param", c_void_p), ("grad_mc", c_void_p), ("momentum_buf", c_void_p), ("out_bf16_mc", c_void_p),
                ("n", c_longlong), ("lr", c_float), ("weight_decay", c_float)]


SGD_NVLS_MAX_TENSORS = 8


class SgdTensor(ctypes.Structure):
    """soswsod_sgd_tensor (include/soswsod_b200.h)."""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("momentum_buf", c_void_p), ("out_bf16", c_void_p),
                ("out_f32", c_void_p), ("n", c_longlong), ("lr", c_float), ("weight_decay", c_float)]


SGD_MAX_TENSORS = 32

_lib = None


def load() -> ctypes.CDLL:
    """Loads the library (once), checks the ABI version and installs the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built "
            "(run `python -m sos_wsod_b200.build`).  There is no CPU fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    v = lib.soswsod_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"libsoswsod_b200.so ABI version {v} != expected {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().soswsod_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
