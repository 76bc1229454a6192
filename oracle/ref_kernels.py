"""ctypes loader for oracle/ref_kernels.c (TEST INFRASTRUCTURE ONLY; see that file's header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libref_kernels.so")


def build() -> str:
    src = os.path.join(_HERE, "ref_kernels.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ref_nms.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def roi_pool_forward(feat: np.ndarray, rois: np.ndarray, P: int, scale: float):
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    N, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.empty((R, C, P, P), np.float32)
    arg = np.empty((R, C, P, P), np.int32)
    lib().ref_roi_pool_forward(_p(feat, ctypes.c_float), N, C, H, W, _p(rois, ctypes.c_float), R, P, P,
                               ctypes.c_float(scale), _p(out, ctypes.c_float), _p(arg, ctypes.c_int32))
    return out, arg


def roi_pool_backward(grad_out: np.ndarray, argmax: np.ndarray, rois: np.ndarray, shape):
    N, C, H, W = shape
    g = np.ascontiguousarray(grad_out, dtype=np.float32)
    a = np.ascontiguousarray(argmax, dtype=np.int32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    R, _, P, _ = g.shape
    gf = np.zeros((N, C, H, W), np.float32)
    lib().ref_roi_pool_backward(_p(g, ctypes.c_float), _p(a, ctypes.c_int32), _p(rois, ctypes.c_float), R, C,
                                H, W, P, P, _p(gf, ctypes.c_float))
    return gf


def nms(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    keep = np.empty((max(n, 1),), np.int64)
    nk = lib().ref_nms(_p(boxes, ctypes.c_float), _p(scores, ctypes.c_float), n, ctypes.c_float(thr),
                       _p(keep, ctypes.c_int64))
    return keep[:nk].copy()


def pairwise_iou(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().ref_pairwise_iou(_p(a, ctypes.c_float), a.shape[0], _p(b, ctypes.c_float), b.shape[0],
                           _p(out, ctypes.c_float))
    return out


def tta_transform_proposals(boxes: np.ndarray, hw, new_hw, flipped: bool, min_box_size: float = 0.0):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    out = np.empty((n, 4), np.float32)
    keep = np.empty((n,), np.uint8)
    lib().ref_tta_transform_proposals(_p(boxes, ctypes.c_float), n, int(hw[0]), int(hw[1]), int(new_hw[0]), int(new_hw[1]),
                                      int(flipped), ctypes.c_float(min_box_size), _p(out, ctypes.c_float),
                                      _p(keep, ctypes.c_uint8))
    return out, keep.astype(bool)


def tta_inverse_boxes(boxes: np.ndarray, hw, new_hw, flipped: bool, post_hw=None) -> np.ndarray:
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    out = np.empty((n, 4), np.float32)
    ph, pw = (int(post_hw[0]), int(post_hw[1])) if post_hw is not None else (0, 0)
    lib().ref_tta_inverse_boxes(_p(boxes, ctypes.c_float), n, int(hw[0]), int(hw[1]), int(new_hw[0]), int(new_hw[1]),
                                int(flipped), ph, pw, _p(out, ctypes.c_float))
    return out
