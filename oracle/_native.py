"""ORACLE (test infrastructure): builds and loads oracle/nms_ref.cpp via ctypes."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "nms_ref.cpp")
_OUT_DIR = os.path.join(_HERE, "_build")
_OUT = os.path.join(_OUT_DIR, "liboracle_nms.so")
_lib = None


def build(force=False):
    """g++ -O2 -ffp-contract=off (no FMA) -> oracle/_build/liboracle_nms.so"""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(_OUT)
            and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC)):
        return _OUT
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-fno-fast-math",
           "-o", _OUT, _SRC]
    subprocess.check_call(cmd)
    return _OUT


def lib():
    global _lib
    if _lib is None:
        path = _OUT if os.path.exists(_OUT) and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC) \
            else build()
        _lib = ctypes.CDLL(path)
        _lib.oracle_nms_tf113.restype = ctypes.c_int
        _lib.oracle_nms_tf113.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]
        _lib.oracle_heap_pop_order.restype = None
        _lib.oracle_heap_pop_order.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_iou.restype = ctypes.c_float
        _lib.oracle_iou.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return _lib


def nms_tf113(boxes, scores, max_out, iou_thr, return_pop_order=False):
    """tf.image.non_max_suppression (TF 1.13 semantics). Returns int32 selected indices."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    n = boxes.shape[0]
    assert scores.shape[0] == n
    sel = np.zeros(max(max_out, 1), dtype=np.int32)
    pop = np.zeros(max(n, 1), dtype=np.int32)
    npop = ctypes.c_int32(0)
    cnt = lib().oracle_nms_tf113(boxes.ctypes.data, scores.ctypes.data, n, int(max_out),
                                 ctypes.c_float(float(iou_thr)), sel.ctypes.data,
                                 pop.ctypes.data, ctypes.byref(npop))
    if return_pop_order:
        return sel[:cnt].copy(), pop[:npop.value].copy()
    return sel[:cnt].copy()


def heap_pop_order(scores):
    scores = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    out = np.zeros(max(scores.shape[0], 1), dtype=np.int32)
    lib().oracle_heap_pop_order(scores.ctypes.data, scores.shape[0], out.ctypes.data)
    return out[:scores.shape[0]]
