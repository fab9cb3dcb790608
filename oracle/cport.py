"""ctypes wrapper of oracle/c/nbctc_oracle.c (plain-C float64 + OpenMP restatement).

TEST INFRASTRUCTURE: used by tests/, and by bench.py only for the reported ``cpu_baseline`` and the
``--impl reference`` arm.  Never imported by ``ctc_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libnbctc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "c", "nbctc_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "_build/libnbctc_oracle.so"], check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        l.nbctc_oracle.restype = C.c_int
        l.nbctc_oracle.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.nbctc_oracle_threads.restype = C.c_int
        _lib = l
    return _lib


def threads() -> int:
    return int(lib().nbctc_oracle_threads())


def loss_grad(kind, logits, targets, input_length, target_length, reduction="mean", want_grad=True):
    """kind 'ctc' | 'bctc'.  Returns dict(loss, per_seq, grad) in float64 (same contract as restatement)."""
    x = np.ascontiguousarray(logits, dtype=np.float32)
    T, B, Cc = x.shape
    if kind == "ctc":
        tg = np.ascontiguousarray(targets, dtype=np.int32)
    else:
        tg = np.ascontiguousarray(targets, dtype=np.float32)
    Lmax = tg.shape[1]
    il = np.ascontiguousarray(input_length, dtype=np.int64)
    tl = np.ascontiguousarray(target_length, dtype=np.int64)
    wv = np.full(B, 1.0 / B if reduction == "mean" else 1.0)
    per = np.empty(B, dtype=np.float64)
    grad = np.empty((T, B, Cc), dtype=np.float64) if want_grad else None
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = lib().nbctc_oracle(0 if kind == "ctc" else 1, p(x), T, B, Cc, p(tg), Lmax, p(il), p(tl), p(wv), p(per), p(grad))
    if rc != 0:
        raise MemoryError("nbctc_oracle: allocation failed")
    loss = per.mean() if reduction == "mean" else per.sum() if reduction == "sum" else per
    return dict(loss=loss, per_seq=per, grad=grad)
