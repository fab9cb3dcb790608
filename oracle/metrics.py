"""CPU restatement of the reference's matching metrics (TEST INFRASTRUCTURE).

Plain-Python/numpy ports of ``train.py:41-56`` (accuracy_s), ``train.py:59-78`` (accuracy),
``train.py:82-107`` (recall_time) and ``train.py:111-136`` (accuracy_time).  All outputs are
integer 0/1 flags (bit-exact bar); the percentages the reference returns are simple ratios
of their sums.  Pinned against the reference by ``tests/golden/metrics_*.npz``
(``tests/golden/make_golden_metrics.py`` imports the unmodified ``train.py`` with a stub for the
absent, unused ``matplotlib``).  Never import this from ``ctc_b200``.

Tie order: ``torch.topk`` (train.py:48,65,90,118) gives no guarantee; here and in the CUDA
kernel ties go to the LOWER class index.  The fixtures use random floats (no ties).
"""
from __future__ import annotations

import numpy as np


def frame_topk(x: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k largest entries of the last axis, best first, ties -> lower index, NaN ranks above
    +inf (as torch.topk does), -0 == +0; -1 where k exceeds the row length."""
    xs = x.astype(np.float64)
    xs = np.where(np.isposinf(xs), 1e300, xs)     # float32 scores: every finite value is below 1e300
    xs = np.where(np.isnan(xs), np.inf, xs)
    order = np.argsort(-xs, axis=-1, kind="stable")[..., :k].astype(np.int32)
    if order.shape[-1] < k:
        pad = np.full(order.shape[:-1] + (k - order.shape[-1],), -1, np.int32)
        order = np.concatenate([order, pad], axis=-1)
    return order


def accuracy_time_flags(pred: np.ndarray, target: np.ndarray, time: int) -> np.ndarray:
    """train.py:111-136.  pred (temporal, K) top-k classes per frame; target (>=time, C) multi-hot.
    Returns correct (K, temporal) int32: frame j of rank i matches the first target row t >= current_id[i]
    that contains the predicted class; current_id[i] then becomes t (monotone, rows may repeat)."""
    temporal, K = pred.shape
    correct = np.zeros((K, temporal), np.int32)
    for i in range(K):
        cur = 0
        for j in range(temporal):
            for t in range(cur, time):
                if target[t, pred[j, i]] > 0.5:
                    correct[i, j] = 1
                    cur = t
                    break
    return correct


def recall_time_flags(pred: np.ndarray, target: np.ndarray, trans: int) -> np.ndarray:
    """train.py:82-107.  Returns correct (K, trans) int32: target row t is recalled by rank i.
    Quirk kept: the frame loop runs over ``correct.shape[1]`` = trans (train.py:96), i.e. only the FIRST
    `trans` frames are looked at (fewer if the clip is shorter)."""
    temporal, K = pred.shape
    correct = np.zeros((K, trans), np.int32)
    for i in range(K):
        cur = 0
        for j in range(min(trans, temporal)):
            for t in range(cur, trans):
                if target[t, pred[j, i]] > 0.5:
                    correct[i, t] = 1
                    cur = t
                    break
    return correct


def accuracy_s_flags(pred: np.ndarray, label: np.ndarray) -> np.ndarray:
    """train.py:41-56.  pred (B, K) top-k classes, label (B,) class index.  Returns correct (K, B) int32."""
    return (pred.T == label[None, :]).astype(np.int32)


def accuracy_flags(pred: np.ndarray, target: np.ndarray) -> np.ndarray:
    """train.py:59-78.  pred (B, K), target (B, C) multi-hot.  Returns correct (K, B) int32."""
    B, K = pred.shape
    return (target[np.arange(B)[None, :], pred.T] > 0.5).astype(np.int32)


def percentages(correct: np.ndarray, denom: int, topk=(1, 5)):
    """res[k] = 100 * sum(correct[:k]) / denom, as the reference returns them (train.py:52-55 etc.)."""
    return [100.0 * float(correct[:k].sum()) / denom for k in topk]
