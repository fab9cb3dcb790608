"""Device versions of the reference's evaluation helpers (train.py:41-136).

Same names, argument meaning and return triple as the reference -- ``(res[0], res[1], correct[:1].view(-1).float())`` --
for one sample, plus ``*_batch`` forms that evaluate every sample of a batch in two kernel launches (the reference
loops over samples, frames and transcript rows in Python).  Everything runs in ``libnbctc.so``; there is no CPU path.

Differences, on purpose: ties in the top-k go to the lower class index (``torch.topk`` leaves them open), and
``accuracy_s`` works for k > 1 (train.py:54 raises on current PyTorch because it ``.view``s a non-contiguous slice).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _ffi

__all__ = ["frame_topk", "match_time", "accuracy_time", "recall_time", "accuracy_s", "accuracy",
           "accuracy_time_batch", "recall_time_batch"]


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _ffi.NbctcError(f"{what} must be a CUDA tensor (ctc_b200 has no CPU path)")


def frame_topk(scores: torch.Tensor, k: int, sample_major: bool = False) -> torch.Tensor:
    """Top-k classes of every row of `scores` (..., C), best first -> int32 (..., k).
    With ``sample_major`` and a (T, B, C) input the result is laid out (B, T, k) -- `output[b]` slices, ready for
    :func:`match_time` -- without transposing the scores."""
    _need_cuda(scores, "scores")
    if scores.dtype != torch.float32:
        scores = scores.float()
    if scores.stride(-1) != 1:
        scores = scores.contiguous()
    C = scores.shape[-1]
    with torch.cuda.device(scores.device):
        if sample_major:
            if scores.dim() != 3:
                raise ValueError("sample_major needs a (T, B, C) tensor")
            T, B, _ = scores.shape
            pred = torch.empty((B, T, k), dtype=torch.int32, device=scores.device)
            rc = _ffi.lib().nbctc_frame_topk_i32(scores.data_ptr(), B, T, scores.stride(1), scores.stride(0), C, k,
                                                 pred.data_ptr(), _stream(scores))
        else:
            flat = scores.reshape(-1, C)
            if flat.stride(-1) != 1:
                flat = flat.contiguous()
            pred = torch.empty(tuple(scores.shape[:-1]) + (k,), dtype=torch.int32, device=scores.device)
            rc = _ffi.lib().nbctc_frame_topk_i32(flat.data_ptr(), 1, flat.shape[0], 0, flat.stride(0), C, k,
                                                 pred.data_ptr(), _stream(scores))
    _ffi.check(rc, "nbctc_frame_topk_i32")
    return pred


def match_time(pred: torch.Tensor, target: torch.Tensor, time: Optional[torch.Tensor], recall: bool
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Greedy monotone matching for a batch.  pred (B, frames, K) int32, target (B, Lt, C) multi-hot, time (B) or None.
    Returns (correct, counts): correct (B, K, frames) [accuracy] or (B, K, Lt) [recall], counts (B, K), int32."""
    _need_cuda(pred, "pred")
    B, frames, K = pred.shape
    target = target.to(device=pred.device, dtype=torch.float32).contiguous()
    Lt, C = target.shape[1], target.shape[2]
    pred = pred.to(torch.int32).contiguous()
    tptr = None
    if time is not None:
        time = torch.as_tensor(time).to(device=pred.device, dtype=torch.int32).contiguous()
        tptr = time.data_ptr()
    correct = torch.empty((B, K, Lt if recall else frames), dtype=torch.int32, device=pred.device)
    counts = torch.empty((B, K), dtype=torch.int32, device=pred.device)
    with torch.cuda.device(pred.device):
        rc = _ffi.lib().nbctc_match_time_i32(pred.data_ptr(), target.data_ptr(), tptr, B, frames, K, Lt, C,
                                             1 if recall else 0, correct.data_ptr(), counts.data_ptr(), _stream(pred))
    _ffi.check(rc, "nbctc_match_time_i32")
    return correct, counts


def _triple(correct: torch.Tensor, denom: float, topk: Sequence[int]):
    """(res[0], res[1], top-1 flags) shaped like the reference's return value.  correct: (K, n) int32."""
    res = [correct[:k].sum().float().reshape(1) * (100.0 / denom) for k in topk]
    return res[0], res[1], correct[:1].reshape(-1).float()


def accuracy_time(output: torch.Tensor, target: torch.Tensor, time: int, topk=(1,)):
    """train.py:111-136.  output (temporal, C) scores, target (>= time, C) multi-hot."""
    pred = frame_topk(output, max(topk))
    tg = target[:time].unsqueeze(0)
    correct, _ = match_time(pred.unsqueeze(0), tg, None, recall=False)
    return _triple(correct[0], output.shape[0], topk)


def recall_time(output: torch.Tensor, target: torch.Tensor, trans: int, topk=(1,)):
    """train.py:82-107 (only the first `trans` frames are scanned, as there)."""
    pred = frame_topk(output, max(topk))
    tg = target[:trans].unsqueeze(0)
    correct, _ = match_time(pred.unsqueeze(0), tg, None, recall=True)
    return _triple(correct[0], trans, topk)


def _match_frame(output, label, target, topk):
    pred = frame_topk(output, max(topk))
    B, K = pred.shape
    correct = torch.empty((K, B), dtype=torch.int32, device=pred.device)
    lp = tp = None
    if label is not None:
        label = label.to(device=pred.device, dtype=torch.int32).contiguous()
        lp = label.data_ptr()
    else:
        target = target.to(device=pred.device, dtype=torch.float32).contiguous()
        tp = target.data_ptr()
    with torch.cuda.device(pred.device):
        rc = _ffi.lib().nbctc_match_frame_i32(pred.data_ptr(), lp, tp, B, K, output.shape[-1], correct.data_ptr(), _stream(pred))
    _ffi.check(rc, "nbctc_match_frame_i32")
    return _triple(correct, B, topk)


def accuracy_s(output: torch.Tensor, target: torch.Tensor, topk=(1,)):
    """train.py:41-56.  output (batch, C) scores, target (batch) class indices."""
    return _match_frame(output, target.reshape(-1), None, topk)


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)):
    """train.py:59-78.  output (batch, C) scores, target (batch, C) multi-hot."""
    return _match_frame(output, None, target, topk)


def accuracy_time_batch(output: torch.Tensor, target: torch.Tensor, time: torch.Tensor, k: int = 5):
    """All samples at once: output (T, B, C) scores, target (B, Lt, C), time (B).  -> (correct (B,k,T), counts (B,k))."""
    return match_time(frame_topk(output, k, sample_major=True), target, time, recall=False)


def recall_time_batch(output: torch.Tensor, target: torch.Tensor, trans: torch.Tensor, k: int = 5):
    """All samples at once -> (correct (B,k,Lt), counts (B,k))."""
    return match_time(frame_topk(output, k, sample_major=True), target, trans, recall=True)
