"""torch.autograd.Function wrappers around the C ABI (PyTorch = tensor plumbing only).

Follows the reference's own Function idiom (models/layers/BalanceLabels.py:11-21:
staticmethod forward/backward, tuple of grads).  The gradient with respect to the logits is
produced by the same fused kernel that computes the loss (one read of the logits, one write
of the gradient); ``backward`` only rescales it by the upstream gradient, and that rescale is
skipped on the device when the upstream gradient is exactly 1 (``loss.backward()``).
"""
from __future__ import annotations

import contextlib

import torch

from . import _ffi

_REDUCTIONS = ("mean", "sum", "none")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _prep_lengths(t, device, B, name):
    if torch.is_tensor(t) and t.dtype == torch.int64 and t.device == device and t.dim() == 1 and t.shape[0] == B \
            and t.is_contiguous():
        return t  # the usual case (train.py:397,399): no tensor ops at all
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    if t.numel() != B:
        raise ValueError(f"{name} must have {B} elements, got {tuple(t.shape)}")
    return t.reshape(B).to(device=device, dtype=torch.int64).contiguous()


_WS_BYTES = {}  # (T, B, C, Lmax, binary, flags) -> workspace bytes (a pure function of the shape: one FFI call per shape)


def _on_device(dev):
    """Context that makes ``dev`` current; free when it already is (small batches are host-bound)."""
    if torch.cuda.current_device() == dev.index:
        return contextlib.nullcontext()
    return torch.cuda.device(dev)


def _launch(x, targets, in_len, tgt_len, binary, want_grad, w_scalar, seq_w, flags):
    """Run the fused loss(+grad) on the current stream.  Returns (per_seq, sum64, reduced, grad)."""
    T, B, C = x.shape
    Lmax = targets.shape[1]
    dev = x.device
    lib = _ffi.lib()
    per_seq = torch.empty(B, dtype=torch.float32, device=dev)
    loss_sum = torch.empty((), dtype=torch.float64, device=dev)
    reduced = torch.empty((), dtype=torch.float32, device=dev)
    grad = torch.empty_like(x) if want_grad else None
    if not want_grad:
        flags |= _ffi.FLAG_NO_GRAD
    if x.data_ptr() % 16 == 0 and (grad is None or grad.data_ptr() % 16 == 0):
        flags |= _ffi.FLAG_ALIGNED16    # no workspace for the unaligned-tensor fallback (torch allocations are aligned)
    key = (T, B, C, Lmax, binary, flags)
    ws_bytes = _WS_BYTES.get(key)
    if ws_bytes is None:
        ws_bytes = _WS_BYTES[key] = int(lib.nbctc_workspace_bytes(T, B, C, Lmax, 1 if binary else 0, flags))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    fn = lib.nbbctc_loss_grad_f32 if binary else lib.nbctc_loss_grad_f32
    with _on_device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = fn(x.data_ptr(), T, B, C, targets.data_ptr(), Lmax, in_len.data_ptr(), tgt_len.data_ptr(),
                per_seq.data_ptr(), loss_sum.data_ptr(), reduced.data_ptr(), _ptr(grad), _ptr(seq_w),
                float(w_scalar), ws.data_ptr(), ws_bytes, flags, stream)
    _ffi.check(rc, "nbbctc_loss_grad_f32" if binary else "nbctc_loss_grad_f32")
    return per_seq, loss_sum, reduced, grad


class _NoBlankCTCFunction(torch.autograd.Function):
    """forward(logits, targets, input_length, target_length, reduction, binary, total_batch, flags, out64, want_grad)"""

    @staticmethod
    def forward(ctx, logits, targets, input_length, target_length, reduction, binary, total_batch, flags, out64,
                want_grad):
        if reduction not in _REDUCTIONS:
            raise ValueError(f"reduction must be one of {_REDUCTIONS}, got {reduction!r}")
        if logits.dim() != 3:
            raise ValueError(f"logits must be (T,B,C), got {tuple(logits.shape)}")
        if not logits.is_cuda:
            raise _ffi.NbctcError("ctc_b200 is CUDA-only: logits must be a CUDA tensor (there is no CPU fallback)")
        T, B, C = logits.shape
        dev = logits.device
        x = logits.detach()
        if x.dtype != torch.float32:
            x = x.float()
        if not x.is_contiguous():
            x = x.contiguous()
        if binary:
            if targets.dim() != 3 or targets.shape[0] != B or targets.shape[2] != C:
                raise ValueError(f"multi-hot targets must be (B,Lmax,C)=({B},L,{C}), got {tuple(targets.shape)}")
            tg = targets if (targets.dtype == torch.float32 and targets.device == dev and targets.is_contiguous()
                             and not targets.requires_grad) \
                else targets.detach().to(device=dev, dtype=torch.float32).contiguous()
        else:
            if targets.dim() != 2 or targets.shape[0] != B:
                raise ValueError(f"labels must be (B,Lmax)=({B},L), got {tuple(targets.shape)}")
            tg = targets if (targets.dtype == torch.int32 and targets.device == dev and targets.is_contiguous()) \
                else targets.detach().to(device=dev, dtype=torch.int32).contiguous()
        il = _prep_lengths(input_length, dev, B, "input_length")
        tl = _prep_lengths(target_length, dev, B, "target_length")
        # grad mode is always off inside forward(); the caller samples it (validate() runs under no_grad)
        want_grad = bool(want_grad) and bool(ctx.needs_input_grad[0])
        nb = int(total_batch) if total_batch else B
        w = 1.0 / nb if reduction == "mean" else 1.0
        if out64 and reduction != "none":
            flags = int(flags) | _ffi.FLAG_SUM_WEIGHTED   # the kernel writes w * sum in float64: no torch op after the call
        per_seq, loss_sum, reduced, grad = _launch(x, tg, il, tl, binary, want_grad, w, None, int(flags))
        ctx.grad = grad
        ctx.per_seq_out = reduction == "none"
        ctx.in_dtype = logits.dtype
        ctx.shape = (T, B, C)
        if reduction == "none":
            return per_seq
        if out64:
            return loss_sum
        return reduced

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        grad = ctx.grad
        if grad is None:
            raise RuntimeError("NoBlankCTC backward called twice (or without a gradient buffer); "
                               "the fused gradient is consumed by the first backward")
        ctx.grad = None
        T, B, C = ctx.shape
        go = grad_out.detach().to(device=grad.device, dtype=torch.float32).contiguous()
        lib = _ffi.lib()
        with _on_device(grad.device):
            stream = torch.cuda.current_stream(grad.device).cuda_stream
            rc = lib.nbctc_scale_grad_f32(grad.data_ptr(), T, B, C, go.data_ptr(), 1 if ctx.per_seq_out else 0, stream)
        _ffi.check(rc, "nbctc_scale_grad_f32")
        if grad.dtype != ctx.in_dtype:
            grad = grad.to(ctx.in_dtype)
        return grad, None, None, None, None, None, None, None, None, None


def no_blank_ctc_loss(logits, labels, input_length, target_length, reduction="mean", *, total_batch=None,
                      flags=_ffi.FLAG_DEFAULT, out64=False):
    """Functional form of :class:`ctc_b200.NoBlankCTC` (NoBlankCTC.py:129-141)."""
    want_grad = torch.is_grad_enabled() and logits.requires_grad
    return _NoBlankCTCFunction.apply(logits, labels, input_length, target_length, reduction, False, total_batch,
                                     flags, out64, want_grad)


def no_blank_binary_ctc_loss(logits, targets, input_length, target_length, reduction="mean", *, total_batch=None,
                             flags=_ffi.FLAG_DEFAULT, out64=False):
    """Functional form of :class:`ctc_b200.NoBlankBinaryCTC` (NoBlankBinaryCTC.py:139-151)."""
    want_grad = torch.is_grad_enabled() and logits.requires_grad
    return _NoBlankCTCFunction.apply(logits, targets, input_length, target_length, reduction, True, total_batch,
                                     flags, out64, want_grad)


class _CtcPlusCeFunction(torch.autograd.Function):
    """CTC + alpha * CE on one frame per sequence (train.py:353, models/__init__.py:85-86).

    The CTC gradient comes from the fused kernel; ``nbctc_aux_ce_f32`` then ADDS alpha * d CE / d logits into the B rows
    it touches, on the same stream.  Returns (total, ctc, ce) -- the three meters of train.py:439-441."""

    @staticmethod
    def forward(ctx, logits, labels, input_length, target_length, ce_targets, alpha, frame_index, binary, total_batch,
                flags, want_grad):
        if not logits.is_cuda:
            raise _ffi.NbctcError("ctc_b200 is CUDA-only: logits must be a CUDA tensor (there is no CPU fallback)")
        T, B, C = logits.shape
        dev = logits.device
        x = logits.detach()
        if x.dtype != torch.float32:
            x = x.float()
        if not x.is_contiguous():
            x = x.contiguous()
        if binary:
            tg = labels.detach().to(device=dev, dtype=torch.float32).contiguous()
        else:
            tg = labels.detach().to(device=dev, dtype=torch.int32).contiguous()
        il = _prep_lengths(input_length, dev, B, "input_length")
        tl = _prep_lengths(target_length, dev, B, "target_length")
        want_grad = bool(want_grad) and bool(ctx.needs_input_grad[0])
        nb = int(total_batch) if total_batch else B
        w = 1.0 / nb
        per_seq, loss_sum, reduced, grad = _launch(x, tg, il, tl, binary, want_grad, w, None, int(flags))
        ce_t = ce_targets.detach().to(device=dev)
        if ce_t.dim() == 1:
            if ce_t.shape[0] != B:
                raise ValueError(f"class-index CE targets must be (B,)=({B},), got {tuple(ce_t.shape)}")
            y_idx, y_hot = ce_t.to(torch.int32).contiguous(), None
        else:
            if tuple(ce_t.shape) != (B, C):
                raise ValueError(f"multi-hot CE targets must be (B,C)=({B},{C}), got {tuple(ce_t.shape)}")
            y_idx, y_hot = None, ce_t.to(torch.float32).contiguous()
        fi = None if frame_index is None else _prep_lengths(frame_index, dev, B, "frame_index")
        ce_per_seq = torch.empty(B, dtype=torch.float32, device=dev)
        lib = _ffi.lib()
        with _on_device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.nbctc_aux_ce_f32(x.data_ptr(), T, B, C, _ptr(fi), il.data_ptr(), _ptr(y_idx), _ptr(y_hot),
                                      float(alpha) * w, None, ce_per_seq.data_ptr(), _ptr(grad), stream)
        _ffi.check(rc, "nbctc_aux_ce_f32")
        ce = ce_per_seq.sum() * w
        ctx.grad = grad
        ctx.in_dtype = logits.dtype
        ctx.shape = (T, B, C)
        ctx.mark_non_differentiable(reduced, ce)
        return reduced + float(alpha) * ce, reduced, ce

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out, _g_ctc, _g_ce):
        grad = ctx.grad
        if grad is None:
            raise RuntimeError("backward called twice (or without a gradient buffer); the fused gradient is consumed "
                               "by the first backward")
        ctx.grad = None
        T, B, C = ctx.shape
        go = grad_out.detach().to(device=grad.device, dtype=torch.float32).contiguous()
        lib = _ffi.lib()
        with _on_device(grad.device):
            stream = torch.cuda.current_stream(grad.device).cuda_stream
            rc = lib.nbctc_scale_grad_f32(grad.data_ptr(), T, B, C, go.data_ptr(), 0, stream)
        _ffi.check(rc, "nbctc_scale_grad_f32")
        if grad.dtype != ctx.in_dtype:
            grad = grad.to(ctx.in_dtype)
        return (grad,) + (None,) * 10


def ctc_plus_cross_entropy(logits, labels, input_length, target_length, ce_targets, alpha, *, frame_index=None,
                           binary=False, total_batch=None, flags=_ffi.FLAG_DEFAULT):
    """``Loss = CTC + alpha * CE`` of the reference's trainer (opts.py:74 --alpha, train.py:353) in one pass over the
    gradient: the CE is taken on the scores of ONE frame per sequence (``frame_index``, default ``input_length-1`` as
    train.py:434 classifies ``v_output[temporal-1]``).  ``ce_targets``: (B,) class indices -> ``nn.CrossEntropyLoss``
    (models/__init__.py:85); (B,C) multi-hot -> the reference's ``CrossEntropy`` module (CrossEntropy.py:17-32).
    Both losses are means over the batch.  Returns ``(total, ctc, ce)``; only ``total`` carries a gradient."""
    want_grad = torch.is_grad_enabled() and logits.requires_grad
    return _CtcPlusCeFunction.apply(logits, labels, input_length, target_length, ce_targets, alpha, frame_index,
                                    binary, total_batch, flags, want_grad)


def best_path(logits, labels, input_length, target_length, want_argmax=True):
    """Viterbi alignment on the no-blank lattice + per-frame argmax.

    Returns ``(states (B,T) int32, score (B,) float64, argmax (T,B) int32 or None)``.
    """
    if not logits.is_cuda:
        raise _ffi.NbctcError("ctc_b200 is CUDA-only")
    T, B, C = logits.shape
    dev = logits.device
    x = logits.detach().float().contiguous()
    lab = labels.detach().to(device=dev, dtype=torch.int32).contiguous()
    Lmax = lab.shape[1]
    il = _prep_lengths(input_length, dev, B, "input_length")
    tl = _prep_lengths(target_length, dev, B, "target_length")
    states = torch.empty((B, T), dtype=torch.int32, device=dev)
    score = torch.empty(B, dtype=torch.float64, device=dev)
    amax = torch.empty((T, B), dtype=torch.int32, device=dev) if want_argmax else None
    lib = _ffi.lib()
    ws_bytes = int(lib.nbctc_best_path_workspace_bytes(T, B, C, Lmax))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.nbctc_best_path_i32(x.data_ptr(), T, B, C, lab.data_ptr(), Lmax, il.data_ptr(), tl.data_ptr(),
                                     states.data_ptr(), score.data_ptr(), _ptr(amax), ws.data_ptr(), ws_bytes, stream)
    _ffi.check(rc, "nbctc_best_path_i32")
    return states, score, amax
