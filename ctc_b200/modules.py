"""Drop-in ``nn.Module`` mirrors of the reference losses (same names, same call signature).

    reference                                   this package
    NoBlankCTC()        NoBlankCTC.py:22-26     ctc_b200.NoBlankCTC()
    NoBlankBinaryCTC()  NoBlankBinaryCTC.py:22  ctc_b200.NoBlankBinaryCTC()
    loss = m(yseq, label, input_length, target_length)      (train.py:427, :576)

Both are parameter- and buffer-free (``state_dict()`` is empty, so existing checkpoints written by
checkpoints.py:25-75 load unchanged) and, unlike the reference (which caches shapes on ``self``,
NoBlankCTC.py:130-131), re-entrant.
"""
from __future__ import annotations

import torch.nn as nn

from . import _ffi
from .function import no_blank_binary_ctc_loss, no_blank_ctc_loss


class NoBlankCTC(nn.Module):
    """No-blank CTC loss over raw logits ``yseq`` (T,B,C); ``label`` (B,Lmax) int, padded with -1.

    Returns the mean over the batch of ``-log p(label_b | yseq_b)`` (NoBlankCTC.py:139-140) as a
    0-dim float32 CUDA tensor with a ``grad_fn``.  ``reduction`` in {'mean','sum','none'} is an
    addition; the default reproduces the reference.
    """

    def __init__(self, reduction: str = "mean", flags: int = _ffi.FLAG_DEFAULT):
        super().__init__()
        self.reduction = reduction
        self.flags = flags

    def forward(self, yseq, label, input_length, target_length, reduction=None):
        return no_blank_ctc_loss(yseq, label, input_length, target_length, reduction or self.reduction,
                                 flags=self.flags)

    def extra_repr(self):
        return f"reduction={self.reduction!r}"


class NoBlankBinaryCTC(nn.Module):
    """Multi-label variant: ``label`` is a (B,Lmax,C) float multi-hot tensor; per-state emission is
    ``-BCELoss(sigmoid(yseq[t,b]), label[b,s])`` (NoBlankBinaryCTC.py:109-112,:146)."""

    def __init__(self, reduction: str = "mean", flags: int = _ffi.FLAG_DEFAULT):
        super().__init__()
        self.reduction = reduction
        self.flags = flags

    def forward(self, yseq, label, input_length, target_length, reduction=None):
        return no_blank_binary_ctc_loss(yseq, label, input_length, target_length, reduction or self.reduction,
                                        flags=self.flags)

    def extra_repr(self):
        return f"reduction={self.reduction!r}"
