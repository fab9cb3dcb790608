"""Batch-sharded loss across the GPUs of one box.

Sequences are independent (the only cross-sequence op in the reference is the final
``torch.mean``, NoBlankCTC.py:140), so rank r owns a contiguous slice of the batch and the
only collective is ONE all-reduce (sum) of a float64 scalar -- NCCL over NVLink on the GPU
box, gloo in the CPU tests.  Gradients need no collective: they are with respect to the
rank's own logits and already carry the 1/B_global factor.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x.  d(global loss)/dx on every rank is the upstream grad itself."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.detach().clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


class _AllReduceSumAsync(torch.autograd.Function):
    """As _AllReduceSum, but the collective is only ENQUEUED (``async_op=True``): the caller's stream is not made to
    wait for it, so the next step's kernels overlap the all-reduce latency.  The handle is kept on ``holder``."""

    @staticmethod
    def forward(ctx, x, group, holder, inplace):
        if inplace:          # x is a fresh buffer nobody else reads (the loss call's own output): reduce it where it is
            ctx.mark_dirty(x)
            y = x
        else:
            y = x.detach().clone()
        holder.append(dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group, async_op=True))
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None, None, None


class _Done:
    def wait(self):
        return True


def all_reduce_sum_async(x: torch.Tensor, group=None, inplace: bool = False):
    """Autograd-transparent all-reduce(sum) that does not block the current stream.

    Returns ``(y, work)``.  ``y.backward()`` may be called at once (the gradient of a sum over ranks does not depend
    on its value); ``work.wait()`` must be called before ``y`` is READ on the current stream (it makes that stream
    wait for the collective, not the host).  With one rank: ``(x, <no-op handle>)``.  ``inplace=True`` reduces ``x``'s own
    buffer (no copy kernel in front of the collective): for the output of ``no_blank_ctc_loss(..., out64=True)``, which
    nobody else holds.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x, _Done()
    holder = []
    y = _AllReduceSumAsync.apply(x, group, holder, bool(inplace) and not x.is_leaf)
    return y, holder[0]


def all_reduce_sum(x: torch.Tensor, group=None) -> torch.Tensor:
    """Autograd-transparent all-reduce(sum); identity when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    return _AllReduceSum.apply(x, group)


def shard_batch(rank: int, world: int, B: int):
    """Contiguous slice [lo, hi) of the batch owned by ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedLoss(torch.nn.Module):
    """Wrap a per-rank loss so it returns the mean over the GLOBAL batch.

    ``local_sum_fn(logits, targets, input_length, target_length, total_batch)`` must return
    ``sum_{b local} loss_b / total_batch`` as a float64 scalar whose gradient w.r.t. the local
    logits already includes ``1/total_batch`` -- :func:`ctc_b200.no_blank_ctc_loss` with
    ``total_batch=..., out64=True`` does exactly that.
    """

    def __init__(self, local_sum_fn, group=None, async_reduce=False):
        super().__init__()
        self.local_sum_fn = local_sum_fn
        self.group = group
        self.async_reduce = async_reduce

    def forward(self, logits, targets, input_length, target_length, total_batch=None):
        """Mean loss over the global batch.  ``async_reduce=True``: returns ``(loss_float64, work)`` -- call
        ``loss.backward()`` right away and ``work.wait()`` before reading the value (see all_reduce_sum_async)."""
        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        if total_batch is None:
            total_batch = logits.shape[1] * world      # equal shards
        local = self.local_sum_fn(logits, targets, input_length, target_length, total_batch)
        if self.async_reduce:
            return all_reduce_sum_async(local, self.group, inplace=True)
        return all_reduce_sum(local, self.group).to(torch.float32)


def sharded_no_blank_ctc(binary: bool = False, group=None, async_reduce: bool = False) -> ShardedLoss:
    from .function import no_blank_binary_ctc_loss, no_blank_ctc_loss
    fn = no_blank_binary_ctc_loss if binary else no_blank_ctc_loss

    def local(logits, targets, il, tl, total_batch):
        return fn(logits, targets, il, tl, "mean", total_batch=total_batch, out64=True)

    return ShardedLoss(local, group, async_reduce)
