"""world_size-2 gloo test of the batch-sharding host logic: each rank evaluates its own
contiguous slice of the batch, ONE all-reduce of a float64 scalar gives the global mean, and the local
gradients already carry 1/B_global.  The per-rank loss here is the CPU oracle (test infrastructure):
the CUDA loss cannot run in this container; what is under test is ctc_b200.dist."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ctc_b200.dist import ShardedLoss, shard_batch
    from helpers import make_ctc_case
    from oracle import restatement as R

    x, lab, il, tl = make_ctc_case(77, 20, 10, 9, 5)
    B = x.shape[1]
    lo, hi = shard_batch(rank, world, B)

    class _OracleSum(torch.autograd.Function):
        @staticmethod
        def forward(ctx, logits, targets, il_, tl_, total):
            r = R.nbctc_loss_grad(logits.numpy(), targets.numpy(), il_.numpy(), tl_.numpy(), "sum")
            ctx.g = torch.tensor(r["grad"] / total)
            return torch.tensor(r["loss"] / total, dtype=torch.float64)

        @staticmethod
        def backward(ctx, g):
            return ctx.g * g, None, None, None, None

    loss_mod = ShardedLoss(lambda *a: _OracleSum.apply(*a))
    xl = torch.tensor(np.ascontiguousarray(x[:, lo:hi]), dtype=torch.float64, requires_grad=True)
    loss = loss_mod(xl, torch.tensor(lab[lo:hi]), torch.tensor(il[lo:hi]), torch.tensor(tl[lo:hi]), total_batch=B)
    loss.backward()
    ref = R.nbctc_loss_grad(x, lab, il, tl, "mean")
    ok = abs(float(loss) - ref["loss"]) < 1e-5 * abs(ref["loss"])
    ok &= np.allclose(xl.grad.numpy(), ref["grad"][:, lo:hi], atol=1e-12)
    # the same with the all-reduce only enqueued: backward first, the value after work.wait()
    amod = ShardedLoss(lambda *a: _OracleSum.apply(*a), async_reduce=True)
    xa = torch.tensor(np.ascontiguousarray(x[:, lo:hi]), dtype=torch.float64, requires_grad=True)
    aloss, work = amod(xa, torch.tensor(lab[lo:hi]), torch.tensor(il[lo:hi]), torch.tensor(tl[lo:hi]), total_batch=B)
    aloss.backward()
    work.wait()
    ok &= abs(float(aloss) - ref["loss"]) < 1e-5 * abs(ref["loss"])
    ok &= np.allclose(xa.grad.numpy(), ref["grad"][:, lo:hi], atol=1e-12)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_mean_two_ranks_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_batch_covers_everything():
    from ctc_b200.dist import shard_batch
    for B in (1, 7, 8, 65536):
        for W in (1, 2, 3, 8):
            spans = [shard_batch(r, W, B) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
