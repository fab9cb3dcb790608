"""GPU parity at the BASELINE.json shapes themselves (configs[1..4]) and on peaked logits.

cfg2 / cfg3 run the whole batch against the C port of the oracle (oracle/c, float64, pinned to the restatement and
through it to the reference: tests/test_oracle_cport.py); cfg4 and a cfg5 shard run at full size on the device and a
random subset of >= 64 sequences is compared (sequences are independent: NoBlankCTC.py:90-126 never mixes b).
Tolerances (north star): loss 1e-5, gradient L2 1e-5 and Linf 1e-5 relative; rows t >= T_b exactly zero.
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import cport

pytestmark = pytest.mark.gpu

TOL = 1e-5
DEV = "cuda:0"


@pytest.fixture(scope="module")
def nb():
    import ctc_b200
    assert torch.cuda.is_available()
    return ctc_b200


def lengths(rs, B, T, Lmin, Lmax, ragged):
    tl = rs.randint(Lmin, Lmax + 1, size=B).astype(np.int64)
    if ragged:
        il = np.array([rs.randint(max(l, T // 2), T + 1) for l in tl], dtype=np.int64)  # T_b ~ U{max(L_b,T/2)..T}
    else:
        il = np.full(B, T, dtype=np.int64)
    return il, tl


def labels(rs, B, C, Lmax, tl):
    lab = rs.randint(0, C, size=(B, Lmax)).astype(np.int32)
    lab[np.arange(Lmax)[None, :] >= tl[:, None]] = -1
    return lab


def check_subset(per, grad_sub, x_sub, tg_sub, il_sub, tl_sub, B_total, kind="ctc"):
    """per: device losses of the subset; grad_sub: (T, n, C) device gradient columns of the subset (weight 1/B_total)."""
    ref = cport.loss_grad(kind, x_sub, tg_sub, il_sub, tl_sub, reduction="sum")
    rl = np.max(np.abs(per - ref["per_seq"]) / np.abs(ref["per_seq"]))
    g = grad_sub.astype(np.float64) * B_total
    rg = rel_l2(g, ref["grad"])
    linf = np.max(np.abs(g - ref["grad"])) / np.max(np.abs(ref["grad"]))
    assert rl < TOL, f"loss rel err {rl}"
    assert rg < TOL, f"grad L2 rel err {rg}"
    assert linf < TOL, f"grad Linf rel err {linf}"
    T = x_sub.shape[0]
    dead = np.arange(T)[:, None] >= il_sub[None, :]
    assert np.all(grad_sub[dead] == 0.0)
    return rl, rg, linf


@pytest.mark.parametrize("flags", [0, 8, 32], ids=["default", "lockstep", "seqwarp"])
@pytest.mark.parametrize("ragged", [False, True], ids=["fixedT", "raggedT"])
def test_cfg2_full_batch(nb, ragged, flags):
    """BASELINE configs[1]: B=4096, T=256, C=157, ragged L in [1,32]; every sequence against the oracle."""
    T, B, C, Lmax = 256, 4096, 157, 32
    rs = np.random.RandomState(20 + ragged)
    x = rs.standard_normal((T, B, C)).astype(np.float32)
    il, tl = lengths(rs, B, T, 1, Lmax, ragged)
    lab = labels(rs, B, C, Lmax, tl)
    xt = torch.tensor(x, device=DEV, requires_grad=True)
    per = nb.no_blank_ctc_loss(xt, torch.tensor(lab, device=DEV), torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV), "none",
                               flags=flags)
    (per.sum() / B).backward()
    torch.cuda.synchronize()
    check_subset(per.detach().cpu().double().numpy(), xt.grad.cpu().numpy(), x, lab, il, tl, B)


def test_cfg3_full_batch(nb):
    """BASELINE configs[2]: NoBlankBinaryCTC, B=4096, T=256, C=157, multi-hot density 0.03, L in [1,32], ragged T."""
    T, B, C, Lmax = 256, 4096, 157, 32
    rs = np.random.RandomState(30)
    x = rs.standard_normal((T, B, C)).astype(np.float32)
    il, tl = lengths(rs, B, T, 1, Lmax, True)
    y = (rs.uniform(size=(B, Lmax, C)) < 0.03).astype(np.float32)
    pick = rs.randint(0, C, size=(B, Lmax))
    bi, si = np.meshgrid(np.arange(B), np.arange(Lmax), indexing="ij")
    y[bi, si, pick] = 1.0
    y[np.arange(Lmax)[None, :] >= tl[:, None]] = 0.0
    xt = torch.tensor(x, device=DEV, requires_grad=True)
    per = nb.no_blank_binary_ctc_loss(xt, torch.tensor(y, device=DEV), torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV), "none")
    (per.sum() / B).backward()
    torch.cuda.synchronize()
    check_subset(per.detach().cpu().double().numpy(), xt.grad.cpu().numpy(), x, y, il, tl, B, kind="bctc")


def _device_logits(T, B, C, seed):
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    return torch.randn((T, B, C), device=DEV, dtype=torch.float32, generator=g)


def _big_case(nb, T, B, C, Lmax, seed, n_sub, flags=0):
    rs = np.random.RandomState(seed)
    il, tl = lengths(rs, B, T, 1, Lmax, True)
    lab = labels(rs, B, C, Lmax, tl)
    xt = _device_logits(T, B, C, seed).requires_grad_(True)
    per = nb.no_blank_ctc_loss(xt, torch.tensor(lab, device=DEV), torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV), "none",
                               flags=flags)
    (per.sum() / B).backward()
    torch.cuda.synchronize()
    # whole batch on the device: finite losses, zero rows beyond input_length, gradient rows sum to zero
    assert bool(torch.isfinite(per).all())
    ild = torch.tensor(il, device=DEV)
    dead = torch.arange(T, device=DEV)[:, None] >= ild[None, :]
    assert float(xt.grad.abs().amax(dim=2)[dead].max() if bool(dead.any()) else 0.0) == 0.0
    assert float(xt.grad.sum(dim=2).abs().max()) * B < 2e-4
    sub = np.sort(rs.choice(B, n_sub, replace=False))
    # the longest and the shortest sequences are always part of the sample
    sub[0], sub[1] = int(np.argmax(il * tl)), int(np.argmin(il * tl))
    sub = np.unique(sub)
    subd = torch.tensor(sub, device=DEV)
    x_sub = xt.detach()[:, subd].cpu().numpy()
    g_sub = xt.grad[:, subd].cpu().numpy()
    return check_subset(per.detach()[subd].cpu().double().numpy(), g_sub, x_sub, lab[sub], il[sub], tl[sub], B)


@pytest.mark.parametrize("flags", [0, 8], ids=["default", "lockstep"])
def test_cfg4_full_size_subset(nb, flags):
    """BASELINE configs[3]: B=1024, T=4096, C=1024, L<=256, ragged input_length (16 states per chain lane)."""
    _big_case(nb, 4096, 1024, 1024, 256, seed=40, n_sub=64, flags=flags)


@pytest.mark.parametrize("flags", [0, 8], ids=["default", "lockstep"])
def test_cfg5_shard_full_size_subset(nb, flags):
    """BASELINE configs[4] shard at 8 GPUs: B=8192, T=512, C=157, L<=64, ragged input_length."""
    _big_case(nb, 512, 8192, 157, 64, seed=50, n_sub=96, flags=flags)


PEAKED = [
    # T, B, C, Lmax, Lmin, boost, label0 (the boosted class is the label of this state; -1 = nowhere)
    (256, 24, 157, 32, 20, 14.0, -1),
    (256, 24, 157, 32, 24, 14.0, 0),
    (512, 12, 157, 64, 50, 14.0, 0),     # inf with one shared exponent, oracle loss ~880
    (512, 12, 157, 64, 50, 25.0, 30),
    (300, 10, 64, 100, 80, 20.0, 40),
    (1024, 6, 1024, 256, 200, 8.0, 0),   # inf against ~3200
    (1024, 6, 64, 256, 230, 20.0, 100),
]


@pytest.mark.parametrize("flags", [0, 1, 8, 32], ids=["default", "generic", "lockstep", "seqwarp"])
@pytest.mark.parametrize("case", PEAKED, ids=lambda c: "T%d_B%d_C%d_L%d_%d_boost%g_s%d" % c)
def test_peaked_logits(nb, case, flags):
    """A model that is confident in one class for the whole sequence: the lattice states differ by thousands of binary
    orders of magnitude.  Loss and gradient stay within the bar on the default and on the generic path."""
    T, B, C, Lmax, Lmin, boost, s0 = case
    rs = np.random.RandomState(60 + T + Lmax + int(boost))
    x = rs.standard_normal((T, B, C)).astype(np.float32)
    x[:, :, 0] += boost
    tl = rs.randint(Lmin, Lmax + 1, size=B).astype(np.int64)
    il = np.full(B, T, dtype=np.int64)
    lab = rs.randint(1, C, size=(B, Lmax)).astype(np.int32)
    if s0 >= 0:
        lab[:, s0] = 0
    lab[np.arange(Lmax)[None, :] >= tl[:, None]] = -1
    xt = torch.tensor(x, device=DEV, requires_grad=True)
    per = nb.no_blank_ctc_loss(xt, torch.tensor(lab, device=DEV), torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV), "none",
                               flags=flags)
    (per.sum() / B).backward()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(per).all())
    check_subset(per.detach().cpu().double().numpy(), xt.grad.cpu().numpy(), x, lab, il, tl, B)


@pytest.mark.parametrize("kind", ["ctc", "bctc"])
def test_reduction_none_backward_with_random_weights(nb, kind):
    """reduction='none' backpropagated with arbitrary per-sequence upstream gradients (the per_seq branch of the
    backward rescale), including negative and zero entries."""
    T, B, C, Lmax = 70, 13, 157, 12
    rs = np.random.RandomState(70)
    x = rs.standard_normal((T, B, C)).astype(np.float32)
    il, tl = lengths(rs, B, T, 1, Lmax, True)
    w = rs.standard_normal(B).astype(np.float32)
    w[3] = 0.0
    xt = torch.tensor(x, device=DEV, requires_grad=True)
    if kind == "ctc":
        tg = labels(rs, B, C, Lmax, tl)
        per = nb.no_blank_ctc_loss(xt, torch.tensor(tg, device=DEV), torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV), "none")
    else:
        tg = (rs.uniform(size=(B, Lmax, C)) < 0.05).astype(np.float32)
        tg[:, :, 5] = 1.0
        tg[np.arange(Lmax)[None, :] >= tl[:, None]] = 0.0
        per = nb.no_blank_binary_ctc_loss(xt, torch.tensor(tg, device=DEV), torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV), "none")
    per.backward(torch.tensor(w, device=DEV))
    torch.cuda.synchronize()
    ref = cport.loss_grad(kind, x, tg, il, tl, reduction="sum")
    want = ref["grad"] * w.astype(np.float64)[None, :, None]
    g = xt.grad.cpu().numpy().astype(np.float64)
    assert rel_l2(g, want) < TOL
    assert np.max(np.abs(g - want)) / np.max(np.abs(want)) < TOL
    assert np.all(g[:, 3] == 0.0)
