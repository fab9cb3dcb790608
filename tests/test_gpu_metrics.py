"""Device metrics (ctc_b200.metrics -> libnbctc.so) against the oracle and the reference-generated fixtures.
Integer outputs: bit-exact."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import metrics as M

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "metrics_*.npz")))


@pytest.fixture(scope="module")
def mt():
    import ctc_b200.metrics as mt
    return mt


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


@pytest.mark.parametrize("path", FILES, ids=lambda p: os.path.basename(p)[:-4])
def test_reference_fixtures(mt, path):
    g = np.load(path)
    out, tgt, time = dev(g["output"]), dev(g["target"]), int(g["time"])
    a1, a5, af = mt.accuracy_time(out, tgt, time, topk=(1, 5))
    np.testing.assert_allclose([a1.item(), a5.item()], g["acc_time"], rtol=1e-6)
    assert np.array_equal(af.cpu().numpy(), g["acc_time_top1"])
    r1, r5, rf = mt.recall_time(out, tgt, time, topk=(1, 5))
    np.testing.assert_allclose([r1.item(), r5.item()], g["rec_time"], rtol=1e-6)
    assert np.array_equal(rf.cpu().numpy(), g["rec_time_top1"])
    s1, _, sf = mt.accuracy_s(out, dev(g["label"]), topk=(1, 5))
    np.testing.assert_allclose(s1.item(), g["acc_s"][0], rtol=1e-6)
    assert np.array_equal(sf.cpu().numpy(), g["acc_s_top1"])
    m1, m5, mf = mt.accuracy(out, dev(g["multihot"]), topk=(1, 5))
    np.testing.assert_allclose([m1.item(), m5.item()], g["acc_m"], rtol=1e-6)
    assert np.array_equal(mf.cpu().numpy(), g["acc_m_top1"])


@pytest.mark.parametrize("rows,C,k", [(1, 1, 1), (7, 3, 5), (33, 157, 5), (64, 160, 8), (1000, 157, 1), (5, 1031, 8),
                                      (257, 4, 3), (19, 37, 2), (40, 300, 5), (40, 512, 8), (40, 513, 5), (300, 96, 5), (300, 256, 8)])
def test_frame_topk_matches_oracle(mt, rows, C, k):
    rng = np.random.default_rng(rows * 1000 + C)
    x = rng.standard_normal((rows, C)).astype(np.float32)
    got = mt.frame_topk(dev(x), k).cpu().numpy()
    assert np.array_equal(got, M.frame_topk(x, k))


def test_frame_topk_ties_nan_inf(mt):
    rng = np.random.default_rng(5)
    x = rng.integers(-3, 4, size=(513, 157)).astype(np.float32)  # heavy ties -> lower index first
    x[3, 10] = np.nan
    x[3, 150] = np.nan
    x[4, 7] = np.inf
    x[4, 9] = np.nan
    x[5, :] = -np.inf
    x[6, :] = 0.0
    x[6, 100] = -0.0
    got = mt.frame_topk(dev(x), 5).cpu().numpy()
    assert np.array_equal(got, M.frame_topk(x, 5))
    assert list(got[3, :2]) == [10, 150] and list(got[4, :2]) == [9, 7] and list(got[5]) == [0, 1, 2, 3, 4]
    assert list(got[6]) == [0, 1, 2, 3, 4]


def test_frame_topk_sample_major_and_strided(mt):
    rng = np.random.default_rng(6)
    x = rng.standard_normal((17, 6, 157)).astype(np.float32)
    got = mt.frame_topk(dev(x), 5, sample_major=True).cpu().numpy()
    assert got.shape == (6, 17, 5)
    assert np.array_equal(got, M.frame_topk(x.transpose(1, 0, 2), 5))
    # a sliced (non-16-byte-aligned rows) view
    xs = dev(x)[:, :, 1:150]
    got = mt.frame_topk(xs, 3).cpu().numpy()
    assert np.array_equal(got, M.frame_topk(x[:, :, 1:150], 3))


@pytest.mark.parametrize("B,T,C,Lt,k,dens", [(9, 40, 157, 12, 5, 0.03), (64, 256, 157, 32, 5, 0.01), (5, 7, 20, 30, 3, 0.2),
                                             (3, 50, 11, 1, 8, 0.5), (4, 3, 50, 9, 1, 0.1), (6, 90, 30, 64, 5, 0.05), (6, 200, 30, 80, 5, 0.03),
                                             (2, 33, 7000, 20, 2, 0.001)])
def test_batched_matching_matches_oracle(mt, B, T, C, Lt, k, dens):
    rng = np.random.default_rng(B * 31 + T)
    x = rng.standard_normal((T, B, C)).astype(np.float32)
    tgt = (rng.random((B, Lt, C)) < dens).astype(np.float32)
    time = rng.integers(0, Lt + 1, size=B).astype(np.int32)
    time[0] = Lt
    ca, na = mt.accuracy_time_batch(dev(x), dev(tgt), dev(time), k)
    cr, nr = mt.recall_time_batch(dev(x), dev(tgt), dev(time), k)
    ca, na, cr, nr = (t.cpu().numpy() for t in (ca, na, cr, nr))
    for b in range(B):
        pred = M.frame_topk(x[:, b], k)
        ea = M.accuracy_time_flags(pred, tgt[b], int(time[b]))
        er = M.recall_time_flags(pred, tgt[b], int(time[b]))
        assert np.array_equal(ca[b], ea), b
        assert np.array_equal(na[b], ea.sum(1)), b
        assert np.array_equal(cr[b][:, : time[b]], er), b
        assert not cr[b][:, time[b]:].any()
        assert np.array_equal(nr[b], er.sum(1)), b


def test_metrics_argument_errors(mt):
    from ctc_b200 import NbctcError
    x = torch.randn(4, 10, device="cuda")
    with pytest.raises(NbctcError):
        mt.frame_topk(x, 9)
    with pytest.raises(NbctcError):
        mt.frame_topk(torch.randn(4, 10), 2)
    assert mt.frame_topk(torch.empty(0, 10, device="cuda"), 2).shape == (0, 2)
    assert mt.frame_topk(x, 3).shape == (4, 3)
    assert (mt.frame_topk(torch.randn(4, 2, device="cuda"), 3)[:, 2] == -1).all()
