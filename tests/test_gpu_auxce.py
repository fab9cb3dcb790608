"""GPU parity of the auxiliary cross-entropy: nbctc_aux_ce_f32 through ctc_b200.ctc_plus_cross_entropy
against the float64 oracle (oracle/restatement.py: aux_ce, pinned to the reference's CrossEntropy.py / nn.CrossEntropyLoss
by tests/test_oracle_ce.py) and against the reference-generated fixtures."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_l2
from helpers import make_bctc_case, make_ctc_case
from oracle import restatement as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def _run(x, lab, il, tl, ce_t, alpha, fi=None, binary=False):
    import ctc_b200
    xt = torch.tensor(x, device=DEV, requires_grad=True)
    labt = torch.tensor(lab, device=DEV)
    total, ctc, ce = ctc_b200.ctc_plus_cross_entropy(
        xt, labt, torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV), torch.tensor(ce_t, device=DEV), alpha,
        frame_index=None if fi is None else torch.tensor(fi, device=DEV), binary=binary)
    total.backward()
    torch.cuda.synchronize()
    return float(total), float(ctc), float(ce), xt.grad.cpu().double().numpy()


@pytest.mark.parametrize("mode", ["index", "multihot"])
@pytest.mark.parametrize("shape", [(12, 5, 7, 4), (64, 8, 157, 8), (100, 33, 157, 32), (40, 3, 300, 40)])
def test_ctc_plus_ce_matches_oracle(mode, shape):
    T, B, C, Lmax = shape
    x, lab, il, tl = make_ctc_case(11, T, B, C, Lmax)
    rs = np.random.RandomState(5)
    y_idx = rs.randint(0, C, size=B).astype(np.int64)
    y_hot = (rs.uniform(size=(B, C)) < 0.05).astype(np.float32)
    y_hot[np.arange(B), y_idx] = 1.0
    tgt = y_idx if mode == "index" else y_hot
    alpha = 0.7
    total, ctc, ce, grad = _run(x, lab, il, tl, tgt, alpha)
    ref_ctc = R.nbctc_loss_grad(x, lab, il, tl, "mean")
    ref_ce = R.aux_ce(x, tgt, il - 1, mode)
    assert abs(ctc - ref_ctc["loss"]) < TOL * abs(ref_ctc["loss"])
    assert abs(ce - ref_ce["loss"]) < TOL * abs(ref_ce["loss"])
    assert abs(total - (ref_ctc["loss"] + alpha * ref_ce["loss"])) < TOL * abs(total)
    ref_grad = ref_ctc["grad"] + alpha * ref_ce["grad"]
    assert rel_l2(grad, ref_grad) < TOL
    assert np.max(np.abs(grad - ref_grad)) < TOL * np.max(np.abs(ref_grad))


def test_binary_ctc_plus_ce_explicit_frames():
    T, B, C, Lmax = 50, 6, 157, 9
    x, y, il, tl = make_bctc_case(3, T, B, C, Lmax)
    rs = np.random.RandomState(9)
    fi = np.array([rs.randint(0, t) for t in il], dtype=np.int64)
    y_hot = (rs.uniform(size=(B, C)) < 0.04).astype(np.float32)
    total, ctc, ce, grad = _run(x, y, il, tl, y_hot, 0.25, fi=fi, binary=True)
    ref_ctc = R.nbbctc_loss_grad(x, y, il, tl, "mean")
    ref_ce = R.aux_ce(x, y_hot, fi, "multihot")
    ref_grad = ref_ctc["grad"] + 0.25 * ref_ce["grad"]
    assert abs(total - (ref_ctc["loss"] + 0.25 * ref_ce["loss"])) < TOL * abs(total)
    assert rel_l2(grad, ref_grad) < TOL


@pytest.mark.parametrize("name", ["ce_small", "ce_charades", "ce_wide"])
@pytest.mark.parametrize("mode", ["index", "multihot"])
def test_aux_ce_c_abi_against_reference_fixtures(name, mode):
    """The C entry point alone (gradient ADDED into a zeroed buffer) against what the reference's modules produced."""
    from ctc_b200 import _ffi
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    x = torch.tensor(z["logits"], device=DEV)
    T, B, C = x.shape
    fi = torch.tensor(z["frame_index"], device=DEV)
    grad = torch.zeros_like(x)
    ce = torch.empty(B, device=DEV)
    yi = torch.tensor(z["y_index"], device=DEV).int() if mode == "index" else None
    yh = torch.tensor(z["y_multihot"], device=DEV) if mode == "multihot" else None
    rc = _ffi.lib().nbctc_aux_ce_f32(x.data_ptr(), T, B, C, fi.data_ptr(), None, None if yi is None else yi.data_ptr(),
                                     None if yh is None else yh.data_ptr(), 1.0 / B, None, ce.data_ptr(), grad.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream)
    _ffi.check(rc, "nbctc_aux_ce_f32")
    torch.cuda.synchronize()
    loss = float(ce.double().mean())
    assert abs(loss - float(z[f"loss_{mode}"])) < TOL * abs(float(z[f"loss_{mode}"]))
    assert rel_l2(grad.cpu().numpy(), z[f"grad_{mode}"]) < TOL


def test_aux_ce_rejects_both_target_forms():
    from ctc_b200 import _ffi
    x = torch.zeros(2, 2, 4, device=DEV)
    il = torch.ones(2, dtype=torch.int64, device=DEV)
    ce = torch.empty(2, device=DEV)
    rc = _ffi.lib().nbctc_aux_ce_f32(x.data_ptr(), 2, 2, 4, None, il.data_ptr(), None, None, 1.0, None, ce.data_ptr(), None, None)
    assert rc != 0
