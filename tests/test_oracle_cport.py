"""The plain-C/OpenMP restatement (oracle/c) against the numpy oracle and the reference's goldens."""
import numpy as np
import pytest

from conftest import rel_l2
from helpers import make_bctc_case, make_ctc_case
from oracle import cport
from oracle import restatement as R


def test_cport_matches_reference_goldens(golden):
    kind = str(golden["kind"])
    out = cport.loss_grad(kind, golden["logits"], golden["targets"], golden["input_length"], golden["target_length"])
    assert abs(out["loss"] - float(golden["loss"])) < 1e-12 * abs(float(golden["loss"]))
    assert rel_l2(out["grad"], golden["grad"]) < 1e-12


@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_cport_matches_numpy_oracle(reduction):
    x, lab, il, tl = make_ctc_case(1, 70, 9, 40, 33, dup=True)
    a = cport.loss_grad("ctc", x, lab, il, tl, reduction)
    b = R.nbctc_loss_grad(x, lab, il, tl, reduction)
    np.testing.assert_allclose(a["loss"], b["loss"], rtol=1e-13)
    assert rel_l2(a["grad"], b["grad"]) < 1e-13
    xb, y, ilb, tlb = make_bctc_case(2, 50, 6, 30, 12, density=0.2)
    a = cport.loss_grad("bctc", xb, y, ilb, tlb, reduction)
    b = R.nbbctc_loss_grad(xb, y, ilb, tlb, reduction)
    np.testing.assert_allclose(a["loss"], b["loss"], rtol=1e-13)
    assert rel_l2(a["grad"], b["grad"]) < 1e-13


def test_cport_infeasible():
    x, lab, il, tl = make_ctc_case(3, 6, 3, 5, 8, ragged_T=False)
    tl[:] = [7, 2, 0]
    lab[:] = np.random.RandomState(0).randint(0, 5, size=lab.shape)
    out = cport.loss_grad("ctc", x, lab, il, tl, "none")
    assert np.isinf(out["per_seq"][0]) and np.isfinite(out["per_seq"][1]) and np.isinf(out["per_seq"][2])
    assert np.all(out["grad"][:, [0, 2]] == 0)
