"""Pin the float64 oracle to the reference's own outputs (tests/golden/*.npz).

The fixtures were produced by importing /root/reference/NoBlankCTC.py and
NoBlankBinaryCTC.py unmodified (tests/golden/make_golden.py).  Known-answer constants
for the two fixture inputs held in the reference's scratch test file
(test.py:384-402 and test.py:258-273) are repeated literally below.
"""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from oracle import restatement as R


def _run(d, reduction="mean"):
    if str(d["kind"]) == "ctc":
        return R.nbctc_loss_grad(d["logits"], d["targets"], d["input_length"], d["target_length"], reduction)
    return R.nbbctc_loss_grad(d["logits"], d["targets"], d["input_length"], d["target_length"], reduction)


def test_oracle_matches_reference(golden):
    out = _run(golden)
    assert abs(out["loss"] - float(golden["loss"])) <= 1e-12 * abs(float(golden["loss"]))
    np.testing.assert_allclose(out["per_seq"], golden["per_seq"], rtol=1e-12, atol=0)
    assert rel_l2(out["grad"], golden["grad"]) < 1e-12
    np.testing.assert_allclose(out["grad"], golden["grad"], rtol=0, atol=1e-13)


def test_kat_constants():
    # reference float64 outputs on its own toy inputs (fp32-rounded logits)
    a = load_golden("kat_a_ctc")
    assert float(a["loss"]) == pytest.approx(6.623642288976711, rel=1e-14)
    b = load_golden("kat_b_bctc")
    assert float(b["loss"]) == pytest.approx(4.018208605043238, rel=1e-14)
    out = _run(a)
    np.testing.assert_allclose(out["grad"][0, 0], a["grad"][0, 0], atol=1e-14)
    # SURVEY 8(c) KAT-A constants were produced from float64 literals; fp32-rounded inputs agree to 1e-7
    np.testing.assert_allclose(a["grad"][0, 0],
                               [0.0615268443, 0.1848368551, -0.4248509427, 0.0112399407, 0.1672473027], atol=2e-7)


def test_gamma_rows_sum_to_one(golden):
    out = _run(golden)
    T = golden["logits"].shape[0]
    live = np.arange(T)[:, None] < golden["input_length"][None, :]
    s = out["gamma"].sum(axis=2)
    np.testing.assert_allclose(s[live], 1.0, atol=1e-12)
    assert np.all(s[~live] == 0.0)


def test_grad_zero_beyond_input_length(golden):
    out = _run(golden)
    T = golden["logits"].shape[0]
    dead = np.arange(T)[:, None] >= golden["input_length"][None, :]
    assert np.all(out["grad"][dead] == 0.0)


def test_reductions(golden):
    m = _run(golden, "mean")
    s = _run(golden, "sum")
    n = _run(golden, "none")
    B = golden["logits"].shape[1]
    assert s["loss"] == pytest.approx(m["loss"] * B, rel=1e-13)
    np.testing.assert_allclose(n["loss"], m["per_seq"], rtol=0, atol=0)
    np.testing.assert_allclose(s["grad"], m["grad"] * B, rtol=1e-13, atol=1e-15)


def test_closed_form_gradient_vs_autograd():
    """The closed-form gradient equals autograd through an independent torch alpha pass."""
    import torch
    rs = np.random.RandomState(5)
    T, B, C, L = 9, 3, 6, 4
    x = rs.standard_normal((T, B, C))
    lab = rs.randint(0, C, size=(B, L))
    tl = np.array([4, 2, 1])
    il = np.array([9, 5, 3])
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    lp = torch.log_softmax(xt, dim=2)
    total = 0.0
    for b in range(B):
        a = [lp[0, b, lab[b, 0]]] + [torch.tensor(-1e30, dtype=torch.float64)] * (tl[b] - 1)
        for t in range(1, il[b]):
            na = []
            for s in range(tl[b]):
                prev = a[s] if s == 0 else torch.logaddexp(a[s], a[s - 1])
                na.append(prev + lp[t, b, lab[b, s]])
            a = na
        total = total - a[tl[b] - 1]
    (total / B).backward()
    out = R.nbctc_loss_grad(x, lab, il, tl)
    assert float((total / B).detach()) == pytest.approx(out["loss"], rel=1e-13)
    np.testing.assert_allclose(out["grad"], xt.grad.numpy(), atol=1e-14)


def test_infeasible_sequences_are_inf():
    x = np.zeros((3, 2, 4))
    out = R.nbctc_loss_grad(x, np.array([[0, 1, 2, 3], [1, 1, 1, 1]]), np.array([3, 3]), np.array([4, 2]))
    assert np.isinf(out["per_seq"][0]) and np.isfinite(out["per_seq"][1])
    assert np.all(out["grad"][:, 0] == 0.0)


def test_best_path_kat_c():
    """SURVEY 8(c) KAT-C (restatement-defined; the reference has no alignment code)."""
    a = load_golden("kat_a_ctc")
    states, score = R.best_path(a["logits"], a["targets"], a["input_length"], a["target_length"])
    assert states.tolist() == [[0, 0, 1, 2], [0, 0, 0, 1]]
    assert R.frame_argmax(a["logits"]).T.tolist() == [[1, 4, 2, 3], [4, 1, 3, 2]]


def test_best_path_is_optimal_bruteforce():
    import itertools
    rs = np.random.RandomState(3)
    T, B, C, L = 6, 4, 5, 3
    x = rs.standard_normal((T, B, C)).astype(np.float32)
    lab = rs.randint(0, C, size=(B, L))
    tl = np.array([3, 2, 1, 3])
    il = np.array([6, 4, 2, 3])
    states, score = R.best_path(x, lab, il, tl)
    for b in range(B):
        best = -np.inf
        for adv in itertools.product([0, 1], repeat=il[b] - 1):
            path = np.concatenate([[0], np.cumsum(adv)])
            if path[-1] != tl[b] - 1:
                continue
            sc = sum(float(x[t, b, lab[b, path[t]]]) for t in range(il[b]))
            best = max(best, sc)
        assert score[b] == pytest.approx(best, abs=1e-9)
        p = states[b, : il[b]]
        assert p[0] == 0 and p[-1] == tl[b] - 1 and np.all(np.diff(p) >= 0) and np.all(np.diff(p) <= 1)
        assert np.all(states[b, il[b]:] == -1)
