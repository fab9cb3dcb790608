#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Only runs in the build container (needs /root/reference, which is not shipped to the
GPU box).  The reference modules are imported unmodified; the only shims are
  * ``Tensor.cuda`` / ``Module.cuda`` -> identity (the reference hard-codes .cuda(),
    NoBlankCTC.py:40-41,49,61-63,73,78,92,96), and
  * ``torch.set_default_dtype(torch.float64)`` so its internally allocated ``prob``
    tensor (NoBlankCTC.py:96) is float64.
Outputs: one ``<case>.npz`` per case holding the inputs and the reference's
loss / per-sequence loss / d(loss)/d(logits) in float64.

    python tests/golden/make_golden.py            # regenerates every *.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NBCTC_REFERENCE", "/root/reference")


def _load_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}")
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    import NoBlankCTC as m1
    import NoBlankBinaryCTC as m2
    return m1.NoBlankCTC, m2.NoBlankBinaryCTC


def run_ref(mod_cls, logits, targets, in_len, tgt_len):
    """Run the reference in float64; also return per-sequence losses (B one-hot-weighted calls
    would be slow, so per-sequence values come from single-sequence batches)."""
    torch.set_default_dtype(torch.float64)
    x = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    tg = torch.tensor(targets)
    if tg.dtype.is_floating_point:
        tg = tg.double()
    il = torch.tensor(in_len, dtype=torch.int64)
    tl = torch.tensor(tgt_len, dtype=torch.int64)
    m = mod_cls()
    loss = m(x, tg, il, tl)
    loss.backward()
    grad = x.grad.detach().numpy().copy()
    B = x.shape[1]
    per_seq = np.zeros(B)
    with torch.no_grad():
        for b in range(B):
            per_seq[b] = float(mod_cls()(x[:, b:b + 1].detach(), tg[b:b + 1], il[b:b + 1], tl[b:b + 1]))
    torch.set_default_dtype(torch.float32)
    return float(loss), per_seq, grad


def rand_case(seed, T, B, C, Lmax, binary=False, ragged_T=True, density=0.2, dup=False):
    rs = np.random.RandomState(seed)
    logits = rs.standard_normal((T, B, C)).astype(np.float32)
    tgt_len = rs.randint(1, Lmax + 1, size=B)
    tgt_len[0] = Lmax
    if B > 1:
        tgt_len[1] = 1
    if ragged_T:
        in_len = np.array([rs.randint(max(l, 1), T + 1) for l in tgt_len])
        in_len[0] = T
        if B > 2:
            in_len[2] = tgt_len[2]          # L_b == T_b: single admissible path
    else:
        in_len = np.full(B, T)
    if binary:
        y = (rs.uniform(size=(B, Lmax, C)) < density).astype(np.float32)
        for b in range(B):
            for s in range(Lmax):
                if y[b, s].sum() == 0:
                    y[b, s, rs.randint(C)] = 1.0
                if s >= tgt_len[b]:
                    y[b, s] = 0.0           # padded rows (reference pads -1; BCELoss rejects it on CPU)
        targets = y
    else:
        lab = rs.randint(0, C, size=(B, Lmax)).astype(np.int32)
        if dup:
            lab[:, 1::2] = lab[:, 0::2][:, : lab[:, 1::2].shape[1]]   # repeated labels are distinct states
        for b in range(B):
            lab[b, tgt_len[b]:] = -1
        targets = lab
    return logits, targets, in_len.astype(np.int64), tgt_len.astype(np.int64)


def main():
    NoBlankCTC, NoBlankBinaryCTC = _load_reference()
    cases = {}

    # KAT-A: fixture inputs from reference test.py:384-402 (fed as logits)
    yseq = np.array([[[1.2, 2.3, 1.4, -0.5, 2.2], [-0.1, 1.2, 0.4, 2.5, 3.2]],
                     [[0.5, 1.3, 2.2, 0.1, 2.4], [1.1, 2.2, 0.7, 1.4, 2.2]],
                     [[0.8, -1.5, 2.3, 1.2, 2.1], [0.9, 1.4, 0.6, 2.3, 1.0]],
                     [[0.2, -1.0, 1.3, 2.2, 0.1], [0.2, 1.0, 1.6, 1.3, 1.2]]], dtype=np.float32)
    cases["kat_a_ctc"] = ("ctc", yseq, np.array([[2, 3, 4], [1, 2, 0]], dtype=np.int32),
                          np.array([4, 4]), np.array([3, 2]))
    # KAT-B: fixture inputs from reference test.py:258-273 (binary variant, one-hot rows)
    yseq_b = yseq[:, :1].copy()
    lab_b = np.array([[[0, 0, 1, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 1]]], dtype=np.float32)
    cases["kat_b_bctc"] = ("bctc", yseq_b, lab_b, np.array([4]), np.array([3]))

    cases["rand_ctc_small"] = ("ctc",) + rand_case(11, T=12, B=4, C=7, Lmax=5)
    cases["rand_ctc_dup"] = ("ctc",) + rand_case(12, T=10, B=3, C=5, Lmax=6, dup=True)
    cases["rand_ctc_fullT"] = ("ctc",) + rand_case(13, T=16, B=4, C=33, Lmax=8, ragged_T=False)
    cases["rand_ctc_wide"] = ("ctc",) + rand_case(14, T=40, B=2, C=19, Lmax=36)
    cases["rand_bctc_small"] = ("bctc",) + rand_case(21, T=12, B=4, C=7, Lmax=5, binary=True)
    cases["rand_bctc_dense"] = ("bctc",) + rand_case(22, T=9, B=3, C=11, Lmax=4, binary=True, density=0.5)
    cases["rand_bctc_wide"] = ("bctc",) + rand_case(23, T=40, B=2, C=19, Lmax=34, binary=True, density=0.1)
    # cfg1 of BASELINE.json (the reference's own CPU-runnable case)
    cases["cfg1_ctc"] = ("ctc",) + rand_case(1234, T=64, B=8, C=157, Lmax=8, ragged_T=False)
    cases["cfg1_bctc"] = ("bctc",) + rand_case(1235, T=64, B=8, C=157, Lmax=8, binary=True,
                                               ragged_T=False, density=0.03)

    for name, (kind, logits, targets, in_len, tgt_len) in cases.items():
        cls = NoBlankCTC if kind == "ctc" else NoBlankBinaryCTC
        loss, per_seq, grad = run_ref(cls, logits, targets, in_len, tgt_len)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind=kind, logits=logits,
                            targets=targets, input_length=np.asarray(in_len, dtype=np.int64),
                            target_length=np.asarray(tgt_len, dtype=np.int64),
                            loss=np.float64(loss), per_seq=per_seq, grad=grad)
        print(f"{name:18s} kind={kind} shape={logits.shape} loss={loss!r}")


if __name__ == "__main__":
    main()
