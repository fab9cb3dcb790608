#!/usr/bin/env python
"""Golden fixtures of the auxiliary cross-entropy, produced by RUNNING THE REFERENCE'S OWN MODULES in
float64: /root/reference/CrossEntropy.py (unmodified; `.cuda()` shimmed to identity as in make_golden.py) and
torch.nn.CrossEntropyLoss, which models/__init__.py:85 instantiates.  The classified frame is input_length-1
(train.py:434 classifies v_output[temporal-1]).

    python tests/golden/make_golden_ce.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NBCTC_REFERENCE", "/root/reference")


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}")
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    from CrossEntropy import CrossEntropy
    torch.set_default_dtype(torch.float64)
    for name, (seed, T, B, C, scale) in {"ce_small": (1, 6, 4, 7, 1.0), "ce_charades": (2, 10, 8, 157, 1.0),
                                          "ce_wide": (3, 5, 3, 300, 4.0)}.items():
        rs = np.random.RandomState(seed)
        x = (rs.standard_normal((T, B, C)) * scale).astype(np.float32)
        il = rs.randint(1, T + 1, size=B).astype(np.int64)
        fi = il - 1
        y_idx = rs.randint(0, C, size=B).astype(np.int64)
        y_hot = (rs.uniform(size=(B, C)) < 0.05).astype(np.float32)
        y_hot[np.arange(B), y_idx] = 1.0
        out = {"logits": x, "input_length": il, "frame_index": fi, "y_index": y_idx, "y_multihot": y_hot}
        for mode in ("index", "multihot"):
            xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
            rows = xt[torch.tensor(fi), torch.arange(B)]
            if mode == "index":
                loss = torch.nn.CrossEntropyLoss()(rows, torch.tensor(y_idx))
            else:
                loss = CrossEntropy()(rows, torch.tensor(y_hot, dtype=torch.float64))
            loss.backward()
            out[f"loss_{mode}"] = float(loss)
            out[f"grad_{mode}"] = xt.grad.numpy().copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, out["loss_index"], out["loss_multihot"])
    torch.set_default_dtype(torch.float32)


if __name__ == "__main__":
    main()
