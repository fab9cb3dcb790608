#!/usr/bin/env python
"""Multi-label loss at growing logit scales (tiled path and generic kernels) against the float64 oracle.
Run on a GPU box: python tools/extreme_inputs.py"""
import sys, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ctc_b200
from helpers import make_bctc_case
from oracle import restatement as R
dev = torch.device('cuda:0')
for scale in (1.0, 4.0, 8.0, 15.0):
    for (T, B, C, L, dens) in ((300, 6, 157, 20, 0.03), (64, 4, 8, 6, 0.3), (500, 3, 64, 32, 0.05)):
        x, y, il, tl = make_bctc_case(11, T, B, C, L, density=dens)
        x = (x * scale).astype(np.float32)
        xt = torch.tensor(x, device=dev, requires_grad=True)
        out = {}
        for flags, name in ((0, 'tiled'), (1, 'generic')):
            xt.grad = None
            loss = ctc_b200.NoBlankBinaryCTC(flags=flags)(xt, torch.tensor(y, device=dev), torch.tensor(il, device=dev), torch.tensor(tl, device=dev))
            loss.backward()
            out[name] = (float(loss.detach()), xt.grad.detach().cpu().numpy().astype(np.float64))
        ref = R.nbbctc_loss_grad(x, y, il, tl)
        for name, (l, g) in out.items():
            rl = abs(l - ref['loss']) / abs(ref['loss'])
            rg = np.linalg.norm(g - ref['grad']) / np.linalg.norm(ref['grad'])
            print(f"scale {scale:5.1f} T{T} C{C} L{L} {name:8s} rel_loss {rl:.2e} rel_grad {rg:.2e} finite {np.isfinite(g).all()}")
