import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctc_b200
from ctc_b200 import _ffi
from oracle import cport
from pipe_check import case
x, lab, il, tl = case(2, 37, 5, 157, 20)
dev = torch.device("cuda:0")
xt = torch.tensor(x, device=dev, requires_grad=True)
loss = ctc_b200.no_blank_ctc_loss(xt, torch.tensor(lab, device=dev), torch.tensor(il, device=dev), torch.tensor(tl, device=dev), "none", flags=_ffi.FLAG_SEQWARP)
loss.sum().backward()
g = xt.grad.cpu().numpy().astype(np.float64)
ref = cport.loss_grad("ctc", x, lab, il, tl, reduction="none")
bad = ~np.isfinite(g)
print("il", il, "tl", tl)
idx = np.argwhere(bad)
print("nonfinite count", len(idx))
for t, b, c in idx[:40]:
    s = [i for i in range(tl[b]) if lab[b, i] == c]
    print(t, b, c, g[t, b, c], "states", s)
d = np.abs(np.where(bad, 0, g) - ref["grad"])
print("max err finite", d.max(), np.unravel_index(d.argmax(), d.shape))
for b in range(5):
    print(b, "err by t", [f"{d[t, b].max():.1e}" for t in range(0, 37, 3)])
