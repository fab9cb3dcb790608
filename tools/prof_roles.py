#!/usr/bin/env python
"""Role profiler for the fused kernel (needs a build with NBCTC_EXTRA_NVCC_FLAGS=-DNBCTC_PROF).
Prints per-role cycle buckets averaged over the CTAs of one cfg-sized launch."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import ctc_b200
from ctc_b200 import _ffi

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
B, T, Cc, Lmax = w["B"], w["T"], w["C"], w["Lmax"]
dev = torch.device("cuda:0")
tg, il, tl = bench.make_inputs_np(w, 1234)
x = torch.randn((T, B, Cc), device=dev)
tgt, ilt, tlt = torch.tensor(tg, device=dev), torch.tensor(il, device=dev), torch.tensor(tl, device=dev)
lib = _ffi.lib()
lib.nbctc_debug_set_prof.argtypes = [C.c_void_p]
prof = torch.zeros((B, 6, 8), dtype=torch.int64, device=dev)
assert lib.nbctc_debug_set_prof(prof.data_ptr()) == 0
per = torch.empty(B, device=dev)
grad = torch.empty_like(x)
ws_bytes = int(lib.nbctc_workspace_bytes(T, B, Cc, Lmax, 0, 0))
ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
for it in range(3):
    rc = lib.nbctc_loss_grad_f32(x.data_ptr(), T, B, Cc, tgt.data_ptr(), Lmax, ilt.data_ptr(), tlt.data_ptr(),
                                 per.data_ptr(), None, None, grad.data_ptr(), None, 1.0 / B, ws.data_ptr(), ws_bytes, 0,
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == 0
torch.cuda.synchronize()
p = prof.cpu().numpy().astype(np.float64)
names = {
    "chain": ["wait pfull ph1", "alpha steps", "wait pfull ph2", "beta+replay steps", "wait gempty", "gamma pass", "-", "total"],
    "row": ["wait slot ph1", "forward tile", "wait slot ph2", "emit tile (A)", "wait gamma", "backward tile (B)", "-", "total"],
    "producer": ["wait sempty ph1", "wait sempty ph2", "-", "-", "-", "-", "-", "total"],
}
print(f"workload {sys.argv[1] if len(sys.argv) > 1 else 'cfg2'}: mean cycles per CTA (sequence)")
for role, sl in (("chain", p[:, 0]), ("row", p[:, 1:5].reshape(-1, 8)), ("producer", p[:, 5])):
    m = sl.mean(axis=0)
    print(f"  {role:9s} " + "  ".join(f"{n}={v:,.0f}" for n, v in zip(names[role], m) if n != "-"))
