#!/usr/bin/env python
"""Role profiler for the block-streaming kernel (needs a build with NBCTC_EXTRA_NVCC_FLAGS=-DNBCTC_PROF).
Prints per-role cycle buckets, averaged per warp and CTA, for one workload-sized launch.
Tuning knobs are read by the library from the environment: NBCTC_LPR, NBCTC_CTAS."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import ctc_b200
from ctc_b200 import _ffi

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = bench.WORKLOADS[name]
B, T, Cc, Lmax = w["B"], w["T"], w["C"], w["Lmax"]
nch = Cc // 4 if Cc % 4 == 0 else (Cc + 6) // 4
LPR = int(os.environ.get("NBCTC_LPR", "0")) or (4 if nch <= 16 else 8 if nch <= 64 else 32)
GB = 32 // LPR
NRW = 2 if Lmax > 128 else 8            # row warps = time steps per tile
NCH = GB * (2 if Lmax > 32 else 1)      # chain warps
dev = torch.device("cuda:0")
tg, il, tl = bench.make_inputs_np(w, 1234)
x = torch.randn((T, B, Cc), device=dev)
tgt, ilt, tlt = torch.tensor(tg, device=dev), torch.tensor(il, device=dev), torch.tensor(tl, device=dev)
lib = _ffi.lib()
lib.nbctc_debug_set_prof.argtypes = [C.c_void_p]
prof = torch.zeros(24 + 160 * 32 * 2, dtype=torch.int64, device=dev)
per = torch.empty(B, device=dev)
grad = torch.empty_like(x)
ws_bytes = int(lib.nbctc_workspace_bytes(T, B, Cc, Lmax, 0, 0))
ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)


def run():
    rc = lib.nbctc_loss_grad_f32(x.data_ptr(), T, B, Cc, tgt.data_ptr(), Lmax, ilt.data_ptr(), tlt.data_ptr(),
                                 per.data_ptr(), None, None, grad.data_ptr(), None, 1.0 / B, ws.data_ptr(), ws_bytes, 0,
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.nbctc_last_error()


for _ in range(3):
    run()
torch.cuda.synchronize()
assert lib.nbctc_debug_set_prof(prof.data_ptr()) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
pall = prof.cpu().numpy().astype(np.float64)
p = pall[:24].reshape(3, 8)
trace = pall[24:].reshape(160, 32, 2)
groups = (B + GB - 1) // GB
names = {
    0: ("chain", NCH, ["phase1 steps", "phase2 steps", "-", "-", "-", "-", "barrier", "total"]),
    1: ("row", NRW, ["wait rows", "forward step", "-", "-", "emit (ph2)", "scatter", "barrier", "total"]),
    2: ("mover", 2, ["store issue", "wait store reads", "load issue", "-", "-", "-", "barrier", "total"]),
}
print(f"workload {name}: GB={GB} NRW={NRW} groups={groups} kernel {e0.elapsed_time(e1):.3f} ms (instrumented); "
      f"mean cycles per warp per CTA")
for role, (rn, nw, nm) in names.items():
    m = p[role] / (groups * nw)
    print(f"  {rn:9s} " + "  ".join(f"{n}={v:,.0f}" for n, v in zip(nm, m) if n != "-"))

# per-iteration trace of the middle CTA: cycles each role spends working in the iteration (work end - previous barrier end)
nw = NCH + NRW + 2
prev = np.zeros(nw)
print("trace of one CTA: iteration | iteration cycles | busy cycles chain(max) rows(max)")
for itx in range(160):
    if trace[itx, :nw, 1].max() == 0:
        break
    busy = trace[itx, :nw, 0] - prev
    end = trace[itx, :nw, 1]
    print(f"  it={itx:3d} iter={end.max() - prev.max():7.0f}  chain={busy[:NCH].max():6.0f} rows={busy[NCH:NCH + NRW].max():6.0f} movers={busy[NCH + NRW:].max():6.0f}")
    prev = end
