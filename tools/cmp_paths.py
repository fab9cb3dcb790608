#!/usr/bin/env python
"""Time the fused kernel against the generic three-kernel path on a bench workload (kernel time, CUDA events)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ctc_b200 import _ffi

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
w = bench.WORKLOADS[name]
B, T, C, Lmax = w["B"], w["T"], w["C"], w["Lmax"]
dev = torch.device("cuda:0")
tg, il, tl = bench.make_inputs_np(w, 1234)
x = torch.randn((T, B, C), device=dev)
tgt, ilt, tlt = torch.tensor(tg, device=dev), torch.tensor(il, device=dev), torch.tensor(tl, device=dev)
lib = _ffi.lib()
per = torch.empty(B, device=dev)
grad = torch.empty_like(x)
keep = {}
only_fused = os.environ.get('CMP_ONLY_FUSED') == '1'
reps = int(os.environ.get('CMP_REPS', '3'))
for flags, label in ((0, "fused"), (1, "generic")):
    ws_bytes = int(lib.nbctc_workspace_bytes(T, B, C, Lmax, 1 if w['binary'] else 0, flags))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    def run():
        fn = lib.nbbctc_loss_grad_f32 if w['binary'] else lib.nbctc_loss_grad_f32
        rc = fn(x.data_ptr(), T, B, C, tgt.data_ptr(), Lmax, ilt.data_ptr(), tlt.data_ptr(),
                                     per.data_ptr(), None, None, grad.data_ptr(), None, 1.0 / B, ws.data_ptr(), ws_bytes,
                                     flags, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.nbctc_last_error()
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ab = bench.algorithmic_bytes(w, B)
    print(f"{name} {label}: {ms:.3f} ms  ({ab / ms / 1e6 / 6540.2 * 100:.1f}% of HBM roofline)  loss_sum={float(per.double().sum()):.3f}  ws={ws_bytes/1e9:.2f} GB")
    del ws
    keep[label] = grad.clone() if x.numel() < (1 << 29) else None
    if only_fused:
        break
if keep.get('generic') is not None:
    print(f"{name} max |grad fused - generic| = {float((keep['fused'] - keep['generic']).abs().max()):.3e}")
