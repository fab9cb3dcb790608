"""Kernel time of one path on a bench workload (CUDA events, C ABI, device-resident inputs).

    python tools/time_paths.py cfg2|cfg5|cfg2r [seqwarp|lockstep|generic] [reps]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctc_b200 import _ffi  # noqa: E402

SHAPES = {"cfg2": (256, 4096, 157, 32, False), "cfg2r": (256, 4096, 157, 32, True), "cfg5": (512, 8192, 157, 64, False),
          "cfg1": (64, 8, 157, 8, False), "cfg4": (4096, 1024, 1024, 256, False), "cfg4q": (1024, 1024, 1024, 256, False), "cfg4s": (512, 1024, 1024, 256, False), "cfg5s": (512, 4096, 157, 64, False), "b1024": (256, 1024, 157, 32, False), "b2048": (256, 2048, 157, 32, False), "b3072": (256, 3072, 157, 32, False),
          "b1536": (256, 1536, 157, 32, False), "w256": (1024, 256, 1024, 256, False), "w512": (1024, 512, 1024, 256, False), "w768": (1024, 768, 1024, 256, False),
          "m1024": (512, 1024, 157, 64, False), "m2048": (512, 2048, 157, 64, False), "m3072": (512, 3072, 157, 64, False)}


def main():
    name = sys.argv[1]
    path = sys.argv[2] if len(sys.argv) > 2 else "seqwarp"
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    T, B, C, Lmax, ragged = SHAPES[name]
    flags = {"seqwarp": _ffi.FLAG_SEQWARP, "lockstep": _ffi.FLAG_LOCKSTEP, "generic": _ffi.FLAG_GENERIC}[path] | _ffi.FLAG_ALIGNED16
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(T, B, C, device=dev, generator=g)
    rs = np.random.RandomState(2)
    tl = rs.randint(1, Lmax + 1, size=B).astype(np.int64)
    il = np.array([rs.randint(max(l, T // 2), T + 1) for l in tl], dtype=np.int64) if ragged else np.full(B, T, dtype=np.int64)
    lab = rs.randint(0, C, size=(B, Lmax)).astype(np.int32)
    labt, ilt, tlt = torch.tensor(lab, device=dev), torch.tensor(il, device=dev), torch.tensor(tl, device=dev)
    grad = torch.empty_like(x)
    per = torch.empty(B, device=dev)
    lib = _ffi.lib()
    wsb = int(lib.nbctc_workspace_bytes(T, B, C, Lmax, 0, flags))
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def call():
        rc = lib.nbctc_loss_grad_f32(x.data_ptr(), T, B, C, labt.data_ptr(), Lmax, ilt.data_ptr(), tlt.data_ptr(), per.data_ptr(),
                                     None, None, grad.data_ptr(), None, 1.0 / B, ws.data_ptr(), wsb, flags, st)
        _ffi.check(rc, "nbctc_loss_grad_f32")

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = 2 * 4 * T * B * C
    print(f"{name} {path}: {ms:.4f} ms  {alg / ms / 1e6:.0f} GB/s algorithmic = {alg / ms / 1e6 / 6540.2:.3f} of 6540 GB/s; loss mean {float(per.mean()):.4f}", flush=True)


if __name__ == "__main__":
    main()
