#!/usr/bin/env python
"""Times the metric kernels at the cfg2 shape (T=256, B=4096, C=157, Lt=32, k=5); CUDA events, L2-exceeding input."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctc_b200.metrics as mt

T, B, C, Lt, k = 256, 4096, 157, 32, 5
x = torch.randn(T, B, C, device="cuda")
tgt = (torch.rand(B, Lt, C, device="cuda") < 0.02).float()
time = torch.randint(1, Lt + 1, (B,), device="cuda", dtype=torch.int32)


def timeit(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


pred = mt.frame_topk(x, k, sample_major=True)
t_topk = timeit(lambda: mt.frame_topk(x, k, sample_major=True))
t_flat = timeit(lambda: mt.frame_topk(x, k))
t_match = timeit(lambda: mt.match_time(pred, tgt, time, recall=False))
t_rec = timeit(lambda: mt.match_time(pred, tgt, time, recall=True))
t_torch = timeit(lambda: x.topk(k, -1, True, True))
print(json.dumps({"shape": [T, B, C], "k": k, "topk_sample_major_ms": t_topk, "topk_flat_ms": t_flat,
                  "topk_GBps": x.numel() * 4 / t_flat / 1e6, "torch_topk_ms": t_torch,
                  "match_accuracy_ms": t_match, "match_recall_ms": t_rec}))
