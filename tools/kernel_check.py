"""Development check of one single-label kernel against the CPU oracle port on a few shapes (GPU box).

    python tools/kernel_check.py [quick|full] [seqwarp|lockstep]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctc_b200  # noqa: E402
from ctc_b200 import _ffi  # noqa: E402
from oracle import cport  # noqa: E402


DEFAULT_FLAGS = _ffi.FLAG_SEQWARP


def case(seed, T, B, C, Lmax, ragged=True, boost=0.0, dup=False, Lmin=1):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((T, B, C)).astype(np.float32)
    if boost:
        x[:, :, 0] += boost
    hi = min(Lmax, T)
    tl = rs.randint(min(Lmin, hi), hi + 1, size=B).astype(np.int64)
    il = np.array([rs.randint(max(l, (T + 1) // 2), T + 1) for l in tl], dtype=np.int64) if ragged else np.full(B, T, dtype=np.int64)
    lab = rs.randint(0, C, size=(B, Lmax)).astype(np.int32)
    if dup:
        lab[:, 1::2] = lab[:, 0::2][:, : lab[:, 1::2].shape[1]]
    for b in range(B):
        lab[b, tl[b]:] = -1
    return x, lab, il, tl


def run(name, x, lab, il, tl, flags=None, want_grad=True):
    flags = DEFAULT_FLAGS if flags is None else flags
    dev = torch.device("cuda:0")
    xt = torch.tensor(x, device=dev, requires_grad=want_grad)
    t0 = time.time()
    with torch.set_grad_enabled(want_grad):
        loss = ctc_b200.no_blank_ctc_loss(xt, torch.tensor(lab, device=dev), torch.tensor(il, device=dev),
                                          torch.tensor(tl, device=dev), "none", flags=flags)
        if want_grad:
            loss.sum().backward()
    torch.cuda.synchronize()
    dt = time.time() - t0
    ref = cport.loss_grad("ctc", x, lab, il, tl, reduction="none")
    per = loss.detach().cpu().numpy().astype(np.float64)
    rl = np.abs(per - ref["per_seq"]).max() / np.abs(ref["per_seq"]).max()
    msg = f"{name}: loss rel {rl:.2e}"
    ok = rl < 1e-5
    if want_grad:
        g = xt.grad.cpu().numpy().astype(np.float64)
        rg = np.linalg.norm(g - ref["grad"]) / np.linalg.norm(ref["grad"])
        li = np.abs(g - ref["grad"]).max() / np.abs(ref["grad"]).max()
        # rows beyond input_length are exact zeros
        T = x.shape[0]
        tail = np.arange(T)[:, None] >= il[None, :]
        z = float(np.abs(g[tail]).max()) if tail.any() else 0.0
        msg += f" grad L2 {rg:.2e} Linf {li:.2e} tail {z:.1e}"
        ok = ok and rg < 1e-5 and li < 1e-5 and z == 0.0
    print(msg + f" ({dt * 1e3:.0f} ms) {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def main():
    global DEFAULT_FLAGS
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    if len(sys.argv) > 2 and sys.argv[2] == "lockstep":
        DEFAULT_FLAGS = _ffi.FLAG_LOCKSTEP
    if len(sys.argv) > 2 and sys.argv[2] == "seqwarp":
        DEFAULT_FLAGS = _ffi.FLAG_SEQWARP
    print("sequence-per-warp kernel;" if DEFAULT_FLAGS == _ffi.FLAG_SEQWARP else "lock-step kernel;", torch.cuda.get_device_name(0), flush=True)
    ok = True
    ok &= run("tiny T=4 B=2", *case(0, 4, 2, 157, 3, ragged=False))
    ok &= run("cfg1-like", *case(1, 64, 8, 157, 8, ragged=False))
    ok &= run("ragged T=37 B=5", *case(2, 37, 5, 157, 20))
    ok &= run("ragged T=100 B=33 dup", *case(3, 100, 33, 150, 32, dup=True))
    ok &= run("T=1..", *case(4, 3, 9, 157, 2))
    ok &= run("L<=64", *case(5, 130, 19, 157, 64))
    ok &= run("peaked 14 L<=32", *case(6, 256, 16, 157, 32, ragged=False, boost=14.0, Lmin=20))
    ok &= run("peaked 14 L<=64", *case(7, 512, 8, 157, 64, ragged=False, boost=14.0, Lmin=50))
    ok &= run("no grad", *case(8, 64, 12, 157, 32), want_grad=False)
    ok &= run("dup heavy", *case(14, 70, 40, 6, 32, dup=True))
    ok &= run("peaked 30 L<=32", *case(15, 256, 16, 157, 32, ragged=False, boost=30.0, Lmin=20))
    ok &= run("C=33 odd", *case(16, 50, 37, 33, 17))
    ok &= run("C=256", *case(17, 50, 20, 256, 40))
    if mode == "full":
        ok &= run("C=1024 L<=256", *case(9, 600, 6, 1024, 256))
        ok &= run("peaked 8 L<=256", *case(10, 1024, 4, 1024, 256, ragged=False, boost=8.0, Lmin=200))
        ok &= run("C=512 L<=100 dup", *case(19, 77, 9, 512, 100, dup=True))
        ok &= run("C=160 L<=128", *case(20, 300, 7, 160, 128))
        ok &= run("wide multi-wave B=1500", *case(21, 24, 1500, 64, 70))
        ok &= run("multi-wave ragged B=9000", *case(18, 40, 9000, 40, 12))
        ok &= run("cfg2 quarter", *case(11, 256, 1024, 157, 32, ragged=False))
        ok &= run("cfg2 ragged", *case(12, 256, 1024, 157, 32, ragged=True))
        ok &= run("cfg5 slice", *case(13, 512, 512, 157, 64, ragged=False))
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
