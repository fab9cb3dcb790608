"""Randomised parity sweep of every single-label kernel and the multi-label paths against the C port of the oracle
(GPU box): shapes around the template boundaries, T_b close to L_b, peaked logits (on a class, on the path), repeated
labels, ragged lengths.     python tools/fuzz_parity.py [cases] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctc_b200  # noqa: E402
from oracle import cport  # noqa: E402

DEV = "cuda:0"
FLAGS = {"generic": 1, "lockstep": 8, "seqwarp": 32}


def gen(rs, binary):
    C = int(rs.choice([3, 17, 32, 33, 64, 100, 157, 200, 256, 260, 512, 1024] if not binary else [5, 33, 64, 100, 157, 256]))
    Lmax = int(rs.choice([1, 2, 7, 16, 31, 32, 33, 40, 64, 65, 100, 128, 129, 200, 256]))
    T = int(rs.choice([1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 33, 40, 64, 100, 130, 257]))
    T = max(T, int(rs.randint(1, 4)))
    B = int(rs.choice([1, 2, 3, 5, 8, 13]))
    if rs.uniform() < 0.04 and C <= 64:        # batches of more than one wave: the longest-first work queue
        B, T, Lmax = int(rs.choice([1300, 3100, 4500])), min(T, 17), min(Lmax, 100)
    if C >= 512:
        T, B = min(T, 64), min(B, 3)
    hi = min(Lmax, T)
    mode = rs.choice(["rand", "tight", "equal", "one"])
    tl = rs.randint(1, hi + 1, size=B)
    if mode == "equal":
        il = tl.copy()
    elif mode == "tight":
        il = np.minimum(T, tl + rs.randint(0, 3, size=B))
    elif mode == "one":
        tl = np.ones(B, dtype=np.int64)
        il = rs.randint(1, T + 1, size=B)
    else:
        il = np.array([rs.randint(l, T + 1) for l in tl])
    tl, il = tl.astype(np.int64), il.astype(np.int64)
    scale = float(rs.choice([1.0, 1.0, 4.0, 10.0] if binary else [1.0, 1.0, 4.0, 10.0, 40.0]))  # (multi-label: |x| > 100 is the documented BCE-clamp deviation)
    x = (rs.standard_normal((T, B, C)) * scale).astype(np.float32)
    peak = rs.choice(["none", "class", "path"])
    if binary:
        dens = float(rs.choice([0.02, 0.1, 0.3]))
        y = (rs.uniform(size=(B, Lmax, C)) < dens).astype(np.float32)
        y[np.arange(B)[:, None], np.arange(Lmax)[None, :], rs.randint(0, C, size=(B, Lmax))] = 1.0
        for b in range(B):
            y[b, tl[b]:] = 0.0
        return x, y, il, tl, dict(T=T, B=B, C=C, Lmax=Lmax, mode=mode, scale=scale)
    lab = rs.randint(0, C, size=(B, Lmax)).astype(np.int32)
    if rs.uniform() < 0.3:
        lab[:, 1::2] = lab[:, 0::2][:, : lab[:, 1::2].shape[1]]
    if rs.uniform() < 0.1:
        lab[:] = lab[:, :1]
    if peak == "class":
        x[:, :, 0] += float(rs.choice([8.0, 20.0]))
    for b in range(B):
        lab[b, tl[b]:] = -1
        if peak == "path":
            t_idx = np.arange(il[b])
            s_idx = np.minimum((t_idx * tl[b]) // max(il[b], 1), tl[b] - 1)
            x[t_idx, b, lab[b, s_idx]] += 12.0
    return x, lab, il, tl, dict(T=T, B=B, C=C, Lmax=Lmax, mode=mode, scale=scale, peak=str(peak))


def check(kind, x, tg, il, tl, flags):
    xt = torch.tensor(x, device=DEV, requires_grad=True)
    f = ctc_b200.no_blank_binary_ctc_loss if kind == "bctc" else ctc_b200.no_blank_ctc_loss
    per = f(xt, torch.tensor(tg, device=DEV), torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV), "none", flags=flags)
    per.sum().backward()
    torch.cuda.synchronize()
    ref = cport.loss_grad(kind, x, tg, il, tl, reduction="sum")
    per = per.detach().cpu().double().numpy()
    g = xt.grad.cpu().double().numpy()
    # 1e-5 relative, plus what float32 itself resolves: a row log-partition held in float32 (as the reference's
    # log_softmax holds it) is good to one ulp of the row's largest |logit|, the loss of T_b rows to T_b of those
    xmax = np.array([np.abs(x[:max(int(il[b]), 1), b]).max() for b in range(x.shape[1])], dtype=np.float32)
    res = il * np.spacing(xmax).astype(np.float64)
    err = np.abs(per - ref["per_seq"])
    le = float(np.max(err / np.maximum(np.abs(ref["per_seq"]), 1e-1 * il)))
    le_ok = bool(np.all(err <= 1e-5 * np.maximum(np.abs(ref["per_seq"]), 1e-1 * il) + res))
    den = max(np.linalg.norm(ref["grad"]), 1e-6)
    ge = float(np.linalg.norm(g - ref["grad"]) / den)
    ae = float(np.max(np.abs(g - ref["grad"])))
    T = x.shape[0]
    dead = np.arange(T)[:, None] >= il[None, :]
    tail = float(np.abs(g[dead]).max()) if dead.any() else 0.0
    ok = np.isfinite(per).all() and le_ok and (ge < 1e-5 or ae < 5e-6) and tail == 0.0
    return ok, le, ge, ae


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rs = np.random.RandomState(seed)
    bad = 0
    for i in range(n):
        binary = rs.uniform() < 0.3
        x, tg, il, tl, info = gen(rs, binary)
        kind = "bctc" if binary else "ctc"
        paths = [("default", 0), ("generic", 1)] if binary else list(FLAGS.items())
        for name, fl in paths:
            ok, le, ge, ae = check(kind, x, tg, il, tl, fl)
            if not ok:
                bad += 1
                print(f"FAIL case {i} {kind} {name} {info} il={il.tolist()} tl={tl.tolist()} loss {le:.1e} grad {ge:.1e} abs {ae:.1e}", flush=True)
    print(f"{n} cases, {bad} failures", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
