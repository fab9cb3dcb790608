"""Numpy emulation of the pipeline kernel's chain stage (pipe_kernel.cuh: chain_sequence), lane by lane.

Development tool: checks the meet-in-the-middle alpha/beta recursion with per-lane block exponents against the
float64 oracle on the CPU before the CUDA version runs on a GPU.  Not imported by the product path.

    python tools/emul_chain.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import restatement as R  # noqa: E402

SENT = -(1 << 28)
RB = 8
PMIN = 2.0 ** -120


def pow2z(e):
    if e < -1022:
        return 0.0
    return float(np.ldexp(1.0, min(e, 1023)))


def expo(v):
    """exponent field - 1023 of the largest entry (None for an all-zero / denormal lane)."""
    m = float(np.max(v))
    if m < 2.0 ** -1022:
        return None
    return int(np.floor(np.log2(m))) if np.isfinite(m) else None


def chain(p, Tb, Lb, NS, wgt=1.0):
    """p: (T, Lpad) emissions (float32 values as float64).  Returns loss, gamma' (T, Lpad) = -w*gamma."""
    W = 16
    Lpad = W * NS
    T = p.shape[0]
    gam = np.zeros((T, Lpad))
    if Tb == 1:
        z = p[0, 0]
        gam[0, 0] = -wgt
        return -np.log(z), gam
    Ha = (Tb + 1) // 2
    odd = Tb & 1
    nblk = (Ha + RB - 1) // RB
    # state index of (dir, lane, j)
    sidx = np.zeros((2, W, NS), dtype=int)
    for hl in range(W):
        for j in range(NS):
            sidx[0, hl, j] = hl * NS + j
            sidx[1, hl, j] = Lpad - 1 - (hl * NS + j)
    x = np.zeros((2, W, NS))
    e = np.zeros((2, W), dtype=int)
    fac = np.ones((2, W))
    fac[:, 0] = 0.0
    x[0][sidx[0] == 0] = 1.0
    x[1][sidx[1] == Lb - 1] = 1.0
    firstp = [True, True]
    AB = np.zeros((T, 2, W, NS))          # storer-position order, indexed [t][dir of storer]
    E = np.zeros((2, nblk, W), dtype=int)

    def t_first(d, k):
        return k if d == 0 else Tb - 1 - k + odd

    def t_second(d, k):
        return Ha + k - odd if d == 0 else Ha - 1 - k

    DEC = {2: 208, 4: 420, 8: 850, 16: 850}[NS]

    def block_entry(store_blk=None):
        for d in range(2):
            A = np.full(W, SENT, dtype=np.int64)
            for hl in range(W):
                ex = expo(x[d, hl])
                if ex is not None:
                    A[hl] = e[d, hl] + ex
            # one-sided envelope: e_l = max_{k<=l} (A_k - DEC*(l-k)), log-step scan as on the device
            env = A.copy()
            o = 1
            while o < W:
                sh = np.full(W, SENT, dtype=np.int64)
                sh[o:] = env[:-o]
                cand = np.where(sh > SENT // 2, sh - DEC * o, SENT)
                env = np.maximum(env, cand)
                o *= 2
            enew = np.where(env > SENT // 2, env, e[d])
            for hl in range(W):
                x[d, hl] *= pow2z(int(e[d, hl] - enew[hl]))
            e[d] = enew
            for hl in range(1, W):
                fac[d, hl] = pow2z(int(e[d, hl - 1] - e[d, hl]))
            if store_blk is not None:
                E[d, store_blk] = e[d]

    def load_p(d, t):
        pj = np.where(sidx[d] < Lb, p[t][sidx[d]], 0.0)
        if firstp[d]:
            start = 0 if d == 0 else Lb - 1
            pj = np.where(sidx[d] == start, pj, 0.0)
        return pj

    def step(d, pj, second):
        up = np.zeros(W)
        up[1:] = x[d, :-1, NS - 1]
        up[0] = x[d, 0, NS - 1]
        sums = None
        if second:
            sums = np.empty((W, NS))
            sums[:, 1:] = x[d, :, 1:] + x[d, :, :-1]
            sums[:, 0] = up * fac[d] + x[d, :, 0]
        xn = np.empty((W, NS))
        tt = x[d] * pj
        xn[:, 1:] = x[d, :, :-1] * pj[:, 1:] + tt[:, 1:]
        xn[:, 0] = up * (pj[:, 0] * fac[d]) + tt[:, 0]
        return xn, sums

    # ---- first half
    for k in range(Ha):
        if k % RB == 0:
            block_entry(store_blk=k // RB)
        for d in range(2):
            valid = not (d == 1 and k == 0 and odd)
            if not valid:
                continue
            t = t_first(d, k)
            pj = load_p(d, t)
            xn, _ = step(d, pj, False)
            x[d] = xn
            firstp[d] = False
            AB[t, d] = x[d]
    # ---- boundary
    block_entry()
    # Z pre-step: beta lanes, t = Ha-1
    d = 1
    up = np.zeros(W)
    up[1:] = x[d, :-1, NS - 1]
    sums = np.empty((W, NS))
    sums[:, 1:] = x[d, :, 1:] + x[d, :, :-1]
    sums[:, 0] = up * fac[d] + x[d, :, 0]
    kp = Ha - 1
    part = np.zeros(W)
    ep = np.full(W, SENT, dtype=np.int64)
    for hl in range(W):
        o = AB[Ha - 1, 0, W - 1 - hl][::-1]
        eo = E[0, kp >> 3, W - 1 - hl]
        ms, mo = expo(sums[hl]), expo(o)
        if ms is None or mo is None:
            continue
        part[hl] = float(np.sum((sums[hl] * pow2z(-ms)) * (o * pow2z(-mo))))
        if part[hl] > 0:
            ep[hl] = e[d, hl] + eo + ms + mo
    emax = int(ep.max())
    zsum = sum(part[hl] * pow2z(int(ep[hl] - emax)) for hl in range(W) if ep[hl] != SENT)
    if not (zsum > 0):
        return np.inf, gam
    ezz = int(np.floor(np.log2(zsum)))
    zhat = zsum * pow2z(-ezz)
    Ez = emax + ezz
    loss = -(np.log(zhat) + Ez * np.log(2.0))
    zinv = -wgt / zhat
    # ---- second half
    for k in range(Ha):
        if k % RB == 0 and k > 0:
            block_entry()
        for d in range(2):
            valid = not (d == 0 and k == 0 and odd)
            if not valid:
                continue
            t = t_second(d, k)
            kp = (Tb - 1 + odd - t) if d == 0 else t
            od = 1 - d
            pj = load_p(d, t)
            xn, sums = step(d, pj, True)
            for hl in range(W):
                o = AB[t, od, W - 1 - hl][::-1]
                eo = E[od, kp >> 3, W - 1 - hl]
                dd = int(e[d, hl] + eo - Ez)
                sA = pow2z(dd >> 1)
                sB = pow2z(dd - (dd >> 1)) * zinv
                g = (sums[hl] * sA) * (o * sB)
                for j in range(NS):
                    s = sidx[d, hl, j]
                    if s < Lb:
                        gam[t, s] = np.float32(g[j])
            x[d] = xn
    return loss, gam


def run_case(seed, T, C, L, Tb, Lb, boost=0.0, NS=None):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((T, 1, C)).astype(np.float32)
    if boost:
        x[:, :, 0] += boost
    lab = rs.randint(0, C, size=(1, L)).astype(np.int32)
    lab[0, Lb:] = -1
    ref = R.nbctc_loss_grad(x, lab, np.array([Tb]), np.array([Lb]), reduction="sum")
    lp = R.log_softmax(x.astype(np.float64))[:, 0, :]
    NS = NS or max(2, (L + 15) // 16)
    Lpad = 16 * NS
    p = np.zeros((T, Lpad))
    for s in range(Lb):
        p[:, s] = np.maximum(np.exp(lp[:, lab[0, s]]).astype(np.float32), PMIN)
    loss, gam = chain(p, Tb, Lb, NS)
    # gradient from gamma'
    g = np.exp(lp)
    g[Tb:] = 0
    for s in range(Lb):
        g[:Tb, lab[0, s]] += gam[:Tb, s]
    gref = ref["grad"][:, 0, :]
    rl = abs(loss - ref["per_seq"][0]) / abs(ref["per_seq"][0])
    rg = np.linalg.norm(g - gref) / np.linalg.norm(gref)
    li = np.abs(g - gref).max() / np.abs(gref).max()
    return rl, rg, li, loss, ref["per_seq"][0]


if __name__ == "__main__":
    cases = [
        dict(seed=0, T=20, C=11, L=8, Tb=20, Lb=5),
        dict(seed=1, T=21, C=11, L=8, Tb=21, Lb=8),
        dict(seed=2, T=33, C=30, L=32, Tb=32, Lb=32),
        dict(seed=3, T=33, C=30, L=32, Tb=33, Lb=1),
        dict(seed=4, T=9, C=7, L=4, Tb=2, Lb=2),
        dict(seed=5, T=9, C=7, L=4, Tb=3, Lb=3),
        dict(seed=6, T=9, C=7, L=4, Tb=2, Lb=1),
        dict(seed=7, T=64, C=157, L=32, Tb=57, Lb=17),
        dict(seed=8, T=256, C=157, L=32, Tb=256, Lb=32, boost=14.0),
        dict(seed=9, T=512, C=157, L=64, Tb=512, Lb=64, boost=14.0),
        dict(seed=10, T=300, C=157, L=64, Tb=299, Lb=40, boost=25.0),
        dict(seed=11, T=1024, C=64, L=256, Tb=1000, Lb=256, boost=8.0),
        dict(seed=12, T=700, C=64, L=128, Tb=640, Lb=100, boost=20.0, NS=8),
    ]
    for c in cases:
        rl, rg, li, loss, ref = run_case(**c)
        print(f"{c}: loss {loss:.6f} ref {ref:.6f} rel {rl:.2e} grad L2 {rg:.2e} Linf {li:.2e}")
        assert rl < 1e-6 and rg < 1e-5 and li < 1e-5, c
    print("ok")


def run_dominant(seed, T, C, L, Tb, Lb, boost, s0):
    """boosted class 0 is the label of state s0 (the advisor's 'confident in one label' case)."""
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((T, 1, C)).astype(np.float32)
    x[:, :, 0] += boost
    lab = rs.randint(1, C, size=(1, L)).astype(np.int32)
    lab[0, s0] = 0
    lab[0, Lb:] = -1
    ref = R.nbctc_loss_grad(x, lab, np.array([Tb]), np.array([Lb]), reduction="sum")
    lp = R.log_softmax(x.astype(np.float64))[:, 0, :]
    NS = max(2, (L + 15) // 16)
    p = np.zeros((T, 16 * NS))
    for s in range(Lb):
        p[:, s] = np.maximum(np.exp(lp[:, lab[0, s]]).astype(np.float32), PMIN)
    loss, gam = chain(p, Tb, Lb, NS)
    g = np.exp(lp)
    g[Tb:] = 0
    for s in range(Lb):
        g[:Tb, lab[0, s]] += gam[:Tb, s]
    gref = ref["grad"][:, 0, :]
    return abs(loss - ref["per_seq"][0]) / abs(ref["per_seq"][0]), np.abs(g - gref).max() / np.abs(gref).max(), loss


if __name__ == "__main__":
    for c in [dict(seed=20, T=512, C=157, L=64, Tb=512, Lb=64, boost=14.0, s0=0),
              dict(seed=21, T=512, C=157, L=64, Tb=512, Lb=64, boost=14.0, s0=20),
              dict(seed=22, T=512, C=157, L=64, Tb=512, Lb=64, boost=30.0, s0=63),
              dict(seed=23, T=1024, C=64, L=256, Tb=1024, Lb=256, boost=8.0, s0=0),
              dict(seed=24, T=1024, C=64, L=256, Tb=1024, Lb=256, boost=20.0, s0=100)]:
        rl, li, loss = run_dominant(**c)
        print(f"dominant {c}: loss {loss:.4f} rel {rl:.2e} Linf {li:.2e}")
        assert rl < 1e-6 and li < 1e-5
    print("dominant ok")
