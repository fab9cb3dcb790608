"""Shared input generators for the parity tests (seeded, numpy)."""
import numpy as np


def make_ctc_case(seed, T, B, C, Lmax, ragged_T=True, Lmin=1, dup=False):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((T, B, C)).astype(np.float32)
    hi = min(Lmax, T)
    tl = rs.randint(min(Lmin, hi), hi + 1, size=B).astype(np.int64)
    if ragged_T:
        il = np.array([rs.randint(max(l, (T + 1) // 2), T + 1) for l in tl], dtype=np.int64)
    else:
        il = np.full(B, T, dtype=np.int64)
    lab = rs.randint(0, C, size=(B, Lmax)).astype(np.int32)
    if dup:
        lab[:, 1::2] = lab[:, 0::2][:, : lab[:, 1::2].shape[1]]
    for b in range(B):
        lab[b, tl[b]:] = -1
    return x, lab, il, tl


def make_bctc_case(seed, T, B, C, Lmax, density=0.03, ragged_T=True, pad=0.0):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((T, B, C)).astype(np.float32)
    hi = min(Lmax, T)
    tl = rs.randint(1, hi + 1, size=B).astype(np.int64)
    if ragged_T:
        il = np.array([rs.randint(max(l, (T + 1) // 2), T + 1) for l in tl], dtype=np.int64)
    else:
        il = np.full(B, T, dtype=np.int64)
    y = (rs.uniform(size=(B, Lmax, C)) < density).astype(np.float32)
    pick = rs.randint(0, C, size=(B, Lmax))
    bi, si = np.meshgrid(np.arange(B), np.arange(Lmax), indexing="ij")
    y[bi, si, pick] = 1.0                       # at least one active class per row
    for b in range(B):
        y[b, tl[b]:] = pad                      # padded rows (the dataset pads with -1)
    return x, y, il, tl
