"""Float64 CPU restatement of the reference's no-blank CTC losses (TEST INFRASTRUCTURE).

This file is the *oracle*: a vectorised (over batch and states, sequential over time)
numpy/float64 restatement of what ``/root/reference/NoBlankCTC.py`` and
``/root/reference/NoBlankBinaryCTC.py`` compute, plus the closed-form gradient that the
reference obtains through autograd.  It is pinned against the reference itself by
``tests/test_oracle_golden.py`` (fixtures in ``tests/golden/*.npz`` were produced by
importing the unmodified reference modules in float64 -- see
``tests/golden/make_golden.py``).

Never import this from ``ctc_b200``: the product path is CUDA only.

Reference map (file:line in /root/reference):
  * log-softmax of the logits ............... NoBlankCTC.py:136
  * emission gather lp[t,b,label[b,s]] ...... NoBlankCTC.py:96-102
  * alpha initial state [0,-inf,...] ........ NoBlankCTC.py:92-93, shift guard t>0 :75-76
  * alpha step logaddexp(stay, advance) ..... NoBlankCTC.py:73-85 (+ _logsumexp :16-19)
  * state mask s >= target_length ........... NoBlankCTC.py:79-80
  * read-out alpha[T_b-1, L_b-1] ............ NoBlankCTC.py:58-68 (flip) and :139
  * mean over the batch ..................... NoBlankCTC.py:140
  * sigmoid + (-BCELoss) emissions .......... NoBlankBinaryCTC.py:146, :109-112, :85-88
The reference uses -1e13 where this file uses -inf; inside the parity domain
(1 <= L_b <= T_b <= T) every reachable cell is identical (logaddexp(a, -1e13) == a in
float64 for |a| < 1e12) and unreachable cells never reach the read-out.
"""
from __future__ import annotations

import numpy as np

NEG_INF = -np.inf


def _as_f64(x):
    return np.ascontiguousarray(np.asarray(x), dtype=np.float64)


def _as_i64(x):
    return np.ascontiguousarray(np.asarray(x), dtype=np.int64)


def log_softmax(x: np.ndarray) -> np.ndarray:
    """Row log-softmax over the last axis (NoBlankCTC.py:136)."""
    m = np.max(x, axis=-1, keepdims=True)
    z = x - m
    return z - np.log(np.sum(np.exp(z), axis=-1, keepdims=True))


def _logaddexp(a, b):
    with np.errstate(invalid="ignore"):
        return np.logaddexp(a, b)


def nbctc_alpha_beta(E: np.ndarray, input_length, target_length):
    """Forward/backward DP on log emissions ``E`` (T,B,Lmax) float64.

    Returns ``(ll, alpha, beta)``: per-sequence log-likelihood (B,), and the log-domain
    alpha/beta lattices (T,B,Lmax) with -inf outside the reachable region.
    alpha includes the emission at t; beta excludes it.
    Infeasible sequences (L_b < 1, T_b < L_b, T_b > T) get ll = -inf.
    """
    E = _as_f64(E)
    T, B, L = E.shape
    Tb = _as_i64(input_length).copy()
    Lb = _as_i64(target_length).copy()
    # reference quirk: input_length == 0 behaves like T (NoBlankCTC.py:62, modulo)
    s_idx = np.arange(L)[None, :]
    outside = s_idx >= Lb[:, None]                       # NoBlankCTC.py:79
    feasible = (Lb >= 1) & (Lb <= L) & (Tb >= Lb) & (Tb <= T) & (Tb >= 1)

    alpha = np.full((T, B, L), NEG_INF)
    a = np.full((B, L), NEG_INF)
    a[:, 0] = 0.0                                        # NoBlankCTC.py:92-93
    for t in range(T):
        if t > 0:                                        # NoBlankCTC.py:75
            shifted = np.concatenate([np.full((B, 1), NEG_INF), a[:, :-1]], axis=1)
            a = _logaddexp(a, shifted)
        a = a + E[t]
        a[outside] = NEG_INF
        alpha[t] = a

    ll = np.full(B, NEG_INF)
    ok = np.nonzero(feasible)[0]
    ll[ok] = alpha[Tb[ok] - 1, ok, Lb[ok] - 1]

    beta = np.full((T, B, L), NEG_INF)
    b = np.full((B, L), NEG_INF)
    for t in range(T - 1, -1, -1):
        if t < T - 1:
            be = b + E[t + 1]
            be[outside] = NEG_INF
            shifted = np.concatenate([be[:, 1:], np.full((B, 1), NEG_INF)], axis=1)
            b = _logaddexp(be, shifted)
        last = feasible & (Tb - 1 == t)
        if last.any():
            b = b.copy()
            b[last] = NEG_INF
            b[last, Lb[last] - 1] = 0.0
        b[Tb - 1 < t] = NEG_INF
        b[outside] = NEG_INF
        beta[t] = b
    return ll, alpha, beta


def _gamma(ll, alpha, beta):
    with np.errstate(invalid="ignore", over="ignore"):
        g = np.exp(alpha + beta - ll[None, :, None])
    g[~np.isfinite(g)] = 0.0
    return g


def _reduce(per_seq, B, reduction):
    if reduction == "mean":
        return per_seq.mean(), np.full(B, 1.0 / B)
    if reduction == "sum":
        return per_seq.sum(), np.ones(B)
    if reduction == "none":
        return per_seq, np.ones(B)
    raise ValueError(reduction)


def nbctc_loss_grad(logits, labels, input_length, target_length, reduction="mean"):
    """NoBlankCTC (NoBlankCTC.py:129-141) and its autograd gradient in closed form.

    logits (T,B,C) raw scores; labels (B,Lmax) int, slots >= L_b ignored (-1 padding ok).
    Returns dict(loss, per_seq, grad, gamma, ll) with grad = d(loss)/d(logits) for
    'mean'/'sum', and d(sum_b loss_b)/d(logits) for 'none'.
    grad[t,b,c] = w_b * (softmax(x[t,b])[c] - sum_{s<L_b, label_s=c} gamma_t(s)), t < T_b; 0 else.
    """
    x = _as_f64(logits)
    T, B, C = x.shape
    lab = _as_i64(labels)
    Lmax = lab.shape[1]
    Lb = _as_i64(target_length)
    valid = np.arange(Lmax)[None, :] < Lb[:, None]
    lab_safe = np.where(valid, lab, 0)
    if ((lab_safe < 0) | (lab_safe >= C)).any():
        raise ValueError("label out of range inside target_length")
    lp = log_softmax(x)
    E = np.take_along_axis(lp, np.broadcast_to(lab_safe[None], (T, B, Lmax)), axis=2)
    ll, alpha, beta = nbctc_alpha_beta(E, input_length, target_length)
    gamma = _gamma(ll, alpha, beta) * valid[None]
    per_seq = -ll
    loss, w = _reduce(per_seq, B, reduction)
    Tb = _as_i64(input_length)
    live = (np.arange(T)[:, None] < Tb[None, :]) & np.isfinite(ll)[None, :]
    grad = np.exp(lp)
    sub = np.zeros_like(grad)
    tt, bb, ss = np.nonzero(gamma != 0.0)
    np.add.at(sub, (tt, bb, lab_safe[bb, ss]), gamma[tt, bb, ss])   # duplicates accumulate
    grad = (grad - sub) * live[:, :, None] * w[None, :, None]
    return dict(loss=loss, per_seq=per_seq, grad=grad, gamma=gamma, ll=ll)


def nbbctc_emissions(logits, targets, clamp=True):
    """Binary emissions e[t,b,s] = -BCELoss(sigmoid(x[t,b,:]), y[b,s,:]) (mean over C).

    NoBlankBinaryCTC.py:146 (sigmoid), :109-112 (bce * -1).  ``clamp`` reproduces
    nn.BCELoss's clamp of each log term at -100.
    """
    x = _as_f64(logits)
    y = _as_f64(targets)
    C = x.shape[2]
    # log(sigmoid(x)) = -softplus(-x); log(1-sigmoid(x)) = -softplus(x)
    ls = -np.logaddexp(0.0, -x)
    l1s = -np.logaddexp(0.0, x)
    if clamp:
        ls = np.maximum(ls, -100.0)
        l1s = np.maximum(l1s, -100.0)
    # e[t,b,s] = (1/C) sum_c y[b,s,c]*ls[t,b,c] + (1-y[b,s,c])*l1s[t,b,c]
    E = (np.einsum("bsc,tbc->tbs", y, ls) + np.einsum("bsc,tbc->tbs", 1.0 - y, l1s)) / C
    return E


def nbbctc_loss_grad(logits, targets, input_length, target_length, reduction="mean"):
    """NoBlankBinaryCTC (NoBlankBinaryCTC.py:139-151) with closed-form gradient.

    targets (B,Lmax,C) float multi-hot (rows >= L_b ignored).
    grad[t,b,c] = w_b * (sigmoid(x[t,b,c]) - sum_s gamma_t(s) y[b,s,c]) / C for t < T_b.
    """
    x = _as_f64(logits)
    y = _as_f64(targets)
    T, B, C = x.shape
    Lmax = y.shape[1]
    Lb = _as_i64(target_length)
    valid = np.arange(Lmax)[None, :] < Lb[:, None]
    y = y * valid[:, :, None]
    E = nbbctc_emissions(x, y)
    ll, alpha, beta = nbctc_alpha_beta(E, input_length, target_length)
    gamma = _gamma(ll, alpha, beta) * valid[None]
    per_seq = -ll
    loss, w = _reduce(per_seq, B, reduction)
    Tb = _as_i64(input_length)
    live = (np.arange(T)[:, None] < Tb[None, :]) & np.isfinite(ll)[None, :]
    sig = 1.0 / (1.0 + np.exp(-x))
    sub = np.einsum("tbs,bsc->tbc", gamma, y)
    grad = (sig - sub) / C * live[:, :, None] * w[None, :, None]
    return dict(loss=loss, per_seq=per_seq, grad=grad, gamma=gamma, ll=ll)


def best_path(E_or_logits, labels, input_length, target_length, gathered=False):
    """Max-plus (Viterbi) twin of the alpha recursion on the same lattice.

    Not in the reference; defined here as the oracle for the
    integer outputs.  Scores are accumulated in float64 over the RAW logits
    x[t,b,label_s] (the per-row log-partition is common to all paths so it does not
    change the argmax), recurrence ``v_t(s) = max(v_{t-1}(s), v_{t-1}(s-1)) + x``;
    ties prefer *stay* (state s).  float64 add/max are exactly reproducible on the GPU,
    which is what makes the bit-exact comparison meaningful.

    Returns (states (B,T) int32 padded with -1 for t >= T_b, score (B,) float64).
    """
    if gathered:
        E = _as_f64(E_or_logits)
    else:
        x = np.asarray(E_or_logits)
        T, B, C = x.shape
        lab = _as_i64(labels)
        Lmax = lab.shape[1]
        valid = np.arange(Lmax)[None, :] < _as_i64(target_length)[:, None]
        lab_safe = np.where(valid, lab, 0)
        E = np.take_along_axis(x, np.broadcast_to(lab_safe[None], (T, B, Lmax)), axis=2).astype(np.float64)
    T, B, L = E.shape
    Tb = _as_i64(input_length)
    Lb = _as_i64(target_length)
    outside = np.arange(L)[None, :] >= Lb[:, None]
    feasible = (Lb >= 1) & (Lb <= L) & (Tb >= Lb) & (Tb <= T)
    v = np.full((B, L), NEG_INF)
    v[:, 0] = E[0, :, 0]
    v[outside] = NEG_INF
    back = np.zeros((T, B, L), dtype=np.int8)            # 1 = came from s-1
    finals = np.full(B, NEG_INF)
    for t in range(T):
        if t > 0:
            shifted = np.concatenate([np.full((B, 1), NEG_INF), v[:, :-1]], axis=1)
            adv = shifted > v                            # tie -> stay
            v = np.where(adv, shifted, v) + E[t]
            v[outside] = NEG_INF
            back[t] = adv
        done = np.nonzero(feasible & (Tb - 1 == t))[0]
        finals[done] = v[done, Lb[done] - 1]
    states = np.full((B, T), -1, dtype=np.int32)
    for b in range(B):
        if not feasible[b]:
            continue
        s = int(Lb[b] - 1)
        for t in range(int(Tb[b]) - 1, -1, -1):
            states[b, t] = s
            if t > 0 and back[t, b, s]:
                s -= 1
    return states, finals


def frame_argmax(logits):
    """Per-frame argmax class, ties -> lowest index (train.py:41-56 uses topk on logits)."""
    x = np.asarray(logits)
    return np.argmax(x, axis=2).astype(np.int32)


# ---------------------------------------------------------------------------------------------
# Auxiliary cross-entropy on one frame per sequence: the reference mixes the CTC loss with a
# cross-entropy on the last frame's scores, weighted by --alpha (opts.py:74, main.py:42, train.py:353).
#   mode "index":    nn.CrossEntropyLoss on class indices (models/__init__.py:85):  ce_b = lse(x_b) - x_b[y_b]
#   mode "multihot": the reference's own CrossEntropy module (CrossEntropy.py:17-32) on multi-hot targets:
#                    q = softmax(x_b) (:22); ce_b = log sum_c exp(q_c) (:25) - sum_{n: target[b][n] == 1} q_n (:26-29)
# Both are averaged over the batch (CrossEntropy.py:30; nn.CrossEntropyLoss default reduction).
def aux_ce(logits, targets, frame_index, mode="index", reduction="mean"):
    """logits (T,B,C); frame_index (B) = the frame whose scores are classified (train.py:434 uses temporal-1);
    targets (B,) class indices or (B,C) multi-hot.  Returns dict(loss, per_seq, grad) with grad = d loss / d logits
    (zero outside the B classified rows)."""
    x = _as_f64(logits)
    T, B, C = x.shape
    fi = _as_i64(frame_index)
    rows = x[fi, np.arange(B)]                                  # (B,C)
    q = np.exp(log_softmax(rows))
    if mode == "index":
        y = _as_i64(targets)
        per = -log_softmax(rows)[np.arange(B), y]
        g_rows = q.copy()
        g_rows[np.arange(B), y] -= 1.0
    elif mode == "multihot":
        pos = (_as_f64(targets) == 1.0)                         # CrossEntropy.py:28 tests equality with 1
        eq = np.exp(q)
        S = eq.sum(axis=1)
        per = np.log(S) - (q * pos).sum(axis=1)
        g = eq / S[:, None] - pos                               # d per / d q
        g_rows = q * (g - (g * q).sum(axis=1, keepdims=True))   # through the softmax
    else:
        raise ValueError(mode)
    loss, w = _reduce(per, B, reduction)
    grad = np.zeros_like(x)
    grad[fi, np.arange(B)] = g_rows * w[:, None]
    return dict(loss=loss, per_seq=per, grad=grad)
