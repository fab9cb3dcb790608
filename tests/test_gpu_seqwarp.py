"""GPU tests of the sequence-per-warp kernels (seqwarp_kernel.cuh for C <= 256 / Lmax <= 64, seqwide_kernel.cuh for
C % 4 == 0, C <= 1024, Lmax <= 256) through the C ABI: shapes around their template boundaries, repeated labels, the
multi-wave work queue with ragged lengths, bit reproducibility, and the row log-partition hooks of SURVEY.md 8(f3)
(nbctc_loss_grad_lse_f32).  Oracle: oracle/c (float64), pinned to the reference by tests/test_oracle_cport.py."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from helpers import make_ctc_case
from oracle import cport

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5
SEQWARP = 32


def _call(x, lab, il, tl, flags=SEQWARP, lse_in=None, want_lse=False, w=None, seq_w=None, want_grad=True):
    from ctc_b200 import _ffi
    lib = _ffi.lib()
    T, B, C = x.shape
    Lmax = lab.shape[1]
    xt = torch.tensor(x, device=DEV)
    labt = torch.tensor(lab, device=DEV).int().contiguous()
    ilt, tlt = torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV)
    per = torch.empty(B, device=DEV)
    grad = torch.full_like(xt, float("nan")) if want_grad else None
    lse_out = torch.full((T, B), float("nan"), device=DEV) if want_lse else None
    lse_in_t = None if lse_in is None else torch.tensor(lse_in, device=DEV, dtype=torch.float32)
    swt = None if seq_w is None else torch.tensor(seq_w, device=DEV, dtype=torch.float32)
    flags |= _ffi.FLAG_ALIGNED16
    wsb = int(lib.nbctc_workspace_bytes(T, B, C, Lmax, 0, flags))
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=DEV)
    p = lambda t: None if t is None else t.data_ptr()
    rc = lib.nbctc_loss_grad_lse_f32(xt.data_ptr(), T, B, C, labt.data_ptr(), Lmax, ilt.data_ptr(), tlt.data_ptr(), p(lse_in_t),
                                     p(lse_out), per.data_ptr(), None, None, p(grad), p(swt), float(1.0 if w is None else w),
                                     ws.data_ptr(), wsb, flags, torch.cuda.current_stream().cuda_stream)
    _ffi.check(rc, "nbctc_loss_grad_lse_f32")
    torch.cuda.synchronize()
    return (per.cpu().double().numpy(), None if grad is None else grad.cpu().double().numpy(),
            None if lse_out is None else lse_out.cpu().double().numpy())


def _check(per, grad, x, lab, il, tl, scale=None):
    ref = cport.loss_grad("ctc", x, lab, il, tl, reduction="sum")
    assert np.max(np.abs(per - ref["per_seq"]) / np.maximum(np.abs(ref["per_seq"]), 1e-2)) < TOL
    want = ref["grad"] if scale is None else ref["grad"] * scale
    assert rel_l2(grad, want) < TOL
    assert np.max(np.abs(grad - want)) / max(np.max(np.abs(want)), 1e-30) < TOL
    T = x.shape[0]
    dead = np.arange(T)[:, None] >= il[None, :]
    assert np.all(grad[dead] == 0.0)


SHAPES = [
    # T, B, C, Lmax, dup          narrow kernel: rows of 1..8 elements per lane, 1 or 2 states per lane
    (9, 7, 1, 1, False), (30, 5, 31, 9, False), (30, 5, 32, 32, True), (41, 6, 33, 33, False), (64, 8, 157, 8, False),
    (77, 4, 96, 64, True), (50, 3, 129, 20, False), (50, 3, 160, 40, False), (35, 5, 200, 64, False), (44, 4, 224, 10, False),
    (44, 4, 255, 31, True), (23, 9, 256, 64, False), (1, 4, 157, 1, False), (2, 4, 157, 2, False), (3, 4, 157, 3, False),
    (5, 4, 157, 4, False), (8, 3, 157, 8, False), (13, 3, 157, 5, True),
    # wide kernel: C % 4 == 0 beyond the narrow range, 1 / 2 / 4 / 8 states per lane, partial and full rows
    (40, 3, 260, 20, False), (40, 3, 512, 60, True), (33, 4, 516, 100, False), (29, 3, 1024, 29, False), (70, 2, 1000, 256, False),
    (130, 3, 64, 128, True), (150, 2, 16, 130, False), (6, 3, 384, 5, False), (9, 2, 768, 3, False),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "T%d_B%d_C%d_L%d_%d" % s)
def test_seqwarp_shapes(shape):
    T, B, C, Lmax, dup = shape
    x, lab, il, tl = make_ctc_case(300 + T + C, T, B, C, Lmax, dup=dup)
    per, grad, _ = _call(x, lab, il, tl)
    _check(per, grad, x, lab, il, tl)


def test_seqwarp_all_labels_equal():
    """Every state carries the same class: the follower walk runs its longest chain (31 rounds), in state order."""
    T, B, C, Lmax = 80, 6, 50, 32
    x, lab, il, tl = make_ctc_case(5, T, B, C, Lmax, Lmin=20)
    lab[lab >= 0] = 7
    per, grad, _ = _call(x, lab, il, tl)
    _check(per, grad, x, lab, il, tl)


@pytest.mark.parametrize("shape", [(40, 9000, 40, 12), (24, 1500, 64, 70)], ids=["narrow", "wide"])
def test_seqwarp_multi_wave_work_queue_is_bit_reproducible(shape):
    """More sequences than resident warps with ragged lengths: the longest-first work queue hands them out in a
    run-dependent order; the results must not depend on it."""
    T, B, C, Lmax = shape
    x, lab, il, tl = make_ctc_case(6, T, B, C, Lmax)
    per1, g1, _ = _call(x, lab, il, tl)
    per2, g2, _ = _call(x, lab, il, tl)
    assert np.array_equal(per1, per2) and np.array_equal(g1, g2)
    _check(per1, g1, x, lab, il, tl)


def test_seqwarp_weights_and_infeasible():
    T, B, C, Lmax = 30, 8, 157, 10
    x, lab, il, tl = make_ctc_case(7, T, B, C, Lmax)
    il[2] = tl[2] - 1 if tl[2] > 1 else 0          # T_b < L_b
    lab[5, 0] = C + 3                              # label out of range
    sw = np.array([1.0, 0.5, 1.0, 0.0, 2.0, 1.0, 1.0, 0.25], dtype=np.float32)
    per, grad, _ = _call(x, lab, il, tl, w=0.5, seq_w=sw)
    ok = np.array([b not in (2, 5) for b in range(B)])
    assert np.all(np.isinf(per[~ok])) and np.all(grad[:, ~ok] == 0.0)
    ref = cport.loss_grad("ctc", x[:, ok], lab[ok], il[ok], tl[ok], reduction="sum")
    assert np.max(np.abs(per[ok] - ref["per_seq"]) / np.abs(ref["per_seq"])) < TOL
    want = ref["grad"] * (0.5 * sw[ok].astype(np.float64))[None, :, None]
    assert rel_l2(grad[:, ok], want) < TOL
    assert np.all(grad[:, 3] == 0.0)


def test_seqwarp_no_grad():
    x, lab, il, tl = make_ctc_case(8, 50, 6, 157, 20)
    per, grad, _ = _call(x, lab, il, tl, want_grad=False)
    ref = cport.loss_grad("ctc", x, lab, il, tl, reduction="sum")
    assert grad is None and np.max(np.abs(per - ref["per_seq"]) / np.abs(ref["per_seq"])) < TOL


@pytest.mark.parametrize("shape", [(64, 8, 157, 8), (37, 5, 33, 20), (130, 4, 256, 64)], ids=lambda s: "T%d_B%d_C%d_L%d" % s)
def test_row_lse_out_then_in(shape):
    """the call hands out log sum_c exp(x[t,b,c]) for t < T_b; fed back in (the producer-fusion path, where
    phase 1 reads only the label entries of every row) it reproduces loss and gradient."""
    T, B, C, Lmax = shape
    x, lab, il, tl = make_ctc_case(11, T, B, C, Lmax)
    per, grad, lse = _call(x, lab, il, tl, want_lse=True)
    xd = x.astype(np.float64)
    m = xd.max(axis=2)
    want = m + np.log(np.exp(xd - m[:, :, None]).sum(axis=2))
    live = np.arange(T)[:, None] < il[None, :]
    assert np.max(np.abs(lse[live] - want[live])) < 2e-6
    assert np.all(np.isnan(lse[~live]))            # rows beyond input_length are not written
    per2, grad2, _ = _call(x, lab, il, tl, lse_in=np.where(live, want, 0.0))
    _check(per2, grad2, x, lab, il, tl)
    np.testing.assert_allclose(per2, per, rtol=2e-6)


def test_row_lse_rejected_on_wide_rows():
    from ctc_b200 import _ffi
    x, lab, il, tl = make_ctc_case(12, 8, 2, 512, 4)
    with pytest.raises(_ffi.NbctcError):
        _call(x, lab, il, tl, want_lse=True)


def test_default_dispatch_takes_the_seqwarp_kernel_for_large_batches():
    """B >= 2304 sequences of the narrow shape run the sequence-per-warp kernel by default: same bits as the forced path."""
    import ctc_b200
    T, B, C, Lmax = 12, 3200, 20, 6
    x, lab, il, tl = make_ctc_case(13, T, B, C, Lmax)
    per_f, g_f, _ = _call(x, lab, il, tl)
    xt = torch.tensor(x, device=DEV, requires_grad=True)
    per = ctc_b200.no_blank_ctc_loss(xt, torch.tensor(lab, device=DEV), torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV), "none")
    per.sum().backward()
    torch.cuda.synchronize()
    assert np.array_equal(per.detach().cpu().double().numpy(), per_f)
    assert np.array_equal(xt.grad.cpu().double().numpy(), g_f)


@pytest.mark.parametrize("shape", [(37, 9, 157, 20), (23, 5, 33, 33), (19, 4, 260, 9), (30, 3, 1000, 200), (40, 1500, 12, 5)],
                         ids=lambda s: "T%d_B%d_C%d_L%d" % s)
def test_no_write_outside_the_caller_buffers(shape):
    """Gradient, per-sequence losses and workspace sit between guard regions of a known byte pattern: the kernels write
    every byte of the gradient and nothing outside what nbctc_workspace_bytes() asked for."""
    from ctc_b200 import _ffi
    lib = _ffi.lib()
    T, B, C, Lmax = shape
    x, lab, il, tl = make_ctc_case(500 + T + C, T, B, C, Lmax)
    xt = torch.tensor(x, device=DEV)
    labt = torch.tensor(lab, device=DEV).int().contiguous()
    ilt, tlt = torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV)
    flags = SEQWARP | _ffi.FLAG_ALIGNED16
    wsb = int(lib.nbctc_workspace_bytes(T, B, C, Lmax, 0, flags))
    G = 4096                                              # guard bytes (keeps every buffer 256-byte aligned)
    up = lambda n: (n + 255) // 256 * 256
    n_grad, n_per, n_ws = up(T * B * C * 4), up(B * 4), up(wsb)
    arena = torch.full((G + n_grad + G + n_per + G + n_ws + G,), 0x5A, dtype=torch.uint8, device=DEV)
    o_grad, o_per, o_ws = G, G + n_grad + G, G + n_grad + G + n_per + G
    base = arena.data_ptr()
    assert base % 256 == 0
    rc = lib.nbctc_loss_grad_f32(xt.data_ptr(), T, B, C, labt.data_ptr(), Lmax, ilt.data_ptr(), tlt.data_ptr(), base + o_per, None, None,
                                 base + o_grad, None, 1.0, base + o_ws, wsb, flags, torch.cuda.current_stream().cuda_stream)
    _ffi.check(rc, "nbctc_loss_grad_f32")
    torch.cuda.synchronize()
    a = arena.cpu().numpy()
    for lo, hi in ((0, G), (o_grad + T * B * C * 4, o_per), (o_per + B * 4, o_ws), (o_ws + wsb, len(a))):
        assert np.all(a[lo:hi] == 0x5A), f"guard bytes [{lo},{hi}) were written"
    grad = a[o_grad:o_grad + T * B * C * 4].view(np.float32).reshape(T, B, C).astype(np.float64)
    per = a[o_per:o_per + B * 4].view(np.float32).astype(np.float64)
    _check(per, grad, x, lab, il, tl)


@pytest.mark.parametrize("path", [8, SEQWARP], ids=["lockstep", "seqwarp"])
@pytest.mark.parametrize("shape", [(37, 9, 157, 20), (23, 6, 33, 33), (30, 3, 1000, 200), (21, 70, 16, 150)],
                         ids=lambda s: "T%d_B%d_C%d_L%d" % s)
def test_repair_kernel_stays_inside_the_workspace(shape, path):
    """The guard-region check with every second sequence beyond the float32 emission floor (logits times 40): the
    log-domain repair kernel keeps its row log-partitions and checkpoints inside the workspace the query asked for."""
    from ctc_b200 import _ffi
    lib = _ffi.lib()
    T, B, C, Lmax = shape
    x, lab, il, tl = make_ctc_case(700 + T + C, T, B, C, Lmax)
    x[:, 0::2] *= 40.0
    xt = torch.tensor(x, device=DEV)
    labt = torch.tensor(lab, device=DEV).int().contiguous()
    ilt, tlt = torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV)
    flags = path | _ffi.FLAG_ALIGNED16
    wsb = int(lib.nbctc_workspace_bytes(T, B, C, Lmax, 0, flags))
    G = 4096
    up = lambda n: (n + 255) // 256 * 256
    n_grad, n_per, n_ws = up(T * B * C * 4), up(B * 4), up(wsb)
    arena = torch.full((G + n_grad + G + n_per + G + n_ws + G,), 0x5A, dtype=torch.uint8, device=DEV)
    o_grad, o_per, o_ws = G, G + n_grad + G, G + n_grad + G + n_per + G
    base = arena.data_ptr()
    rc = lib.nbctc_loss_grad_f32(xt.data_ptr(), T, B, C, labt.data_ptr(), Lmax, ilt.data_ptr(), tlt.data_ptr(), base + o_per, None, None,
                                 base + o_grad, None, 1.0, base + o_ws, wsb, flags, torch.cuda.current_stream().cuda_stream)
    _ffi.check(rc, "nbctc_loss_grad_f32")
    torch.cuda.synchronize()
    a = arena.cpu().numpy()
    for lo, hi in ((0, G), (o_grad + T * B * C * 4, o_per), (o_per + B * 4, o_ws), (o_ws + wsb, len(a))):
        assert np.all(a[lo:hi] == 0x5A), f"guard bytes [{lo},{hi}) were written"
    grad = a[o_grad:o_grad + T * B * C * 4].view(np.float32).reshape(T, B, C).astype(np.float64)
    per = a[o_per:o_per + B * 4].view(np.float32).astype(np.float64)
    _check(per, grad, x, lab, il, tl)


def test_misaligned_workspace_is_rejected():
    from ctc_b200 import _ffi
    lib = _ffi.lib()
    x, lab, il, tl = make_ctc_case(1, 8, 2, 16, 3)
    xt, labt = torch.tensor(x, device=DEV), torch.tensor(lab, device=DEV).int()
    ilt, tlt = torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV)
    per, grad = torch.empty(2, device=DEV), torch.empty_like(xt)
    wsb = int(lib.nbctc_workspace_bytes(8, 2, 16, 3, 0, 0))
    ws = torch.empty(wsb + 512, dtype=torch.uint8, device=DEV)
    rc = lib.nbctc_loss_grad_f32(xt.data_ptr(), 8, 2, 16, labt.data_ptr(), 3, ilt.data_ptr(), tlt.data_ptr(), per.data_ptr(), None, None,
                                 grad.data_ptr(), None, 1.0, ws.data_ptr() + 16, wsb, 0, torch.cuda.current_stream().cuda_stream)
    assert rc != 0 and b"aligned" in lib.nbctc_last_error()


@pytest.mark.parametrize("shape", [(37, 9, 157, 20), (41, 6, 40, 32), (26, 5, 100, 50)], ids=lambda s: "T%d_B%d_C%d_L%d" % s)
def test_multilabel_no_write_outside_the_caller_buffers(shape):
    """The same guard-region check for the multi-label variant (Lmax <= 32: the fused emission + lattice kernel)."""
    from ctc_b200 import _ffi
    from helpers import make_bctc_case
    lib = _ffi.lib()
    T, B, C, Lmax = shape
    x, y, il, tl = make_bctc_case(600 + T, T, B, C, Lmax)
    xt, yt = torch.tensor(x, device=DEV), torch.tensor(y, device=DEV).float().contiguous()
    ilt, tlt = torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV)
    flags = _ffi.FLAG_ALIGNED16
    wsb = int(lib.nbctc_workspace_bytes(T, B, C, Lmax, 1, flags))
    G = 4096
    up = lambda n: (n + 255) // 256 * 256
    n_grad, n_per, n_ws = up(T * B * C * 4), up(B * 4), up(wsb)
    arena = torch.full((G + n_grad + G + n_per + G + n_ws + G,), 0x5A, dtype=torch.uint8, device=DEV)
    o_grad, o_per, o_ws = G, G + n_grad + G, G + n_grad + G + n_per + G
    base = arena.data_ptr()
    rc = lib.nbbctc_loss_grad_f32(xt.data_ptr(), T, B, C, yt.data_ptr(), Lmax, ilt.data_ptr(), tlt.data_ptr(), base + o_per, None, None,
                                  base + o_grad, None, 1.0, base + o_ws, wsb, flags, torch.cuda.current_stream().cuda_stream)
    _ffi.check(rc, "nbbctc_loss_grad_f32")
    torch.cuda.synchronize()
    a = arena.cpu().numpy()
    for lo, hi in ((0, G), (o_grad + T * B * C * 4, o_per), (o_per + B * 4, o_ws), (o_ws + wsb, len(a))):
        assert np.all(a[lo:hi] == 0x5A), f"guard bytes [{lo},{hi}) were written"
    grad = a[o_grad:o_grad + T * B * C * 4].view(np.float32).reshape(T, B, C).astype(np.float64)
    per = a[o_per:o_per + B * 4].view(np.float32).astype(np.float64)
    ref = cport.loss_grad("bctc", x, y, il, tl, reduction="sum")
    assert np.max(np.abs(per - ref["per_seq"]) / np.abs(ref["per_seq"])) < TOL
    assert rel_l2(grad, ref["grad"]) < TOL
