"""GPU parity: the CUDA path (through the nn.Module -> ctypes -> C ABI) against the float64 oracle
and the reference-generated golden fixtures.  Tolerance (BASELINE.json north_star): loss and
gradient within 1e-5 relative of the float64 reference; integer outputs bit-exact."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from helpers import make_bctc_case, make_ctc_case
from oracle import restatement as R

pytestmark = pytest.mark.gpu

TOL = 1e-5
DEV = "cuda:0"


@pytest.fixture(scope="module")
def nb():
    import ctc_b200
    assert torch.cuda.is_available()
    return ctc_b200


def run_cuda(nb, kind, x, tg, il, tl, reduction="mean", flags=0, lab_dtype=torch.int32):
    xt = torch.tensor(x, device=DEV, requires_grad=True)
    if kind == "ctc":
        m = nb.NoBlankCTC(reduction=reduction, flags=flags)
        tgt = torch.tensor(np.asarray(tg), device=DEV).to(lab_dtype)
    else:
        m = nb.NoBlankBinaryCTC(reduction=reduction, flags=flags)
        tgt = torch.tensor(np.asarray(tg), device=DEV).float()
    loss = m(xt, tgt, torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV))
    (loss.sum() if reduction == "none" else loss).backward()
    torch.cuda.synchronize()
    return loss.detach().cpu().double().numpy(), xt.grad.detach().cpu().double().numpy()


def oracle(kind, x, tg, il, tl, reduction="mean"):
    f = R.nbctc_loss_grad if kind == "ctc" else R.nbbctc_loss_grad
    return f(x, tg, il, tl, reduction)


def assert_parity(loss, grad, ref, tol=TOL, scale=None):
    rl = np.max(np.abs(loss - ref["loss"]) / np.maximum(np.abs(ref["loss"]), 1e-2))
    rg = rel_l2(grad, ref["grad"])
    assert rl < tol, f"loss rel err {rl}"
    assert rg < tol, f"grad L2 rel err {rg}"
    # Linf relative to the largest gradient entry: the north-star 1e-5 as well
    # (floor: a tenth of the per-sequence weight, the natural size of a gradient entry -- the C=1 case has grad == 0)
    floor = 0.1 / grad.shape[1] if scale is None else scale
    linf = np.max(np.abs(grad - ref["grad"])) / max(np.max(np.abs(ref["grad"])), floor)
    assert linf < tol, f"grad Linf rel err {linf}"


@pytest.mark.parametrize("flags", [0, 1], ids=["default", "generic"])
def test_golden_fixtures(nb, golden, flags):
    kind = str(golden["kind"])
    loss, grad = run_cuda(nb, kind, golden["logits"], golden["targets"], golden["input_length"],
                          golden["target_length"], flags=flags)
    assert abs(loss - float(golden["loss"])) / abs(float(golden["loss"])) < TOL
    assert rel_l2(grad, golden["grad"]) < TOL


CTC_SHAPES = [
    # T, B, C, Lmax, ragged_T, dup
    (1, 3, 5, 1, False, False),
    (7, 1, 3, 7, False, False),
    (64, 8, 157, 8, False, False),        # cfg1
    (64, 8, 157, 8, True, True),
    (33, 5, 33, 31, True, False),         # --v-class 33
    (50, 6, 10, 32, True, True),
    (40, 4, 157, 40, True, False),        # Lmax > 32
    (96, 3, 64, 64, True, False),
    (130, 2, 1024, 100, True, False),
    (300, 2, 40, 256, True, False),       # cfg4-style state count
    (257, 9, 157, 17, True, False),
    (512, 4, 157, 64, True, False),       # cfg5-style
    (20, 7, 1, 1, False, False),          # single class
    (24, 5, 4, 300, True, False),         # Lmax > 256 (clamped to T)
]


@pytest.mark.parametrize("flags", [0, 1, 8, 32], ids=["default", "generic", "lockstep", "seqwarp"])
@pytest.mark.parametrize("shape", CTC_SHAPES, ids=lambda s: "T%d_B%d_C%d_L%d_%d%d" % s)
def test_ctc_random_vs_oracle(nb, shape, flags):
    T, B, C, L, ragged, dup = shape
    x, lab, il, tl = make_ctc_case(100 + T + L, T, B, C, L, ragged_T=ragged, dup=dup)
    loss, grad = run_cuda(nb, "ctc", x, lab, il, tl, flags=flags)
    assert_parity(loss, grad, oracle("ctc", x, lab, il, tl))


BCTC_SHAPES = [
    (1, 2, 5, 1, 0.3),
    (64, 8, 157, 8, 0.03),
    (33, 5, 33, 31, 0.1),
    (40, 4, 157, 40, 0.03),
    (60, 3, 157, 32, 0.5),
    (100, 2, 300, 70, 0.02),
    (257, 5, 157, 17, 0.03),
    (20, 2, 400, 150, 0.02),              # multi-hot rows too large for shared memory: global-memory kernels
    (300, 3, 64, 40, 0.05),               # several 128-step chunks per sequence
    (50, 3, 100, 100, 0.03),              # tiled path, 8 chain states per lane
    (37, 2, 64, 200, 0.05),               # tiled path, 16 chain states per lane
    (130, 6, 256, 20, 0.02),              # tiled path, widest class dimension (byte class indices)
    (24, 3, 7, 5, 0.3),                   # tiny class dimension
]


@pytest.mark.parametrize("flags", [0, 1], ids=["default", "generic"])
@pytest.mark.parametrize("shape", BCTC_SHAPES, ids=lambda s: "T%d_B%d_C%d_L%d_p%g" % s)
def test_bctc_random_vs_oracle(nb, shape, flags):
    T, B, C, L, dens = shape
    x, y, il, tl = make_bctc_case(200 + T + L, T, B, C, L, density=dens)
    loss, grad = run_cuda(nb, "bctc", x, y, il, tl, flags=flags)
    assert_parity(loss, grad, oracle("bctc", x, y, il, tl))


def test_bctc_soft_targets_and_minus_one_padding(nb):
    """Targets need not be {0,1}; padded rows hold -1 in the reference's dataset
    (charades_ctc_pred.py:557-559) and must be ignored, not dereferenced into the maths."""
    T, B, C, L = 30, 4, 21, 6
    x, y, il, tl = make_bctc_case(7, T, B, C, L, density=0.2, pad=-1.0)
    rs = np.random.RandomState(1)
    soft = rs.uniform(size=y.shape).astype(np.float32)
    y_soft = np.where(y < 0, y, soft)
    for tg in (y, y_soft):
        loss, grad = run_cuda(nb, "bctc", x, tg, il, tl)
        clean = np.where(tg < 0, 0.0, tg)
        assert_parity(loss, grad, oracle("bctc", x, clean, il, tl))


def test_bctc_tiled_path_is_taken_and_bit_reproducible(nb):
    """Conforming multi-hot targets run the tiled kernels (Lmax <= 32: emissions + lattice in one sequence-per-warp kernel,
    gradient + the three gated fallback launches that return at once), twice with bit-identical results;
    NBCTC_FLAG_GENERIC runs three launches."""
    import ctc_b200
    x, y, il, tl = make_bctc_case(91, 70, 9, 157, 20, density=0.03)
    n0 = ctc_b200.launch_count()
    per1, g1 = _device_call_bin(nb, x, y, il, tl)
    n1 = ctc_b200.launch_count()
    per2, g2 = _device_call_bin(nb, x, y, il, tl)
    assert n1 - n0 == 5
    assert np.array_equal(per1, per2) and np.array_equal(g1, g2)
    n2 = ctc_b200.launch_count()
    per3, g3 = _device_call_bin(nb, x, y, il, tl, flags=1)
    assert ctc_b200.launch_count() - n2 == 3
    np.testing.assert_allclose(per1, per3, rtol=1e-6)
    assert rel_l2(g1, g3) < 1e-6


def test_bctc_soft_targets_large_batch_gated_fallback(nb):
    """More sequences than the gated fallback's grid (the generic kernels then loop over their virtual blocks)."""
    T, B, C, L = 12, 700, 6, 3
    x, y, il, tl = make_bctc_case(9, T, B, C, L, density=0.3)
    rs = np.random.RandomState(2)
    y_soft = np.where(np.arange(L)[None, :, None] < tl[:, None, None], rs.uniform(size=y.shape), 0.0).astype(np.float32)
    loss, grad = run_cuda(nb, "bctc", x, y_soft, il, tl)
    assert_parity(loss, grad, oracle("bctc", x, y_soft, il, tl))


@pytest.mark.parametrize("kind", ["ctc", "bctc"])
@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_reductions(nb, kind, reduction):
    if kind == "ctc":
        x, tg, il, tl = make_ctc_case(5, 40, 6, 20, 9)
    else:
        x, tg, il, tl = make_bctc_case(5, 40, 6, 20, 9, density=0.1)
    loss, grad = run_cuda(nb, kind, x, tg, il, tl, reduction=reduction)
    ref = oracle(kind, x, tg, il, tl, reduction)
    assert_parity(loss, grad, ref)


def test_edge_lengths(nb):
    """L_b = 1, L_b = T_b (single admissible path), T_b < T (zero rows), -1 padding, int64 labels."""
    T, B, C, L = 12, 5, 9, 12
    rs = np.random.RandomState(3)
    x = rs.standard_normal((T, B, C)).astype(np.float32)
    tl = np.array([1, 12, 5, 5, 3])
    il = np.array([12, 12, 5, 9, 3])
    lab = rs.randint(0, C, size=(B, L)).astype(np.int64)
    for b in range(B):
        lab[b, tl[b]:] = -1
    loss, grad = run_cuda(nb, "ctc", x, lab, il, tl, lab_dtype=torch.int64)
    ref = oracle("ctc", x, lab, il, tl)
    assert_parity(loss, grad, ref)
    for b in range(B):
        assert np.all(grad[il[b]:, b] == 0.0)           # exact zeros beyond input_length (quirk 4)
    # single admissible path: gamma is one-hot => grad = (softmax - onehot)/B
    b = 1
    sm = np.exp(R.log_softmax(x[:, b].astype(np.float64)))
    oh = np.zeros_like(sm)
    oh[np.arange(T), lab[b, :T]] = 1.0
    np.testing.assert_allclose(grad[:, b], (sm - oh) / B, atol=2e-7)


def test_infeasible_sequences(nb):
    """Outside the parity domain (T_b < L_b, L_b = 0, label out of range): loss = +inf, grad = 0,
    and the other sequences of the batch are unaffected."""
    T, B, C, L = 8, 5, 6, 10
    x, lab, il, tl = make_ctc_case(9, T, B, C, L, ragged_T=False)
    tl[:] = [3, 10, 0, 2, 2]
    il[:] = [8, 8, 8, 8, 8]
    lab[:] = np.random.RandomState(1).randint(0, C, size=(B, L))
    lab[3, 1] = C + 3
    m = nb.NoBlankCTC(reduction="none")
    xt = torch.tensor(x, device=DEV, requires_grad=True)
    per = m(xt, torch.tensor(lab, device=DEV), torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV))
    per[[0, 4]].sum().backward()
    torch.cuda.synchronize()
    per = per.detach().cpu().numpy()
    assert np.isinf(per[1]) and np.isinf(per[2]) and np.isinf(per[3])
    g = xt.grad.cpu().numpy()
    assert np.all(g[:, 1:4] == 0.0)
    ok = [0, 4]
    ref = R.nbctc_loss_grad(x[:, ok], lab[ok], il[ok], tl[ok], "none")
    np.testing.assert_allclose(per[ok], ref["per_seq"], rtol=TOL)
    assert rel_l2(g[:, ok], ref["grad"]) < TOL


def test_upstream_gradient_scaling_and_no_grad(nb):
    x, lab, il, tl = make_ctc_case(17, 30, 4, 11, 6)
    ref = oracle("ctc", x, lab, il, tl)
    m = nb.NoBlankCTC()
    args = (torch.tensor(lab, device=DEV), torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV))
    xt = torch.tensor(x, device=DEV, requires_grad=True)
    (2.5 * m(xt, *args)).backward()
    assert rel_l2(xt.grad.cpu().numpy(), 2.5 * ref["grad"]) < TOL
    with torch.no_grad():
        l2 = m(xt, *args)
    assert not l2.requires_grad and abs(float(l2) - ref["loss"]) / ref["loss"] < TOL
    # validate(): logits without grad
    l3 = m(xt.detach(), *args)
    assert not l3.requires_grad
    # float64 logits are accepted; the gradient comes back in the input dtype
    xd = torch.tensor(x, device=DEV, dtype=torch.float64, requires_grad=True)
    m(xd, *args).backward()
    assert xd.grad.dtype == torch.float64 and rel_l2(xd.grad.cpu().numpy(), ref["grad"]) < TOL
    # non-contiguous logits (B,T,C)->(T,B,C) view
    xb = torch.tensor(np.ascontiguousarray(x.transpose(1, 0, 2)), device=DEV, requires_grad=True)
    m(xb.transpose(0, 1), *args).backward()
    assert rel_l2(xb.grad.cpu().numpy().transpose(1, 0, 2), ref["grad"]) < TOL
    assert len(m.state_dict()) == 0


def test_fused_matches_generic_bitwise_properties(nb):
    """Size-independent properties at a larger size: rows of the gradient sum to zero
    (softmax and gamma both sum to one), zero rows beyond input_length, default == generic."""
    T, B, C, L = 256, 512, 157, 32
    x, lab, il, tl = make_ctc_case(1234, T, B, C, L, ragged_T=True)
    l0, g0 = run_cuda(nb, "ctc", x, lab, il, tl, flags=0)
    l1, g1 = run_cuda(nb, "ctc", x, lab, il, tl, flags=1)
    assert abs(l0 - l1) / abs(l1) < 1e-6
    assert rel_l2(g0, g1) < 2e-6
    rowsum = np.abs(g0.sum(axis=2)) * B
    assert rowsum.max() < 1e-5
    live = np.arange(T)[:, None] < il[None, :]
    assert np.all(g0[~live] == 0.0)
    sub = np.random.RandomState(0).choice(B, 24, replace=False)
    ref = R.nbctc_loss_grad(x[:, sub], lab[sub], il[sub], tl[sub], "sum")
    assert rel_l2(g0[:, sub] * B, ref["grad"]) < TOL


def test_best_path_and_argmax_bit_exact(nb):
    for seed, (T, B, C, L) in enumerate([(4, 2, 5, 3), (64, 8, 157, 8), (200, 6, 33, 40), (90, 3, 1024, 64)]):
        x, lab, il, tl = make_ctc_case(40 + seed, T, B, C, L, ragged_T=True, dup=(seed % 2 == 1))
        st, sc, am = nb.best_path(torch.tensor(x, device=DEV), torch.tensor(lab, device=DEV),
                                  torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV))
        torch.cuda.synchronize()
        rst, rsc = R.best_path(x, lab, il, tl)
        assert np.array_equal(st.cpu().numpy(), rst)
        assert np.array_equal(sc.cpu().numpy(), rsc)          # float64 add/max only: bit-exact
        assert np.array_equal(am.cpu().numpy(), R.frame_argmax(x))


def test_kat_c_alignment(nb):
    from conftest import load_golden
    a = load_golden("kat_a_ctc")
    st, sc, am = nb.best_path(torch.tensor(a["logits"], device=DEV), torch.tensor(a["targets"], device=DEV),
                              torch.tensor(a["input_length"], device=DEV), torch.tensor(a["target_length"], device=DEV))
    assert st.cpu().tolist() == [[0, 0, 1, 2], [0, 0, 0, 1]]
    assert am.cpu().numpy().T.tolist() == [[1, 4, 2, 3], [4, 1, 3, 2]]


def test_c_abi_host_entry_points(nb):
    """Call the C ABI directly with HOST buffers (numpy) -- no torch types cross the boundary."""
    import ctypes as C
    from ctc_b200 import _ffi
    lib = _ffi.lib()
    x, lab, il, tl = make_ctc_case(3, 20, 4, 13, 5)
    T, B, Cc = x.shape
    per = np.zeros(B, np.float32)
    s64 = np.zeros(1, np.float64)
    red = np.zeros(1, np.float32)
    g = np.zeros_like(x)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.nbctc_loss_grad_host_f32(0, p(x), T, B, Cc, p(lab), lab.shape[1], p(il), p(tl), p(per), p(s64), p(red),
                                      p(g), 1.0 / B, 0)
    assert rc == 0, lib.nbctc_last_error()
    ref = R.nbctc_loss_grad(x, lab, il, tl)
    np.testing.assert_allclose(per, ref["per_seq"], rtol=TOL)
    assert abs(red[0] - ref["loss"]) / ref["loss"] < TOL and abs(s64[0] / B - ref["loss"]) / ref["loss"] < TOL
    assert rel_l2(g, ref["grad"]) < TOL
    xb, y, ilb, tlb = make_bctc_case(4, 20, 4, 13, 5, density=0.2)
    rc = lib.nbbctc_loss_grad_host_f32(0, p(xb), T, B, Cc, p(y), y.shape[1], p(ilb), p(tlb), p(per), p(s64), p(red),
                                       p(g), 1.0 / B, 0)
    assert rc == 0, lib.nbctc_last_error()
    refb = R.nbbctc_loss_grad(xb, y, ilb, tlb)
    assert abs(red[0] - refb["loss"]) / refb["loss"] < TOL and rel_l2(g, refb["grad"]) < TOL
    # error path: bad shape -> negative code + message, nothing thrown
    rc = lib.nbctc_loss_grad_host_f32(0, p(x), 0, B, Cc, p(lab), lab.shape[1], p(il), p(tl), p(per), p(s64), p(red),
                                      p(g), 1.0, 0)
    assert rc == -1 and b"invalid shape" in lib.nbctc_last_error()


def test_long_sequence_accuracy(nb):
    """T = 4096 (BASELINE configs[3] length): the float64 lattice state keeps the 1e-5 bar."""
    x, lab, il, tl = make_ctc_case(77, 4096, 2, 64, 200, ragged_T=True)
    loss, grad = run_cuda(nb, "ctc", x, lab, il, tl)
    assert_parity(loss, grad, oracle("ctc", x, lab, il, tl))


def _device_call(nb, x, lab, il, tl, seq_w=None, w_scalar=1.0, offset_floats=0):
    """nbctc_loss_grad_f32 through ctypes with DEVICE pointers (what a non-PyTorch host would do)."""
    from ctc_b200 import _ffi
    lib = _ffi.lib()
    T, B, C = x.shape
    # optional 4-byte offsets: the logits / gradient pointers are then not 16-byte aligned
    xs = torch.empty(x.size + offset_floats, device=DEV)
    xd = xs[offset_floats:].view(T, B, C)
    xd.copy_(torch.tensor(x))
    gs = torch.full((x.size + offset_floats,), float("nan"), device=DEV)
    gd = gs[offset_floats:].view(T, B, C)
    labd = torch.tensor(lab, device=DEV, dtype=torch.int32)
    ild, tld = torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV)
    per = torch.empty(B, device=DEV)
    swd = None if seq_w is None else torch.tensor(seq_w, device=DEV, dtype=torch.float32)
    wsb = int(lib.nbctc_workspace_bytes(T, B, C, lab.shape[1], 0, 0))
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=DEV)
    rc = lib.nbctc_loss_grad_f32(xd.data_ptr(), T, B, C, labd.data_ptr(), lab.shape[1], ild.data_ptr(), tld.data_ptr(),
                                 per.data_ptr(), None, None, gd.data_ptr(), None if swd is None else swd.data_ptr(),
                                 w_scalar, ws.data_ptr(), wsb, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.nbctc_last_error()
    torch.cuda.synchronize()
    return per.cpu().double().numpy(), gd.cpu().double().numpy()


def test_sequence_weights_zero_and_negative(nb):
    """w_b = weight_scalar * seq_weights[b] scales the gradient; 0 gives exact zeros and the loss is still right."""
    x, lab, il, tl = make_ctc_case(21, 50, 6, 37, 9, ragged_T=True, dup=True)
    sw = np.array([1.0, 0.0, -2.0, 0.5, 0.0, 3.0], np.float32)
    per, grad = _device_call(nb, x, lab, il, tl, seq_w=sw, w_scalar=0.25)
    ref = oracle("ctc", x, lab, il, tl, "sum")
    np.testing.assert_allclose(per, ref["per_seq"], rtol=TOL)
    want = ref["grad"] * (0.25 * sw)[None, :, None]
    assert rel_l2(grad, want) < TOL
    assert np.all(grad[:, sw == 0] == 0.0)


@pytest.mark.parametrize("off", [1, 2, 3])
def test_unaligned_tensors_take_the_generic_path(nb, off):
    """The fused kernel needs 16-byte aligned logits/grad; a 4-byte offset view must still give the right answer."""
    x, lab, il, tl = make_ctc_case(31, 40, 5, 21, 7, ragged_T=True)
    per, grad = _device_call(nb, x, lab, il, tl, offset_floats=off)
    ref = oracle("ctc", x, lab, il, tl, "sum")
    np.testing.assert_allclose(per, ref["per_seq"], rtol=TOL)
    assert rel_l2(grad, ref["grad"]) < TOL


def _device_call_bin(nb, x, y, il, tl, seq_w=None, w_scalar=1.0, offset_floats=0, grad_offset_floats=0, flags=0):
    """nbbctc_loss_grad_f32 through ctypes with DEVICE pointers; optional 4-byte offsets of logits / gradient."""
    from ctc_b200 import _ffi
    lib = _ffi.lib()
    T, B, C = x.shape
    xs = torch.empty(x.size + offset_floats, device=DEV)
    xd = xs[offset_floats:].view(T, B, C)
    xd.copy_(torch.tensor(x))
    gs = torch.full((x.size + grad_offset_floats,), float("nan"), device=DEV)
    gd = gs[grad_offset_floats:].view(T, B, C)
    yd = torch.tensor(y, device=DEV, dtype=torch.float32)
    ild, tld = torch.tensor(il, device=DEV), torch.tensor(tl, device=DEV)
    per = torch.full((B,), float("nan"), device=DEV)
    swd = None if seq_w is None else torch.tensor(seq_w, device=DEV, dtype=torch.float32)
    wsb = int(lib.nbctc_workspace_bytes(T, B, C, y.shape[1], 1, flags))
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=DEV)
    rc = lib.nbbctc_loss_grad_f32(xd.data_ptr(), T, B, C, yd.data_ptr(), y.shape[1], ild.data_ptr(), tld.data_ptr(),
                                  per.data_ptr(), None, None, None if flags & 2 else gd.data_ptr(),
                                  None if swd is None else swd.data_ptr(), w_scalar, ws.data_ptr(), wsb, flags,
                                  torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.nbctc_last_error()
    torch.cuda.synchronize()
    return per.cpu().double().numpy(), gd.cpu().double().numpy()


@pytest.mark.parametrize("off", [(0, 0), (1, 0), (0, 3), (2, 1)], ids=lambda o: "x%d_g%d" % o)
def test_bctc_tiled_path_pointer_alignment(nb, off):
    """The tiled multi-label path stages logits rows with 16-byte aligned bulk copies: an unaligned logits pointer must
    take the generic kernels; the gradient pointer may have any 4-byte alignment; ragged input lengths."""
    x, y, il, tl = make_bctc_case(77, 45, 6, 157, 11, density=0.04)
    il = np.array([45, 44, 17, 11, 30, 45], np.int64)
    tl = np.minimum(tl, il)
    per, grad = _device_call_bin(nb, x, y, il, tl, offset_floats=off[0], grad_offset_floats=off[1])
    ref = oracle("bctc", x, y, il, tl, "sum")
    np.testing.assert_allclose(per, ref["per_seq"], rtol=TOL)
    assert rel_l2(grad, ref["grad"]) < TOL
    for b in range(6):
        assert np.all(grad[il[b]:, b] == 0.0)


def test_bctc_tiled_path_weights_and_no_grad(nb):
    x, y, il, tl = make_bctc_case(78, 70, 5, 64, 9, density=0.06)
    sw = np.array([1.0, 0.0, -2.0, 0.5, 3.0], np.float32)
    per, grad = _device_call_bin(nb, x, y, il, tl, seq_w=sw, w_scalar=0.25)
    ref = oracle("bctc", x, y, il, tl, "sum")
    np.testing.assert_allclose(per, ref["per_seq"], rtol=TOL)
    assert rel_l2(grad, ref["grad"] * (0.25 * sw)[None, :, None]) < TOL
    assert np.all(grad[:, sw == 0] == 0.0)
    per2, grad2 = _device_call_bin(nb, x, y, il, tl, flags=2)  # NBCTC_FLAG_NO_GRAD: loss only, gradient untouched
    np.testing.assert_allclose(per2, ref["per_seq"], rtol=TOL)
    assert np.all(np.isnan(grad2))


@pytest.mark.parametrize("seed", range(24))
def test_bctc_tiled_path_random_shapes(nb, seed):
    """Random small shapes through the tiled multi-label path: odd and even Lmax (register-fed and staged lattice
    tiles), T below / across the 8-step tiles and the 4-row batches, tiny and wide class dimensions, B = 1."""
    rs = np.random.RandomState(1000 + seed)
    T = int(rs.choice([1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 33, 64, 70, 129, 260]))
    B = int(rs.choice([1, 2, 3, 5, 9]))
    C = int(rs.choice([1, 2, 3, 4, 5, 31, 32, 33, 64, 100, 157, 255, 256]))
    L = int(rs.choice([1, 2, 3, 8, 15, 16, 31, 32, 33, 40, 64, 65, 130]))
    dens = float(rs.choice([0.0, 0.02, 0.1, 0.3]))
    if dens * C > 20:
        dens = 20.0 / C  # stay on the tiled path (at most 31 classes per state)
    x, y, il, tl = make_bctc_case(3000 + seed, T, B, C, L, density=dens, ragged_T=bool(seed & 1))
    x *= float(rs.choice([0.1, 1.0, 4.0]))
    per, grad = _device_call_bin(nb, x, y, il, tl)
    ref = oracle("bctc", x, y, il, tl, "sum")
    np.testing.assert_allclose(per, ref["per_seq"], rtol=TOL)
    assert rel_l2(grad, ref["grad"]) < TOL
    assert np.all(np.isfinite(grad))


def test_bctc_tiled_path_last_row_ends_inside_a_chunk(nb):
    """T*B*C*4 not a multiple of 16: the bulk copy of the tensor's last row stops early and the rest goes by hand."""
    for (T, B, C) in ((9, 3, 7), (8, 1, 157), (17, 5, 33)):
        x, y, il, tl = make_bctc_case(79 + C, T, B, C, 4, density=0.2)
        il[:] = T
        per, grad = _device_call_bin(nb, x, y, il, tl)
        ref = oracle("bctc", x, y, il, tl, "sum")
        np.testing.assert_allclose(per, ref["per_seq"], rtol=TOL)
        assert rel_l2(grad, ref["grad"]) < TOL


@pytest.mark.parametrize("flags", [8, 32], ids=["lockstep", "seqwarp"])
@pytest.mark.parametrize("shape", [(19, 3, 7, 5), (23, 7, 157, 12), (16, 6, 66, 9), (9, 13, 5, 4), (40, 9, 1030, 20)],
                         ids=lambda s: "T%d_B%d_C%d_L%d" % s)
def test_slab_alignment_phases(nb, shape, flags):
    """B*C not a multiple of 4: the 16-byte phase of a time step's rows changes with t; partial last groups;
    tensors whose byte size is not a multiple of 16."""
    T, B, C, L = shape
    x, lab, il, tl = make_ctc_case(500 + T + C, T, B, C, L, ragged_T=True, dup=True)
    loss, grad = run_cuda(nb, "ctc", x, lab, il, tl, flags=flags)
    assert_parity(loss, grad, oracle("ctc", x, lab, il, tl))


def test_aligned16_flag_contract(nb):
    """NBCTC_FLAG_ALIGNED16: smaller workspace query; a broken promise is an error, not a wrong answer."""
    from ctc_b200 import _ffi
    lib = _ffi.lib()
    T, B, C, L = 64, 16, 157, 8
    full = int(lib.nbctc_workspace_bytes(T, B, C, L, 0, 0))
    lean = int(lib.nbctc_workspace_bytes(T, B, C, L, 0, _ffi.FLAG_ALIGNED16))
    assert 0 < lean <= full
    x, lab, il, tl = make_ctc_case(5, T, B, C, L)
    xs = torch.empty(x.size + 1, device=DEV)
    xd = xs[1:].view(T, B, C)          # 4-byte offset: not 16-byte aligned
    xd.copy_(torch.tensor(x))
    per = torch.empty(B, device=DEV)
    grad = torch.empty((T, B, C), device=DEV)
    ws = torch.empty(full, dtype=torch.uint8, device=DEV)
    args = (xd.data_ptr(), T, B, C, torch.tensor(lab, device=DEV).data_ptr(), L, torch.tensor(il, device=DEV).data_ptr(),
            torch.tensor(tl, device=DEV).data_ptr(), per.data_ptr(), None, None, grad.data_ptr(), None, 1.0 / B,
            ws.data_ptr(), full)
    rc = lib.nbctc_loss_grad_f32(*args, _ffi.FLAG_ALIGNED16, torch.cuda.current_stream().cuda_stream)
    assert rc == -1 and b"16-byte aligned" in lib.nbctc_last_error()


@pytest.mark.parametrize("flags", [1, 8, 32], ids=["generic", "lockstep", "seqwarp"])
def test_repeated_labels_are_bit_reproducible(nb, flags):
    """the gammas of states that share a class accumulate.  Every path adds them in ascending state
    order (rank rounds / follower walk, no atomics): two runs give identical bits."""
    T, B, C, L = 60, 40, 7, 32
    x, lab, il, tl = make_ctc_case(77, T, B, C, L, dup=True, Lmin=16)
    l1, g1 = run_cuda(nb, "ctc", x, lab, il, tl, flags=flags)
    l2, g2 = run_cuda(nb, "ctc", x, lab, il, tl, flags=flags)
    assert np.array_equal(l1, l2) and np.array_equal(g1, g2)
    assert_parity(l1, g1, oracle("ctc", x, lab, il, tl))


SINGLE_PATH = [
    # Lmax, target lengths (all but the last sequence have T_b == L_b: one admissible path, gamma one-hot, and alpha AND
    # beta of every state on the path are "mass that has just arrived" in their lanes' scales)
    (32, [5, 17, 23, 32, 20]), (40, [5, 17, 23, 33, 40, 20]), (64, [40, 50, 64, 30]), (100, [70, 100, 35, 50]),
    (256, [130, 200, 256, 100]),
]


def _single_path_case(kind, Lmax, Ls, boost, seed=0):
    rs = np.random.RandomState(seed + Lmax)
    B, T = len(Ls), max(Ls)
    C = 100 if kind == "bctc" else 64
    x = rs.standard_normal((T, B, C)).astype(np.float32)
    tl = np.array(Ls, dtype=np.int64)
    il = tl.copy()
    il[-1] = T
    if kind == "bctc":
        y = (rs.uniform(size=(B, Lmax, C)) < 0.05).astype(np.float32)
        y[np.arange(B)[:, None], np.arange(Lmax)[None, :], rs.randint(0, C, size=(B, Lmax))] = 1.0
        for b in range(B):
            y[b, tl[b]:] = 0.0
        return x, y, il, tl
    lab = rs.randint(0, C, size=(B, Lmax)).astype(np.int32)
    for b in range(B):
        lab[b, tl[b]:] = -1
        if boost and il[b] == tl[b]:
            x[np.arange(tl[b]), b, lab[b, :tl[b]]] += boost   # emissions near one along the forced path
    return x, lab, il, tl


@pytest.mark.parametrize("boost", [0.0, 15.0], ids=["plain", "peaked_on_path"])
@pytest.mark.parametrize("flags", [1, 8, 32], ids=["generic", "lockstep", "seqwarp"])
@pytest.mark.parametrize("case", SINGLE_PATH, ids=lambda c: "L%d" % c[0])
def test_single_admissible_path_ctc(nb, case, flags, boost):
    Lmax, Ls = case
    x, lab, il, tl = _single_path_case("ctc", Lmax, Ls, boost)
    loss, grad = run_cuda(nb, "ctc", x, lab, il, tl, reduction="sum", flags=flags)
    ref = oracle("ctc", x, lab, il, tl, "sum")
    assert_parity(loss, grad, ref, scale=1e-3)
    # every sequence on its own (a wrong short sequence hides in the batch norm).  With emissions near one the gradient
    # of a forced path is 1 - p ~ 2e-5 itself, below what a float32 softmax resolves to 1e-5 relative: absolute bar there
    for b in range(len(Ls) - 1):
        if boost:
            assert np.max(np.abs(grad[:, b] - ref["grad"][:, b])) < 2e-6      # a few float32 ulps of the O(1) terms w p and gamma
        else:
            assert rel_l2(grad[:, b], ref["grad"][:, b]) < 1e-5


@pytest.mark.parametrize("flags", [0, 1], ids=["default", "generic"])
@pytest.mark.parametrize("case", SINGLE_PATH, ids=lambda c: "L%d" % c[0])
def test_single_admissible_path_bctc(nb, case, flags):
    """Found in round 2: emissions of the multi-label variant are near one, so alpha and beta stay large in their lanes'
    scales along a forced path and the single power-of-two factor of the beta entries underflowed (gamma = 0)."""
    Lmax, Ls = case
    x, y, il, tl = _single_path_case("bctc", Lmax, Ls, 0.0)
    loss, grad = run_cuda(nb, "bctc", x, y, il, tl, reduction="sum", flags=flags)
    ref = oracle("bctc", x, y, il, tl, "sum")
    assert_parity(loss, grad, ref)
    for b in range(len(Ls)):
        assert rel_l2(grad[:, b], ref["grad"][:, b]) < 1e-5


@pytest.mark.parametrize("flags", [8, 32], ids=["lockstep", "seqwarp"])
@pytest.mark.parametrize("shape", [(64, 8, 157, 20), (50, 6, 157, 40), (40, 3, 512, 40), (70, 4, 64, 150), (33, 5, 1024, 256),
                                   (12, 301, 33, 9)],
                         ids=lambda s: "T%d_B%d_C%d_L%d" % s)
def test_logit_gaps_beyond_the_float32_emission_floor(nb, shape, flags):
    """Logits whose label entries lie more than 83 nats under the row maximum: softmax(x)[label] is below 2^-120, which the
    linear-domain kernels cannot hold.  They flag such sequences and the log-domain repair kernel redoes them; the other
    sequences of the batch (every second one keeps N(0,1) logits) stay on the fast path."""
    T, B, C, Lmax = shape
    x, lab, il, tl = make_ctc_case(900 + T + C, T, B, C, Lmax)
    x[:, 0::2] *= 40.0
    loss, grad = run_cuda(nb, "ctc", x, lab, il, tl, reduction="none", flags=flags)
    ref = oracle("ctc", x, lab, il, tl, "none")
    assert np.all(np.isfinite(loss))
    assert np.max(np.abs(loss - ref["per_seq"]) / np.abs(ref["per_seq"])) < TOL
    for b in range(B):
        gb, rb = grad[:, b], ref["grad"][:, b]
        assert rel_l2(gb, rb) < TOL, f"sequence {b}"
        assert np.max(np.abs(gb - rb)) < TOL * max(np.max(np.abs(rb)), 1e-3)
        assert np.all(gb[il[b]:] == 0.0)
