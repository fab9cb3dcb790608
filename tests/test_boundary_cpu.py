"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/nbctc.h declares; the Python mirror keeps the reference signature; no CPU fallback exists."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nbctc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nbb?ctc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    g.build()
    from ctc_b200 import _ffi
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nbctc.h but not exported"
        assert n in _ffi.SIGNATURES, f"{n} has no ctypes signature in ctc_b200/_ffi.py"
    assert _ffi.lib().nbctc_version() == 100
    assert _ffi.lib().nbctc_last_error() == b""
    # shape queries are host-only and safe without a GPU
    assert _ffi.lib().nbctc_workspace_bytes(0, 1, 1, 1, 0, 0) == 0
    assert _ffi.lib().nbctc_workspace_bytes(256, 4096, 157, 32, 0, 1) > 0


def test_module_signature_matches_reference():
    import ctc_b200
    for cls in (ctc_b200.NoBlankCTC, ctc_b200.NoBlankBinaryCTC):
        m = cls()
        params = list(inspect.signature(m.forward).parameters)
        assert params[:4] == ["yseq", "label", "input_length", "target_length"]   # NoBlankCTC.py:129
        assert len(list(m.parameters())) == 0 and len(m.state_dict()) == 0


def test_no_cpu_fallback():
    import ctc_b200
    x = torch.zeros(4, 2, 5, requires_grad=True)
    with pytest.raises(ctc_b200.NbctcError):
        ctc_b200.NoBlankCTC()(x, torch.zeros(2, 2, dtype=torch.int32), torch.tensor([4, 4]), torch.tensor([2, 2]))
    with pytest.raises(ctc_b200.NbctcError):
        ctc_b200.NoBlankBinaryCTC()(x, torch.zeros(2, 2, 5), torch.tensor([4, 4]), torch.tensor([2, 2]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ctc_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle" not in txt.replace("against the float64 oracle", "").replace(
                    "oracle/restatement.py::best_path", ""), f"{fn} mentions the oracle"


def test_argument_validation():
    import ctc_b200
    from ctc_b200.function import _NoBlankCTCFunction  # noqa: F401
    with pytest.raises(ValueError):
        ctc_b200.no_blank_ctc_loss(torch.zeros(4, 5), None, None, None)
    with pytest.raises(ValueError):
        ctc_b200.no_blank_ctc_loss(torch.zeros(4, 2, 5), None, None, None, reduction="avg")
