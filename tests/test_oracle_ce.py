"""Pin the oracle of the auxiliary cross-entropy to the reference's own modules: fixtures from
tests/golden/make_golden_ce.py (CrossEntropy.py:17-32 unmodified, and nn.CrossEntropyLoss of models/__init__.py:85)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, rel_l2
from oracle import restatement as R


@pytest.mark.parametrize("name", ["ce_small", "ce_charades", "ce_wide"])
@pytest.mark.parametrize("mode", ["index", "multihot"])
def test_aux_ce_oracle_matches_reference(name, mode):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    tg = z["y_index"] if mode == "index" else z["y_multihot"]
    out = R.aux_ce(z["logits"], tg, z["frame_index"], mode)
    assert abs(out["loss"] - float(z[f"loss_{mode}"])) <= 1e-12 * abs(float(z[f"loss_{mode}"]))
    assert rel_l2(out["grad"], z[f"grad_{mode}"]) < 1e-12
