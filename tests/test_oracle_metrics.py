"""The metric restatement (oracle/metrics.py) against fixtures produced by the reference's own train.py helpers."""
import glob
import os

import numpy as np
import pytest

from oracle import metrics as M

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "metrics_*.npz")))


@pytest.mark.parametrize("path", FILES, ids=lambda p: os.path.basename(p)[:-4])
def test_metrics_oracle_matches_reference(path):
    g = np.load(path)
    out, tgt, time = g["output"], g["target"], int(g["time"])
    temporal = out.shape[0]
    pred = M.frame_topk(out, 5)
    acc = M.accuracy_time_flags(pred, tgt, time)
    assert np.array_equal(acc[0], g["acc_time_top1"].astype(np.int32))
    np.testing.assert_allclose(M.percentages(acc, temporal), g["acc_time"], rtol=1e-6)
    rec = M.recall_time_flags(pred, tgt, time)
    assert np.array_equal(rec[0], g["rec_time_top1"].astype(np.int32))
    np.testing.assert_allclose(M.percentages(rec, time), g["rec_time"], rtol=1e-6)
    s = M.accuracy_s_flags(pred, g["label"])
    assert np.array_equal(s[0], g["acc_s_top1"].astype(np.int32))
    np.testing.assert_allclose(M.percentages(s, temporal, (1,))[0], g["acc_s"][0], rtol=1e-6)
    m = M.accuracy_flags(pred, g["multihot"])
    assert np.array_equal(m[0], g["acc_m_top1"].astype(np.int32))
    np.testing.assert_allclose(M.percentages(m, temporal), g["acc_m"], rtol=1e-6)


def test_fixtures_exist():
    assert len(FILES) >= 4
