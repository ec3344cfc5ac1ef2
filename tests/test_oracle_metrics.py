"""numpy metric restatement (oracle/metrics_oracle.py) vs reference-generated fixtures and the
known-answer vectors of SURVEY.md Appendix B."""
import os

import numpy as np
import pytest

from oracle import metrics_oracle as mo


def test_ged_known_answers():
    s = np.zeros((3, 2, 4), np.int64)
    s[0, 0, :2] = 1
    s[1, 0, :1] = 1
    y = np.zeros((2, 2, 4), np.int64)
    y[0, 0, :3] = 1
    assert mo.generalised_energy_distance(s, y, nlabels=1, label_range=range(1, 2)) == pytest.approx(5 / 18, abs=1e-15)
    assert mo.generalised_energy_distance(s, s, nlabels=1, label_range=range(1, 2)) == pytest.approx(0.0, abs=1e-15)
    z2, z3 = np.zeros((2, 2, 4), np.int64), np.zeros((3, 2, 4), np.int64)
    assert mo.generalised_energy_distance(z2, z3, nlabels=1, label_range=range(1, 2)) == 0.0
    s4 = np.array([[[1, 2, 2, 0]], [[1, 1, 0, 0]]])
    y4 = np.array([[[1, 2, 0, 0]]])
    assert mo.generalised_energy_distance(s4, y4, nlabels=2, label_range=range(1, 3)) == pytest.approx(0.625, abs=1e-15)


def test_ncc_identity():
    """The reference's only unit test (test/test_scores.py:31-50) asserts NCC(gt, gt) == 1 on one LIDC image whose
    annotators agree.  The identity that holds for ANY input is the single-sample one: E_ss == E_sy when N == M == 1."""
    rs = np.random.RandomState(0)
    logits = rs.standard_normal((1, 2, 16, 16)).astype(np.float32)
    p = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
    np.testing.assert_allclose(mo.variance_ncc_dist(p, p), [1.0], rtol=1e-6)
    # agreeing annotators, several samples scattered around them: still a positive correlation
    gt = (rs.uniform(size=(1, 1, 16, 16)) < 0.4).astype(np.int64).repeat(4, 0)
    onehot = mo.convert_batch_to_onehot(gt, 2)
    noisy = np.clip(onehot[:3].astype(np.float32) + 0.2 * rs.uniform(size=(3, 2, 16, 16)).astype(np.float32), 0, 1)
    noisy /= noisy.sum(1, keepdims=True)
    assert mo.variance_ncc_dist(noisy, onehot)[0] > 0


def test_ncc_known_answer():
    """SURVEY.md Appendix B NCC-2."""
    import torch
    torch.manual_seed(1)
    p = torch.softmax(torch.randn(3, 2, 2, 3), 1).numpy()
    gt = np.array([[[1, 0, 1], [0, 0, 1]], [[1, 1, 1], [0, 0, 0]]])
    onehot = mo.convert_batch_to_onehot(gt[:, None], 2)
    out = mo.variance_ncc_dist(p, onehot)
    assert out.shape == (1,) and out.dtype == np.float64
    np.testing.assert_allclose(out, [0.25239795], rtol=1e-6)


@pytest.mark.parametrize('tag,C', [('bin', 2), ('tri', 3)])
def test_metrics_match_reference_fixture(golden_dir, tag, C):
    g = np.load(os.path.join(golden_dir, 'metrics.npz'))
    samples, gts = g[tag + '_samples'].astype(np.int64), g[tag + '_gts'].astype(np.int64)
    ged = mo.generalised_energy_distance(samples, gts, nlabels=C - 1, label_range=range(1, C))
    assert ged == pytest.approx(float(g[tag + '_ged']), abs=1e-12)
    logits = g[tag + '_logits']
    e = np.exp(logits - logits.max(1, keepdims=True))
    import torch
    probs = torch.softmax(torch.from_numpy(logits), 1).numpy()
    onehot = mo.convert_batch_to_onehot(gts[:, None], C)
    np.testing.assert_array_equal(onehot, g[tag + '_onehot'])
    np.testing.assert_allclose(mo.variance_ncc_dist(probs, onehot), g[tag + '_ncc'], rtol=1e-6)
