"""numpy restatement of the reference sample-evaluation metrics -- TEST INFRASTRUCTURE.

  generalised_energy_distance   reference utils.py:148-200  (dist_fct :150-170, medpy jc at :166)
  variance_ncc_dist / ncc       reference utils.py:202-247 / :130-145
  convert_batch_to_onehot       reference utils.py:289-311

Integer work (label comparisons, intersection / union counts) is exact; the ratios are formed in
float64 in the same order as the reference (sum of per-pair distances, then the three scaled sums).
Pinned by tests/test_oracle_metrics.py against the reference's own utils.py (build container) and the
known-answer vectors of SURVEY.md Appendix B, including the reference's only unit test
(test/test_scores.py:48-50: NCC(gt, gt) == 1).
"""
import numpy as np


def pair_distance(m1, m2, label_range, nlabels):
    """utils.py:150-170: 1 - mean_lbl IoU with both-empty -> 1 and one-empty -> 0."""
    ious = []
    for lbl in label_range:
        a = (m1 == lbl)
        b = (m2 == lbl)
        na, nb = int(a.sum()), int(b.sum())
        if na == 0 and nb == 0:
            ious.append(1)
        elif na == 0 or nb == 0:
            ious.append(0)
        else:
            ious.append(float(np.count_nonzero(a & b)) / float(np.count_nonzero(a | b)))
    return 1 - (sum(ious) / nlabels)


def generalised_energy_distance(sample_arr, gt_arr, nlabels=1, label_range=None):
    """utils.py:148-200.  sample_arr [N,H,W] ints, gt_arr [M,H,W]; returns python float."""
    sample_arr = np.asarray(sample_arr)
    gt_arr = np.asarray(gt_arr)
    if label_range is None:
        label_range = range(nlabels)
    N, M = sample_arr.shape[0], gt_arr.shape[0]
    d_sy = [pair_distance(sample_arr[i], gt_arr[j], label_range, nlabels) for i in range(N) for j in range(M)]
    d_ss = [pair_distance(sample_arr[i], sample_arr[j], label_range, nlabels) for i in range(N) for j in range(N)]
    d_yy = [pair_distance(gt_arr[i], gt_arr[j], label_range, nlabels) for i in range(M) for j in range(M)]
    return (2. / (N * M)) * sum(d_sy) - (1. / N ** 2) * sum(d_ss) - (1. / M ** 2) * sum(d_yy)


def ged_pair_sums(sample_arr, gt_arr, nlabels, label_range):
    """The three sums the GED is made of (d_sy, d_ss, d_yy) -- what the CUDA kernel reduces to."""
    sample_arr = np.asarray(sample_arr)
    gt_arr = np.asarray(gt_arr)
    N, M = sample_arr.shape[0], gt_arr.shape[0]
    d_sy = sum(pair_distance(sample_arr[i], gt_arr[j], label_range, nlabels) for i in range(N) for j in range(M))
    d_ss = sum(pair_distance(sample_arr[i], sample_arr[j], label_range, nlabels) for i in range(N) for j in range(N))
    d_yy = sum(pair_distance(gt_arr[i], gt_arr[j], label_range, nlabels) for i in range(M) for j in range(M))
    return d_sy, d_ss, d_yy


def ncc(a, v):
    """utils.py:130-145 (zero_norm=True); population std; returns shape (1,) like np.correlate."""
    a = a.flatten()
    v = v.flatten()
    a = (a - np.mean(a)) / (np.std(a) * len(a))
    v = (v - np.mean(v)) / np.std(v)
    return np.correlate(a, v)


def variance_ncc_dist(sample_arr, gt_arr):
    """utils.py:202-247.  sample_arr [N,C,H,W] fp32 probabilities, gt_arr [M,C,H,W] one-hot ints.
    Returns float64 array of shape (1,)."""
    sample_arr = np.asarray(sample_arr)
    gt_arr = np.asarray(gt_arr)
    M = gt_arr.shape[0]
    mean_seg = np.mean(sample_arr, axis=0)
    log_s = np.log(sample_arr + 1e-8)                                  # fp32 logs like the reference
    e_ss = np.mean(np.stack([-1.0 * np.sum(mean_seg * l, axis=0) for l in log_s]).astype(np.float64), axis=0)
    out = []
    for j in range(M):
        e_sy = np.mean(np.stack([-1.0 * np.sum(gt_arr[j] * l, axis=0) for l in log_s]).astype(np.float64), axis=0)
        out.append(ncc(e_ss, e_sy))
    return (1 / M) * sum(out)


def convert_batch_to_onehot(lblbatch, nlabels):
    """utils.py:289-311 for [B,1,H,W] index labels -> int64 [B,nlabels,H,W]."""
    lblbatch = np.asarray(lblbatch)
    return np.stack([(lblbatch[:, 0] == k) for k in range(nlabels)], axis=1).astype(np.int64)
