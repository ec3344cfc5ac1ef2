"""`medpy.metric.dc` / `jc` as used by the reference caller (train_model.py:13,222,424): binary overlap of two
masks on the host.  Definitions: Dice = 2|A&B| / (|A|+|B|) (0 when both empty), Jaccard = |A&B| / |A|B|."""
import numpy as np


def _masks(result, reference):
    return np.asarray(result).astype(bool), np.asarray(reference).astype(bool)


def dc(result, reference):
    a, b = _masks(result, reference)
    total = int(a.sum()) + int(b.sum())
    return 2.0 * int((a & b).sum()) / total if total else 0.0


def jc(result, reference):
    a, b = _masks(result, reference)
    return float((a & b).sum()) / float((a | b).sum())
