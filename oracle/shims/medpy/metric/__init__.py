import numpy as np


def jc(result, reference):
    """Binary Jaccard |A&B| / |A|B| (MedPy metric.binary.jc)."""
    result = np.atleast_1d(result.astype(bool))
    reference = np.atleast_1d(reference.astype(bool))
    intersection = np.count_nonzero(result & reference)
    union = np.count_nonzero(result | reference)
    return float(intersection) / float(union)


def dc(result, reference):
    """Binary Dice 2|A&B| / (|A|+|B|) (MedPy metric.binary.dc); 0.0 when both empty."""
    result = np.atleast_1d(result.astype(bool))
    reference = np.atleast_1d(reference.astype(bool))
    intersection = np.count_nonzero(result & reference)
    size = np.count_nonzero(result) + np.count_nonzero(reference)
    try:
        return 2.0 * intersection / float(size)
    except ZeroDivisionError:
        return 0.0
