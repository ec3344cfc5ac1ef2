"""Stub: h5py is imported by the reference data modules (data/bratsDataset.py:3) but not on the hot path."""


class File:  # pragma: no cover - never opened on the synthetic path
    def __init__(self, *a, **k):
        raise RuntimeError('h5py stub: dataset IO is out of scope (SURVEY.md #13-15)')
