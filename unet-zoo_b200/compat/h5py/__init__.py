"""Stand-in: imported by data/bratsDataset.py:3; dataset IO is out of scope."""


class File:
    def __init__(self, *a, **k):
        raise RuntimeError('h5py is not installed: HDF5 dataset IO is unavailable (use a synthetic data_loader)')
