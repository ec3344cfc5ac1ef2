"""Stand-in: imported by data/BratsProcessing/utils.py:10 (BraTS preprocessing, out of scope)."""
