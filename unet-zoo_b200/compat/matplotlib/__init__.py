"""Stand-in: imported by data/BratsProcessing/augmentation.py:8 only."""
