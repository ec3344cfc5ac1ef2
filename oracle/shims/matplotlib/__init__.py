"""Stub: matplotlib is imported by data/BratsProcessing/augmentation.py:8 only."""
