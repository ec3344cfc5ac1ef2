"""Stub: nibabel is imported by reference utils.py:7 but never used on the hot path."""
