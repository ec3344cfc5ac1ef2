"""Stand-in: imported by the reference utils.py:7, unused on the synthetic / hot path."""
