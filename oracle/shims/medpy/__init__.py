"""Shim for MedPy==0.4.0 (reference requirements.txt:17); only ``medpy.metric.jc`` / ``dc`` are used
(utils.py:5,166; train_model.py:13,222,424).  PARITY UNPINNED (no reference test; integer counts make it exact)."""
