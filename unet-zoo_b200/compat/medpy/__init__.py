"""Stand-in for MedPy when it is not installed (reference requirements.txt:17); see ../README.md."""
