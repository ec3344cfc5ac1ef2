"""Seeded synthetic inputs / weights: lives in the package (``b200.synth``, the benchmark needs it too); re-exported here
for the oracle, the golden generator and the tests."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'unet-zoo_b200')
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from b200.synth import *  # noqa: E402,F401,F403
from b200.synth import _rs  # noqa: E402,F401
