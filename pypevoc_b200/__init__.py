"""pypevoc_b200 -- B200-native (sm_100a) implementation of PyPeVoc's phase-vocoder hot path.

Drop-in for ``pypevoc.PV`` / ``pypevoc.SinSum`` (reference pypevoc/__init__.py:1):

    from pypevoc_b200 import PV
"""
from .pv import PV, PVBatch, PVHarmonic, SinSum, SinSumBatch, RegPartial, Progress  # noqa: F401

__all__ = ["PV", "PVBatch", "PVHarmonic", "SinSum", "SinSumBatch", "RegPartial", "Progress"]
