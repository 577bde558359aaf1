"""CPU: a fixed-seed slice of the randomized kernel-vs-oracle comparisons of tests/fuzz_emu.py
(emulator build of the CUDA sources): every link kernel + id resolution against the oracle's
sequential greedy loop, the analysis kernel against the oracle on its own spectrum, pack +
resynthesis against orc.synth over random hop / edge / minframes."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

from oracle import pv_oracle as orc
import parity_util as pu
import fuzz_emu as fz

eh = pytest.importorskip("emu_harness")


def test_fuzz_tracking_kernels():
    eh.build()
    rng = np.random.RandomState(2024)
    for it in range(80):
        case = fz.track_case(rng)
        assert fz.check_track(eh, orc, case), (it, case[0].shape, case[2])


def test_fuzz_analysis_kernel():
    eh.build()
    rng = np.random.RandomState(2025)
    for it in range(40):
        case = fz.analyze_case(rng, lognfft=(6, 11))
        fz.check_analyze(eh, orc, pu, case)


def test_fuzz_pack_and_resynthesis_kernels():
    eh.build()
    rng = np.random.RandomState(2026)
    for it in range(30):
        fz.check_synth(eh, orc, pu, fz.synth_case(rng))
