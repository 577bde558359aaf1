"""CPU: real-signal regression (SURVEY 8c iv) -- 1.5 s of the reference's bundled saxophone recording
with the parameters of examples/WavResynth.py (nfft 4096, npks 100, hop 1024), outputs of the real
reference in tests/golden/real_wav.npz (oracle/gen_golden_wav.py).  The oracle must reproduce them
bit for bit; the kernels (SIMT emulator build of the CUDA sources) within the stated tolerances, peak
bins exact wherever the decision margin exceeds fp32 round-off."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

from oracle import pv_oracle as orc
import parity_util as pu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# saxophone @ WavResynth.py's parameters; guitar @ the metric's (nfft 2048 / hop 512 / npks 50); speech @ cfg3's
CASES = {"sax": ("real_wav.npz", dict(nfft=4096, hop=1024, npks=100)),
         "guitar": ("real_guitar.npz", dict(nfft=2048, hop=512, npks=50)),
         "speech": ("real_speech.npz", dict(nfft=512, hop=128, npks=20))}


def load(case):
    G = np.load(os.path.join(GOLD, CASES[case][0]))
    return G, CASES[case][1], (G["pcm"] / float(np.iinfo(np.int16).max)).astype(np.float32), int(G["sr"])


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_on_real_signal(case):
    G, PVKW, x, sr = load(case)
    o = orc.analyze(x, sr, **PVKW)
    for k in ("f", "mag", "ph", "realph", "binno"):
        assert np.array_equal(o[k], G[k], equal_nan=True), k
    assert np.array_equal(np.array(o["totalmag"]), G["totalmag"])
    tr = orc.track(G["f"], G["mag"])
    assert np.array_equal(tr["tid"], G["tid"]) and np.array_equal(tr["st"], G["st"]) and np.array_equal(tr["end"], G["end"])
    parts = orc.partials_from_tracks(tr, G["f"], G["mag"], G["ph"], G["realph"])
    w = orc.synth(parts, sr, PVKW["hop"], PVKW["nfft"], PVKW["hop"])
    assert w.shape == G["synth"].shape and np.max(np.abs(w - G["synth"])) == 0.0


@pytest.mark.parametrize("case", sorted(CASES))
def test_emu_kernels_on_real_signal(case):
    eh = pytest.importorskip("emu_harness")
    eh.build()
    G, PVKW, x, sr = load(case)
    o = eh.analyze(x, sr, PVKW["nfft"], PVKW["hop"], PVKW["npks"], spectra=True)
    got = {k: o[k][0] for k in ("f", "mag", "ph", "realph", "binno", "totalmag", "npk")}
    ref = {k: G[k] for k in ("f", "mag", "ph", "realph", "binno", "totalmag")}
    margin = orc.analyze(x, sr, margins=True, **PVKW)["margin"]
    rep = pu.compare_analysis(got, ref, sr, PVKW["nfft"], margin=margin)
    assert rep["mismatched"] <= 2, rep                       # frames below the fp32 margin, if any
    oo = orc.analyze(np.zeros(1), sr, fx_given=o["fx"][0].astype(np.complex64), **PVKW)
    pu.compare_exact_on_spectrum(got, oo)                    # kernel logic on its own spectrum: exact
    tr = eh.track(G["f"], G["mag"])
    assert np.array_equal(tr["tid"][0], G["tid"])
    pk = eh.track_pack(G["f"], G["mag"], G["ph"], G["realph"], tr["tid"][0], tr["link"][0], int(tr["ntracks"][0]))
    assert np.array_equal(pk["tstart"], G["st"]) and np.array_equal(pk["tstart"] + pk["tlen"] - 1, G["end"])
    w = eh.resynth(tr["tid"][0], pk, sr, PVKW["hop"], PVKW["nfft"], PVKW["hop"])
    assert w.shape == G["synth"].shape and pu.snr_db(w, G["synth"]) > 110.0
