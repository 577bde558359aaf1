"""CPU: the non-default arguments of the path -- PV(wind=np.blackman, pkthresh=0.02),
SinSum.add_frame(maxpitchjmp=0.2 / 1.0) (PV.toSinSum ignores its own argument, PVAnalysis.py:320-321),
synth(edge=0.5, minframes=5), time-stretching synth(hop=200, minframes=2) -- against
tests/golden/params.npz from the real reference (oracle/gen_golden_params.py): the oracle bit for
bit, the kernels (SIMT emulator build) within the stated tolerances."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

from oracle import pv_oracle as orc
from pypevoc_b200 import signals
import parity_util as pu

SIGNAL = dict(sr=22050, dur=0.5, f0=300, nharm=20, p=1.0, sigma=0.1, seed=31)
PVKW = dict(nfft=1024, hop=256, npks=30, pkthresh=0.02)
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "params.npz"))
SYNTH = {"synth_edge05_min5": dict(hop=256, edge=0.5, minframes=5), "synth_hop200_min2": dict(hop=200, edge=1.0, minframes=2)}


def signal():
    return np.asarray(signals.harm(**SIGNAL), dtype=np.float32)


def test_oracle_analysis_with_window_and_threshold():
    o = orc.analyze(signal(), SIGNAL["sr"], wind=np.blackman, **PVKW)
    for k in ("f", "mag", "ph", "realph", "binno"):
        assert np.array_equal(o[k], G[k], equal_nan=True), k
    assert np.array_equal(np.array(o["totalmag"]), G["totalmag"])


@pytest.mark.parametrize("mj", [0.2, 0.5, 1.0])
def test_oracle_tracking_with_maxpitchjmp(mj):
    tag = "mj%02d" % int(mj * 10)
    tr = orc.track(G["f"], G["mag"], maxpitchjmp=mj)
    assert np.array_equal(tr["tid"], G["tid_" + tag])
    assert np.array_equal(tr["st"], G["st_" + tag]) and np.array_equal(tr["end"], G["end_" + tag])


def test_tosinsum_argument_is_not_forwarded_in_the_reference():
    assert np.array_equal(G["tid_tosinsum_arg02"], G["tid_mj05"])
    assert not np.array_equal(G["tid_mj02"], G["tid_mj05"]) and not np.array_equal(G["tid_mj10"], G["tid_mj05"])


@pytest.mark.parametrize("key", sorted(SYNTH))
def test_oracle_resynthesis_with_edge_and_minframes(key):
    kw = SYNTH[key]
    tr = orc.track(G["f"], G["mag"])
    parts = orc.partials_from_tracks(tr, G["f"], G["mag"], G["ph"], G["realph"])
    w = orc.synth(parts, SIGNAL["sr"], kw["hop"], PVKW["nfft"], PVKW["hop"], edge=kw["edge"], minframes=kw["minframes"])
    assert w.shape == G[key].shape and np.max(np.abs(w - G[key])) == 0.0


def test_emu_kernels_with_nondefault_arguments():
    eh = pytest.importorskip("emu_harness")
    eh.build()
    sr = SIGNAL["sr"]
    o = eh.analyze(signal(), sr, PVKW["nfft"], PVKW["hop"], PVKW["npks"], pkthresh=PVKW["pkthresh"], wind=np.blackman)
    got = {k: o[k][0] for k in ("f", "mag", "ph", "realph", "binno", "totalmag")}
    ref = {k: G[k] for k in ("f", "mag", "ph", "realph", "binno", "totalmag")}
    margin = orc.analyze(signal(), sr, wind=np.blackman, margins=True, **PVKW)["margin"]
    pu.compare_analysis(got, ref, sr, PVKW["nfft"], margin=margin)
    for mj in (0.2, 0.5, 1.0):
        tr = eh.track(G["f"], G["mag"], maxpitchjmp=mj)
        assert np.array_equal(tr["tid"][0], G["tid_mj%02d" % int(mj * 10)]), mj
    tr = eh.track(G["f"], G["mag"])
    pk = eh.track_pack(G["f"], G["mag"], G["ph"], G["realph"], tr["tid"][0], tr["link"][0], int(tr["ntracks"][0]))
    for key, kw in SYNTH.items():
        w = eh.resynth(tr["tid"][0], pk, sr, kw["hop"], PVKW["nfft"], PVKW["hop"], edge=kw["edge"], minframes=kw["minframes"])
        assert w.shape == G[key].shape and pu.snr_db(w, G[key]) > 110.0, key
