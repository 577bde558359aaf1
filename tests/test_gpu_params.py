"""-m gpu: the non-default arguments of the path through the product classes, against
tests/golden/params.npz from the real reference (see tests/test_params_cpu.py for the CPU side)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import pv_oracle as orc
import parity_util as pu
from test_params_cpu import G, PVKW, SIGNAL, SYNTH, signal

pytestmark = pytest.mark.gpu


def test_pv_with_window_and_threshold():
    from pypevoc_b200 import PV
    sr = SIGNAL["sr"]
    pv = PV(signal(), sr, wind=np.blackman, progress=False, **PVKW)
    pv.run_pv()
    got = dict(f=pv.f, mag=pv.mag, ph=pv.ph, realph=pv.realph, binno=pv.binno, totalmag=pv.totalmag)
    ref = {k: G[k] for k in ("f", "mag", "ph", "realph", "binno", "totalmag")}
    margin = orc.analyze(signal(), sr, wind=np.blackman, margins=True, **PVKW)["margin"]
    pu.compare_analysis(got, ref, sr, PVKW["nfft"], margin=margin)
    assert np.array_equal(pv.win, np.blackman(PVKW["nfft"]))
    # PV.toSinSum does not forward maxpitchjmp (PVAnalysis.py:320-321): same partials as the default
    if np.array_equal(pv.binno, G["binno"]):
        ss = pv.toSinSum(maxpitchjmp=0.2)
        assert np.array_equal(ss.track_ids, G["tid_tosinsum_arg02"])


@pytest.mark.parametrize("mj", [0.2, 0.5, 1.0])
def test_add_frame_with_maxpitchjmp(mj):
    """SinSum.add_frame row by row on the reference's tables (PVAnalysis.py:871-957)."""
    from pypevoc_b200 import SinSum
    tag = "mj%02d" % int(mj * 10)
    ss = SinSum(SIGNAL["sr"], nfft=PVKW["nfft"], hop=PVKW["hop"])
    for fr in range(G["f"].shape[0]):
        ss.add_frame(fr, G["f"][fr], G["mag"][fr], G["ph"][fr], realph=G["realph"][fr], maxpitchjmp=mj)
    assert np.array_equal(ss.track_ids, G["tid_" + tag])
    assert ss.st == G["st_" + tag].tolist() and ss.end == G["end_" + tag].tolist()
    assert len(ss.partial) == len(G["st_" + tag])
    p = ss.partial[3]
    rows = np.arange(p.start_idx, p.start_idx + len(p.f))
    cols = np.array([np.flatnonzero(G["tid_" + tag][r] == 3)[0] for r in rows])
    assert np.array_equal(np.asarray(p.f), G["f"][rows, cols]) and np.array_equal(np.asarray(p.mag), G["mag"][rows, cols])


@pytest.mark.parametrize("key", sorted(SYNTH))
def test_synth_with_edge_minframes_and_stretch(key):
    from pypevoc_b200 import SinSum
    kw = SYNTH[key]
    ss = SinSum(SIGNAL["sr"], nfft=PVKW["nfft"], hop=PVKW["hop"])
    for fr in range(G["f"].shape[0]):
        ss.add_frame(fr, G["f"][fr], G["mag"][fr], G["ph"][fr], realph=G["realph"][fr])
    w = ss.synth(SIGNAL["sr"], kw["hop"], edge=kw["edge"], minframes=kw["minframes"])
    assert w.shape == G[key].shape and w.dtype == np.float64
    assert pu.snr_db(w, G[key]) > 90.0
